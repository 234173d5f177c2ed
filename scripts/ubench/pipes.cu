// Throughput of the integer / packed-float instructions the fused kernel leans on, alone and in pairs (which ones share
// a pipe?), on one B200.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITERS 2048
#define N_ACC 8
enum { OP_LOP3, OP_SHF, OP_PRMT, OP_VADD2, OP_VIADDMNMX, OP_IMAD, OP_IMADHI, OP_IDP2A, OP_IDP4A, OP_FFMA, OP_FFMA2, OP_IADD3, OP_LDS, N_OPS };
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == OP_VADD2) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (OP == OP_VIADDMNMX) d = __viaddmax_s16x2(a, b, c);
    else if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_IDP2A) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_IDP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == OP_IADD3) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else d = a;
    return d;
}
template <int A, int B, int RA, int RB>
__global__ void __launch_bounds__(1024, 1) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t acc[N_ACC], bcc[N_ACC];
    unsigned long long f2[N_ACC];
    for (int i = 0; i < N_ACC; ++i) { acc[i] = seed * (i + 1) + threadIdx.x; bcc[i] = seed ^ (i * 77 + threadIdx.x); f2[i] = (unsigned long long)acc[i] << 32 | bcc[i]; }
    const uint32_t kb = seed | 0x01020304u, kc = seed + 12345u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N_ACC; ++i) {
#pragma unroll
            for (int r = 0; r < RA; ++r) {
                if (A == OP_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[i]) : "l"(f2[(i + 1) % N_ACC]), "l"(f2[(i + 2) % N_ACC]));
                else acc[i] = op<A>(acc[i], kb, kc);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (B == OP_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f2[i]) : "l"(f2[(i + 1) % N_ACC]), "l"(f2[(i + 2) % N_ACC]));
                else bcc[i] = op<B>(bcc[i], kb, kc);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < N_ACC; ++i) s += acc[i] ^ bcc[i] ^ (uint32_t)f2[i] ^ (uint32_t)(f2[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
static const char *NAMES[] = {"LOP3", "SHF", "PRMT", "VADD2", "VIADDMNMX", "IMAD", "IMAD.HI", "IDP.2A", "IDP.4A", "FFMA", "FFMA2", "IADD", "-"};
template <int A, int B, int RA, int RB> void run(uint32_t *out, long long *cyc) {
    k<A, B, RA, RB><<<148, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    k<A, B, RA, RB><<<148, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double n_a = (double)ITERS * N_ACC * RA * 1024, n_b = (double)ITERS * N_ACC * RB * 1024;
    printf("%-10s x%d + %-10s x%d : %8.1f thread-instr/clk/SM  (A %.1f, B %.1f)  [%s]\n", NAMES[A], RA, NAMES[B], RB, (n_a + n_b) / c, n_a / c, n_b / c,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<OP_LOP3, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_SHF, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_PRMT, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_VADD2, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_VIADDMNMX, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_IADD3, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_IMAD, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_IMADHI, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_IDP2A, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_IDP4A, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_FFMA, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_FFMA2, N_OPS - 1, 1, 0>(out, cyc);
    run<OP_LOP3, OP_IMAD, 1, 1>(out, cyc);
    run<OP_LOP3, OP_IDP2A, 1, 1>(out, cyc);
    run<OP_LOP3, OP_IMADHI, 1, 1>(out, cyc);
    run<OP_IMAD, OP_IDP2A, 1, 1>(out, cyc);
    run<OP_FFMA, OP_IDP2A, 1, 1>(out, cyc);
    run<OP_LOP3, OP_FFMA2, 1, 1>(out, cyc);
    run<OP_IMAD, OP_FFMA2, 1, 1>(out, cyc);
    run<OP_LOP3, OP_VADD2, 1, 1>(out, cyc);
    run<OP_LOP3, OP_IDP2A, 2, 1>(out, cyc);
    run<OP_LOP3, OP_IMAD, 2, 1>(out, cyc);
    run<OP_IDP2A, OP_IDP4A, 1, 1>(out, cyc);
    return 0;
}
