// probe: which TMA box configurations work on this box.  usage: tma_probe BOXW BOXH L2PROMO [bulk1d]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
struct __align__(128) Sm { float dem[40 * 256]; unsigned long long mbar; };
__global__ void k(const __grid_constant__ CUtensorMap pmap, const float *src, int bulk, int bytes, int c0, int c1, float *out, int n) {
    __shared__ Sm s;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s.mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s.mbar)), "r"((unsigned)bytes) : "memory");
        if (bulk)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(smem_u32(&s.dem[0])), "l"(src), "r"(bytes), "r"(smem_u32(&s.mbar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(smem_u32(&s.dem[0])), "l"(&pmap), "r"(c0), "r"(c1), "r"(smem_u32(&s.mbar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&s.mbar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s.dem[i];
}
int main(int argc, char **argv) {
    const int bw = atoi(argv[1]), bh = atoi(argv[2]), promo = atoi(argv[3]), c0 = atoi(argv[4]), c1 = atoi(argv[5]), bulk = argc > 6;
    const int rows = 164, pitch = 228;
    std::vector<float> h(rows * pitch);
    for (int i = 0; i < rows * pitch; ++i) h[i] = (float)i;
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 40 * 256 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    alignas(64) CUtensorMap m; memset(&m, 0, sizeof(m));
    cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    std::vector<float> o(bw * bh);
    k<<<1, 128>>>(m, d, bulk, bw * bh * 4, c0, c1, out, bw * bh);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    printf("c0 %d c1 %d box %dx%d promo %d bulk %d: encode %d, %s, out[0]=%g (want %g) out[last]=%g (want %g)\n", c0, c1, bw, bh, promo, bulk, (int)r,
           cudaGetErrorString(e), o[0], bulk ? h[0] : h[c1 * pitch + c0], o[bw * bh - 1],
           bulk ? h[bw * bh - 1] : h[(c1 + bh - 1) * pitch + c0 + bw - 1]);
    return 0;
}
