// pb200_sweep.cuh - brute-force check of the float32 terrain-shadow shortcuts against the exact float64 sequence.
//
// The fused kernels decide the shadow test of D:4264-4281 in float32 with a guard band (shadow_fast: float compares;
// shadow_fast1 / shadow_fast2: sign bits, FAST8) and fall back to shadow_exact (the reference's own operation
// sequence, numpy >= 2 promotion) only inside the band.  The claim "a pixel the shortcut DECIDES always gets the exact
// answer" is checked here on billions of DEM neighbourhoods (l, r, u, d) = the four float32 samples the stencil reads:
//   mode 0  gradients of random magnitude (log-uniform 1e-4 .. 1e5 m per pixel pair) and direction;
//   mode 1  neighbourhoods solved to sit ON the back-slope boundary s = tan_thr, then moved off it by a relative
//           2^-25 .. 2^-12 (both sides) - the band edge of the slope test (diff = +-e);
//   mode 2  the same for the incidence boundary x = cos_thr (D = +-eg), found by bisection in float64;
//   mode 3  special values: zeros, denormals, 1e30, +-inf, NaN, huge against tiny.
// Counted: samples decided by each shortcut, decided AND different from the exact sequence (must be 0), shadow pixels.
#pragma once
#include "pb200_fused.cuh"

namespace pb200 {

struct SweepCounts {
    unsigned long long n, shadow_exact;
    unsigned long long decided_cmp, wrong_cmp;        // shadow_fast (float compares; general variant)
    unsigned long long decided_sign, wrong_sign;      // shadow_fast1 (sign bits, scalar FFMA; FAST8)
    unsigned long long decided_sign2, wrong_sign2;    // shadow_fast2 (sign bits, packed FFMA2; FAST8)
};

__device__ __forceinline__ uint32_t sweep_rng(uint64_t &st) {          // splitmix64 -> 32 bits
    st += 0x9E3779B97F4A7C15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
__device__ __forceinline__ double sweep_u01(uint64_t &st) { return (sweep_rng(st) + 0.5) * (1.0 / 4294967296.0); }

__global__ void __launch_bounds__(256) shadow_sweep_kernel(const __grid_constant__ DevParams P, const __grid_constant__ FastParams F,
                                                           TileDev T, int mode, unsigned long long seed,
                                                           unsigned long long per_thread, SweepCounts *out) {
    // per-tile float32 constants exactly as the fused kernels derive them
    float K[SK_N];
    {
        const double kx = 0.5 / (double)P.dxf, ky = 0.5 / (double)P.dyf;
        K[SK_SA] = (float)(kx * T.sin_az); K[SK_CA] = (float)(ky * T.cos_az);
        K[SK_SX] = (float)(kx * T.sx); K[SK_SY] = (float)(ky * T.sy); K[SK_SZ] = (float)T.sz;
        K[SK_XX] = (float)(kx * kx);
        K[SK_EA] = 1e-6f * fabsf(K[SK_SA]); K[SK_EB] = 1e-6f * fabsf(K[SK_CA]);
        if (F.fast8) {      // same precondition check as the fused kernels make per tile
            const double hz = T.sx * T.sin_az + T.sy * T.cos_az, n2 = T.sx * T.sx + T.sy * T.sy + T.sz * T.sz;
            if (!(hz >= 0.0 && fabs(n2 - 1.0) < 1e-9 && fabs(T.sin_az * T.sin_az + T.cos_az * T.cos_az - 1.0) < 1e-9))
                K[SK_XX] = __int_as_float(0x7fffffff);
        }
    }
    const double kx = 0.5 / (double)P.dxf, ky = 0.5 / (double)P.dyf;   // s = a kx sin az + b ky cos az with a = l - r, b = u - d
    uint64_t st = seed ^ ((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0xD1B54A32D192ED03ull);
    SweepCounts c = {};
    for (unsigned long long it = 0; it < per_thread; ++it) {
        float l, r, u, d;
        const double base = (sweep_u01(st) - 0.3) * 6000.0;                       // elevations -1800 .. 4200 m
        if (mode == 3) {
            const float pool[12] = {0.0f, -0.0f, 1e-42f, -1e-40f, 1e30f, -1e30f, __int_as_float(0x7f800000), __int_as_float(0xff800000),
                                    __int_as_float(0x7fc00000), 3.4e38f, 1.0f, (float)base};
            l = pool[sweep_rng(st) % 12]; r = pool[sweep_rng(st) % 12]; u = pool[sweep_rng(st) % 12]; d = pool[sweep_rng(st) % 12];
        } else {
            const double mag = exp(log(1e-4) + sweep_u01(st) * (log(1e5) - log(1e-4)));
            const double phi = sweep_u01(st) * 6.283185307179586;
            double cphi = cos(phi), sphi = sin(phi);
            if (mode == 2 && cphi * kx * T.sin_az + sphi * ky * T.cos_az > 0.0) { cphi = -cphi; sphi = -sphi; }   // a ray of back slopes
            double a = mag * cphi, b = mag * sphi;
            if (mode == 1) {
                // on the slope boundary: a kx sin_az + b ky cos_az = tan_thr, solved for the component with the larger weight
                const double wa = kx * T.sin_az, wb = ky * T.cos_az;
                if (fabs(wa) >= fabs(wb)) a = (P.tan_thr - b * wb) / wa; else b = (P.tan_thr - a * wa) / wb;
            } else if (mode == 2) {
                // on the incidence boundary along the ray (a, b) = t (cos phi, sin phi): x(t) = dot / nf = cos_thr, bisection
                double lo = 0.0, hi = 1e7;
                auto xval = [&](double t) {
                    const double nx = t * cphi * kx, ny = t * sphi * ky;            // nx = -g_col / dx = (l - r) / (2 dx)
                    return (nx * T.sx + ny * T.sy + T.sz) / sqrt(nx * nx + ny * ny + 1.0);
                };
                const bool lo_above = xval(lo) >= P.cos_thr;
                if ((xval(hi) >= P.cos_thr) != lo_above) {
                    for (int i = 0; i < 60; ++i) {
                        const double mid = 0.5 * (lo + hi);
                        if ((xval(mid) >= P.cos_thr) == lo_above) lo = mid; else hi = mid;
                    }
                    a = lo * cphi; b = lo * sphi;
                }
            }
            if (mode != 0) {
                // step off the boundary by a relative 2^-25 .. 2^-12, either side
                const double rel = ldexp(1.0, -25 + (int)(sweep_rng(st) % 14)) * (sweep_u01(st) + 0.5);
                const double f = (sweep_rng(st) & 1u) ? 1.0 + rel : 1.0 - rel;
                if (sweep_rng(st) & 1u) a *= f; else b *= f;
            }
            r = (float)(base - 0.5 * a); l = (float)((double)r + a);
            d = (float)(base - 0.5 * b); u = (float)((double)d + b);
        }
        const uint32_t exact = shadow_exact(l, r, u, d, P, T);
        c.n += 1;
        c.shadow_exact += exact != 0u;
        {
            bool und = (F.fast_shadow_ok == 0u);
            const uint32_t got = shadow_fast(l, r, u, d, F, K, &und);
            if (!und) { c.decided_cmp += 1; c.wrong_cmp += got != exact; }
        }
        if (F.fast8) {
            uint32_t is, nt;
            float v;
            shadow_fast1(l, r, u, d, F, K, &is, &nt, &v);
            const uint32_t und = (~(is | nt) | (__float_as_uint(v + v + v + v) + 0x00800000u)) >> 31;
            if (!und) { c.decided_sign += 1; c.wrong_sign += (((is >> 31) ? BIG_SHADOWED : 0u) != exact); }
            uint32_t is2[2], nt2[2];
            float2 v2;
            shadow_fast2(make_float2(l, l), make_float2(r, r), make_float2(u, u), make_float2(d, d), F, K, is2, nt2, &v2);
            const uint32_t und2 = (~(is2[1] | nt2[1]) | (__float_as_uint(v2.y + v2.y + v2.x + v2.x) + 0x00800000u)) >> 31;
            if (!und2) { c.decided_sign2 += 1; c.wrong_sign2 += (((is2[1] >> 31) ? BIG_SHADOWED : 0u) != exact); }
        }
    }
    unsigned long long *o = reinterpret_cast<unsigned long long *>(out);
    const unsigned long long *v = reinterpret_cast<const unsigned long long *>(&c);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SweepCounts) / 8); ++i) {
        unsigned long long x = v[i];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&o[i], x);
    }
}

// ---------------------------------------------------------------------------
// FAST8 integer forms (one IDP.2A per rational test on per-pixel packs, wrapped-sum bookkeeping) against numpy's
// arithmetic, exhaustively over what the kernel can meet: every clipped (G, S1) in [1, 32767]^2 for the three MNDWI
// tests and every (N, R) in [1, 32767]^2 for NDVI - int16 sums wrapped like numpy's, quotient by IEEE float64 division
// (D:1872-1884, 1893-1914) - and 4*awesh on all (G, S1) with pseudo-random N, B, S2 (wrapped mbsrn included).
// counts: [0] mndwi > wigt, [1] mndwi > pswt_1_mndwi, [2] mndwi > pswt_2_mndwi, [3] ndvi < pswt_1_ndvi, [4] awesh > awgt
// mismatches; [5] pairs whose G + S1 wrapped (the sweep really visits them).
__global__ void fast8_sweep_kernel(const __grid_constant__ FastParams F, double wigt, double p1_mndwi, double p2_mndwi, double p1_ndvi,
                                   double awgt, unsigned long long *counts) {
    const uint32_t a = blockIdx.x + 1u;                                  // G (and N)
    unsigned int bad[5] = {0u, 0u, 0u, 0u, 0u}, wrapped = 0u;
    uint64_t st = 0x1234ull + a * 0x9E3779B97F4A7C15ull + threadIdx.x;
    for (uint32_t b = 1u + threadIdx.x; b <= 32767u; b += blockDim.x) {  // S1 (and R)
        const uint32_t rnd = sweep_rng(st), rnd2 = sweep_rng(st);
        const uint32_t nir = 1u + rnd % 32767u, blue = 1u + (rnd >> 16) % 32767u, swir2 = 1u + rnd2 % 32767u;
        // lane layout of the kernel: the pixel in the LOW half of every packed register
        const uint32_t gs = (a + b) & 0xffffu, gd = (b - a) & 0xffffu, ns = (nir + b) & 0xffffu;
        const uint32_t nrs = (a + b) & 0xffffu;                           // (N, R) = (a, b) for the NDVI form
        const Fast8Signs m = fast8_signs_wrapped(gs, gd, ns, nrs, nir, 0u, blue, swir2, 0u, F);    // MNDWI forms + awesh
        const Fast8Signs v = fast8_signs_wrapped(0u, 0u, 0u, nrs, a, b, 0u, 0u, 0u, F);             // NDVI form
        const double q16 = (double)(short)gs;
        const double mndwi = __ddiv_rn((double)((int)a - (int)b), q16);  // (G - S1) / wrap16(G + S1)
        const double ndvi = __ddiv_rn((double)((int)a - (int)b), q16);   // (N - R) / wrap16(N + R): the same numbers
        bad[0] += (m.x0 < 0) != (mndwi > wigt);
        bad[1] += (m.x1 < 0) != (mndwi > p1_mndwi);
        bad[2] += (m.x2 < 0) != (mndwi > p2_mndwi);
        bad[3] += (v.x3 < 0) != (ndvi < p1_ndvi);
        const double awesh = (double)blue + 2.5 * (double)a - 1.5 * (double)(short)ns - 0.25 * (double)swir2;   // D:1881
        bad[4] += (m.aw < 0) != (awesh > awgt);
        wrapped += (gs & 0x8000u) != 0u;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const unsigned int t = __reduce_add_sync(0xffffffffu, bad[k]);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&counts[k], (unsigned long long)t);
    }
    wrapped = __reduce_add_sync(0xffffffffu, wrapped);
    if ((threadIdx.x & 31) == 0 && wrapped) atomicAdd(&counts[5], (unsigned long long)wrapped);
}

}  // namespace pb200
