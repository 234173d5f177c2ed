// pb200_device.cuh - device-side parameter block and per-pixel functions shared
// by the fused kernel and the function-granular kernels.
//
// "D:" = /root/reference/src/proteus/dswx_hls.py (nasa/PROTEUS v1.0.2).
//
// Numerics contract (DESIGN.md "Exactness"):
//  * band sums / differences wrap in int16 exactly like numpy int16 arrays
//    (D:1872-1884);
//  * float64(n)/float64(d) > t is evaluated with the exact integer form
//    p*b >= a*q (pb200_ratio_bound) - no floating point at all;
//  * awesh (D:1881) is a multiple of 0.25 -> compared as 4*awesh in int32;
//  * the terrain-shadow test follows numpy >= 2 semantics: float32 up to the
//    normalisation factor (no FMA contraction: explicit *_rn intrinsics and
//    -fmad=false), float64 for the dot product, the division and the compare.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb200 {

// ---- compact WTR-1/WTR-2 class code k = class & 7 -------------------------
// 0..4 = classes, 6 = ocean masked (254), 7 = fill (255)
__device__ __forceinline__ uint32_t expand_class(uint32_t k) {
    return k < 5u ? k : 248u + k;
}

enum : int { RB_WIGT = 0, RB_P1_MNDWI = 1, RB_P2_MNDWI = 2, RB_P1_NDVI = 3 };
enum : uint32_t { PF_AEROSOL = 1u, PF_COLLAPSE = 2u, PF_HISTOGRAM = 4u,
                  PF_DEFER_SNOW = 8u,     // 'cover' flow, phase 1: CLOUD stays the preliminary layer (+ aerosol bit), also on fill pixels
                  PF_NUMPY1 = 16u };      // terrain shadow with numpy 1.x promotion: float32 dot product / division (setup.py:78)

// fmask_lut entry layout
//   bits 0-2 : preliminary CLOUD value (0, 1, 4, 5)          D:1984-1991
//   bit  3   : snow bit (fmask & 16)                          D:2052
//   bits 4-8 : aerosol membership, bit (4 + k) for class k    D:1237
//   bit  15  : fmask == fmask_fill                            D:2204
struct DevParams {
    int32_t r_a[4], r_b[4];       // rational bounds, see pb200_ratio_bound
    int32_t awesh4_thr;           // 4*awesh > thr   <=> awesh > awgt
    int32_t p1_swir1, p1_nir;     // x < thr  (thr = ceil(threshold))
    int32_t p2_blue, p2_swir1, p2_swir2, p2_nir;
    int32_t lc_nir;               // nir > thr (thr = floor(threshold))
    int32_t band_fill[6];
    int32_t fmask_fill;
    uint32_t flags;               // PF_*
    float   dxf, dyf;             // float32(pixel_spacing_x), -|float32(spacing_y)|
    double  cos_thr, tan_thr;
    uint16_t fmask_lut[256];
    uint32_t diag_lut[32];        // binary-representation | k1 << 16
    uint32_t out_lut[128];        // [k2*16 + c] = WTR | BWTR<<8 | CONF<<16 | CLOUD<<24
    uint8_t  cls_lut[8];          // k -> output byte of WTR-1 / WTR-2 (collapsed or not)
};

struct TileDev {
    const int16_t *band[6];
    const uint8_t *fmask;
    const float   *dem;
    const uint8_t *land;
    const uint8_t *ocean;
    uint16_t *diag;
    uint8_t *wtr1, *wtr1r, *wtr2, *cloud, *shad, *wtr, *bwtr, *conf;
    unsigned long long *counters;
    int32_t height, width;
    int32_t dem_pitch, dem_rows, dem_off_y, dem_off_x;
    int32_t tiles_x, n_ctas;
    uint32_t flags;               // TF_*
    uint32_t pad_;
    double sx, sy, sz, sin_az, cos_az;
};
enum : uint32_t { TF_VEC = 1u, TF_TMA = 2u };

// ---------------------------------------------------------------------------
// exact comparison of float64(n)/float64(d) with a float64 threshold
// ---------------------------------------------------------------------------
struct Ratio { int p, q; };
__device__ __forceinline__ Ratio make_ratio(int n, int d) {
    Ratio r;
    r.p = d < 0 ? -n : n;
    r.q = d < 0 ? -d : d;
    return r;
}
// q > 0 guaranteed (clipped bands)
__device__ __forceinline__ bool ratio_gt(Ratio r, int a, int b) { return r.p * b >= a * r.q; }
__device__ __forceinline__ bool ratio_lt(Ratio r, int a, int b) { return r.p * b <= a * r.q; }
// q may be 0: n/0 = +-inf (n != 0) or nan (n == 0); nan compares false
__device__ __forceinline__ bool ratio_gt_any(Ratio r, int a, int b) {
    return (r.p * b >= a * r.q) && ((r.p | r.q) != 0);
}
__device__ __forceinline__ bool ratio_lt_any(Ratio r, int a, int b) {
    return (r.p * b <= a * r.q) && ((r.p | r.q) != 0);
}

__device__ __forceinline__ int wrap16(int x) { return (int)(short)x; }

// D:1840-1916.  Inputs are the (clipped) int16 band values widened to int.
template <bool MAY_DIV0>
__device__ __forceinline__ uint32_t diagnostic_tests(int B, int G, int R, int N, int S1, int S2,
                                                     const DevParams &P) {
    const Ratio mndwi = make_ratio(wrap16(G - S1), wrap16(G + S1));   // D:1872
    const Ratio ndvi  = make_ratio(wrap16(N - R),  wrap16(N + R));    // D:1884
    const int mbsrv = wrap16(G + R);                                  // D:1875
    const int mbsrn = wrap16(N + S1);                                 // D:1878
    const int awesh4 = 4 * B + 10 * G - 6 * mbsrn - S2;               // 4 x D:1881
    bool m0, m1, m2, nd;
    if (MAY_DIV0) {
        m0 = ratio_gt_any(mndwi, P.r_a[RB_WIGT], P.r_b[RB_WIGT]);
        m1 = ratio_gt_any(mndwi, P.r_a[RB_P1_MNDWI], P.r_b[RB_P1_MNDWI]);
        m2 = ratio_gt_any(mndwi, P.r_a[RB_P2_MNDWI], P.r_b[RB_P2_MNDWI]);
        nd = ratio_lt_any(ndvi, P.r_a[RB_P1_NDVI], P.r_b[RB_P1_NDVI]);
    } else {
        m0 = ratio_gt(mndwi, P.r_a[RB_WIGT], P.r_b[RB_WIGT]);
        m1 = ratio_gt(mndwi, P.r_a[RB_P1_MNDWI], P.r_b[RB_P1_MNDWI]);
        m2 = ratio_gt(mndwi, P.r_a[RB_P2_MNDWI], P.r_b[RB_P2_MNDWI]);
        nd = ratio_lt(ndvi, P.r_a[RB_P1_NDVI], P.r_b[RB_P1_NDVI]);
    }
    const bool t1 = m0;                                               // D:1893
    const bool t2 = mbsrv > mbsrn;                                    // D:1896
    const bool t3 = awesh4 > P.awesh4_thr;                            // D:1899
    const bool t4 = m1 && (S1 < P.p1_swir1) && (N < P.p1_nir) && nd;  // D:1902-1906
    const bool t5 = m2 && (B < P.p2_blue) && (S1 < P.p2_swir1) &&
                    (S2 < P.p2_swir2) && (N < P.p2_nir);              // D:1909-1914
    return (uint32_t)t1 | ((uint32_t)t2 << 1) | ((uint32_t)t3 << 2) |
           ((uint32_t)t4 << 3) | ((uint32_t)t5 << 4);
}

// D:4286-4317: the five test bits written as decimal digits; bit 5 -> 65535
__device__ __forceinline__ uint32_t binary_representation(uint32_t d) {
    if (d & 32u) return 65535u;
    return (d & 1u) + ((d >> 1) & 1u) * 10u + ((d >> 2) & 1u) * 100u +
           ((d >> 3) & 1u) * 1000u + ((d >> 4) & 1u) * 10000u;
}

// D:97-143 as bit planes over the 32 codes (bit i of plane j = bit j of the
// class of code i).  Used by the table builder and the function-level kernel.
__host__ __device__ __forceinline__ uint32_t interpreted_class(uint32_t d) {
    // closed form of the table: tests/test_oracle_golden.py and tests/test_gpu_parity.py check it against the
    // reference's own dict (tests/golden/reference_tables.json) through generate_interpreted_layer.
    constexpr uint32_t C1 = (1u << 0b01111) | (1u << 0b10111) | (1u << 0b11011) |
                            (1u << 0b11101) | (1u << 0b11110) | (1u << 0b11111);
    constexpr uint32_t C2 = (1u << 0b00111) | (1u << 0b01011) | (1u << 0b01101) |
                            (1u << 0b01110) | (1u << 0b10011) | (1u << 0b10101) |
                            (1u << 0b10110) | (1u << 0b11001) | (1u << 0b11010) |
                            (1u << 0b11100);
    constexpr uint32_t C3 = (1u << 0b11000);
    constexpr uint32_t C4 = (1u << 0b00011) | (1u << 0b00101) | (1u << 0b00110) |
                            (1u << 0b01001) | (1u << 0b01010) | (1u << 0b01100) |
                            (1u << 0b10000) | (1u << 0b10001) | (1u << 0b10010) |
                            (1u << 0b10100);
    if (d > 31u) return 255u;              // 32 = fill; anything else not in the table
    const uint32_t m = 1u << d;
    return (m & C1) ? 1u : (m & C2) ? 2u : (m & C3) ? 3u : (m & C4) ? 4u : 0u;
}

// D:1919-1993
__host__ __device__ __forceinline__ uint32_t preliminary_cloud(uint32_t fmask, int mode) {
    uint32_t c = (fmask & 8u) ? 1u : 0u;
    if (mode == 0 && (fmask & 4u)) c = 1u;
    if (fmask & 2u) c += 4u;
    return c;
}

// D:2089-2133 on full 8-bit values (function-level semantics: any uint8 in)
__host__ __device__ __forceinline__ uint32_t cloud_masking(uint32_t w2, uint32_t c) {
    uint32_t w = w2;
    if (c != 0u && c != 8u) w = 253u;
    if (c == 2u || c == 10u) w = 252u;
    if (w2 == 254u) w = 254u;
    if (w2 == 255u) w = 255u;
    return w;
}
// D:1710-1730
__host__ __device__ __forceinline__ uint32_t binary_water(uint32_t w) {
    return (w >= 1u && w <= 4u) ? 1u : w;
}
// D:1733-1837
__host__ __device__ __forceinline__ uint32_t confidence(uint32_t w2, uint32_t c) {
    if (w2 > 4u) return w2;
    const bool cloudy = (c < 16u) && (c & 5u);    // {1,3,4,5,6,7,9,11,12,13,14,15}
    if (cloudy) return w2 + 10u;
    if (c == 2u) return w2 + 20u;
    return w2;
}
// D:2578-2598
__host__ __device__ __forceinline__ uint32_t collapse_class(uint32_t w) {
    if (w == 0u) return 0u;
    if (w == 1u || w == 2u) return 1u;
    if (w == 3u || w == 4u) return 2u;
    if (w >= 252u) return w;
    return 255u;
}

// D:1305-1378 on full 8-bit values.  has_land / has_shad as in the reference's
// "is None" tests; shadow value 0 = masked (D:168).
__device__ __forceinline__ uint32_t landcover_shadow(uint32_t w1, int nir, bool has_land, uint32_t land,
                                                     bool has_shad, uint32_t shad, int lc_nir) {
    const bool water = (w1 >= 1u && w1 <= 4u);
    const bool psw = (w1 == 3u || w1 == 4u);
    bool kill = has_shad && shad == 0u && water && (!has_land || land != 200u);
    if (has_land) {
        const bool bright = nir > lc_nir;
        kill |= (land == 201u || land < 100u) && bright && psw;
        kill |= (land >= 100u && land < 200u) && water;
    }
    return kill ? 0u : w1;
}

// ---------------------------------------------------------------------------
// terrain shadow for one pixel from its four DEM neighbours  (D:4255-4281)
// g_col = d(dem)/d(col), g_row = d(dem)/d(row) already formed in float32.
// returns 1 = not shadow, 0 = shadow
// ---------------------------------------------------------------------------
struct SunTerms { double sx, sy, sz, sin_az, cos_az; };

// np1: numpy 1.x value-based casting (the reference pins numpy 1.23.5, setup.py:78): `float32_array * float64_scalar`
// stays float32, so the sun terms are rounded to float32 and every product, sum, the division, arccos / arctan and
// degrees run in float32 (cos_thr / tan_thr then hold the float32 decision boundaries).  Under numpy >= 2 (NEP 50) the
// same expressions promote to float64.
__device__ __forceinline__ uint32_t shadow_from_gradient(float g_col, float g_row, float dxf, float dyf,
                                                         const SunTerms &S, double cos_thr, double tan_thr,
                                                         bool np1 = false) {
    const float nx = __fdiv_rn(-g_col, dxf);                          // D:4260
    const float ny = __fdiv_rn(-g_row, dyf);                          // D:4261
    const float nf = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), 1.0f));  // D:4264
    if (np1) {
        const float s = __fadd_rn(__fmul_rn(nx, (float)S.sin_az), __fmul_rn(ny, (float)S.cos_az));     // D:4275-4277
        if (!(s <= (float)tan_thr)) return 1u;
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, (float)S.sx), __fmul_rn(ny, (float)S.sy)), (float)S.sz);
        const float x = __fdiv_rn(dot, nf);                                                            // D:4267-4271
        return (x >= (float)cos_thr && x <= 1.0f) ? 1u : 0u;
    }
    const double nxd = (double)nx, nyd = (double)ny, nfd = (double)nf;
    // directional slope: degrees(arctan(s)) <= min_slope  <=>  s <= tan_thr   D:4275-4279
    const double s = __dadd_rn(__dmul_rn(nxd, S.sin_az), __dmul_rn(nyd, S.cos_az));
    const bool backslope = s <= tan_thr;
    if (!backslope) return 1u;        // also taken for NaN (nan <= x is false -> ~False = True)
    // local incidence: degrees(arccos(x)) <= max_inc <=> cos_thr <= x <= 1   D:4267-4280
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(nxd, S.sx), __dmul_rn(nyd, S.sy)), S.sz);
    // x = dot / nfd (IEEE).  Decide without dividing when |dot - thr*nfd| is
    // far outside the rounding error of either side; divide only in the band.
    const double guard = 1e-12 * nfd;
    const double lo = __dadd_rn(dot, -__dmul_rn(cos_thr, nfd));
    const double hi = __dadd_rn(dot, -nfd);
    bool low_inc;
    if (fabs(lo) > guard && fabs(hi) > guard && fabs(cos_thr) <= 1.0) {
        low_inc = (lo > 0.0) && (hi < 0.0);
    } else {
        const double x = __ddiv_rn(dot, nfd);
        low_inc = (x >= cos_thr) && (x <= 1.0);
    }
    return low_inc ? 1u : 0u;
}

// streaming loads / stores: every raster element is touched exactly once
__device__ __forceinline__ int2 ldg_stream_v2(const void *p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u32(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ int4 ldg_stream_v4(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_v4(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void stg_stream_v2(void *p, uint32_t a, uint32_t b) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b));
}

}  // namespace pb200
