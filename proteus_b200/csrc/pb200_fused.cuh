// pb200_fused.cuh - K1 fast path: persistent, packed-SIMD fused classification.
//
// Same per-pixel function as dswx_fused_kernel<> in pb200_kernels.cuh (which
// stays as the generic path for rasters whose width is not a multiple of 4,
// whose planes are not 16-byte aligned or whose DEM cannot be addressed by
// TMA), re-organised for instruction issue - the resource that bounds this
// pass on B200 once the memory side streams (profiles/):
//
//  * a lane owns 4 consecutive pixels of a row (one 8-byte load per band - row
//    pitch 3660 * 2 B is only 8-byte aligned - and 4-byte loads / stores of the
//    byte rasters); one CTA of 24 warps at <= 80 registers per SM: the row
//    loop does not spill, and the loads of the next row are issued in the middle
//    of the current one (an 8-pixel lane with 16-byte accesses needed 123
//    registers -> 16 warps per SM and measured 60 % issue utilisation, profiles/);
//  * int16 bands stay packed two pixels per register: fill test, clip, the
//    wrapping sums and the integer threshold tests run on the 16x2 SIMD
//    integer pipe (VIMNMX[3].x16x2, VIADD.16x2, VIADDMNMX.S16x2); 4*awesh is
//    four IDP.2A dot products per pixel;
//  * the exact rational tests p*b >= a*q need 32-bit products: two IMADs per
//    threshold, arranged so that the SIGN bit carries the result; LOP3 and
//    funnel shifts assemble the 5-bit diagnostic code;
//  * everything after the code is table look-ups in shared memory:
//      diag_lut[code | valid<<5 | not-ocean<<6]      -> DIAG value, WTR-1 class
//      fk_lut  [fmask | class<<8 | (nir<=1000)<<11]  -> remapped class + CLOUD bits
//      big_lut [that | land class | bright | shadow] -> WTR | BWTR | CONF | flags
//  * the terrain-shadow test runs only for lanes holding a pixel whose result
//    it can change (flag from big_lut) or when SHAD is requested: float32 with
//    a guard band covering every rounding error of the shortcut, and the
//    float64 reference sequence (shadow_from_gradient) inside the band;
//  * persistent CTAs loop over 128 x 96 pixel items (a warp owns 4 rows); an item's DEM tile with halo arrives
//    as one TMA box into a double buffer, requested one item ahead; warps synchronise through full / empty
//    mbarriers only (no __syncthreads between the items of a tile); the tables are loaded once per CTA.
#pragma once
#include "pb200_kernels.cuh"

namespace pb200 {

// Geometry: ONE persistent CTA per SM works on 128 x FT_H pixel items.  The lean variant (graded layers) runs 24
// warps x 4 rows at <= 80 registers, the full variant (all layers, needs ~118 registers) 16 warps x 6 rows.  Measured
// (profiles/README.md): one 24-warp CTA beats 3 x 8 and 2 x 12 warps - one copy of the tables, 98/96 instead of 34/32
// DEM rows per item, and more of the SM's 256 KB left to L1, which the streaming loads in flight need.
#ifndef PB200_FT_H
#define PB200_FT_H 96
#endif
#ifndef PB200_FT_WARPS_LEAN
#define PB200_FT_WARPS_LEAN 24
#endif
#ifndef PB200_FT_WARPS_FULL
#define PB200_FT_WARPS_FULL 16
#endif
#ifndef PB200_FT_PRODUCER_WARP
#define PB200_FT_PRODUCER_WARP 0   // DEM requests by thread 0 (0) or by warp 0 with lane 0 issuing (1)
#endif
#ifndef PB200_FT_LATE_REQUEST
#define PB200_FT_LATE_REQUEST 1    // request the DEM tile of item k + 1 in the middle of item k instead of at its top
#endif
#ifndef PB200_FT_PATCH_SLOW
#define PB200_FT_PATCH_SLOW 1      // evaluate the packed fast path unconditionally and patch wrapped pairs afterwards
#endif
#ifndef PB200_FT_LAND_DP4A
#define PB200_FT_LAND_DP4A 0       // land_lut address as ONE IDP.4A (byte select + base add) instead of shift, mask, add
#endif
#ifndef PB200_FT_SHADOW_ALWAYS
#define PB200_FT_SHADOW_ALWAYS 0   // FAST8: evaluate the shadow shortcut for every lane of a tile with a DEM
#endif
#ifndef PB200_FT_SHADOW_SCALAR
#define PB200_FT_SHADOW_SCALAR 0   // FAST8 shadow shortcut with scalar FFMA (1) or packed FFMA2 (0)
#endif
#ifndef PB200_FT_XITEM_PREFETCH
#define PB200_FT_XITEM_PREFETCH 1  // request the first row of the next item in the last row of the current one
#endif
constexpr int FT_W = 128;          // item width: one warp = 32 lanes x 4 pixels
constexpr int FT_H = PB200_FT_H;   // item height
template <bool FULL> struct FtGeom {
    static constexpr int WARPS = FULL ? PB200_FT_WARPS_FULL : PB200_FT_WARPS_LEAN;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int ROWS_PER_WARP = FT_H / WARPS;
    static_assert(FT_H % WARPS == 0, "a warp owns a whole number of rows of an item");
};
// DEM staging: one TMA box per item (a box is at most 256 elements wide and 256
// rows high).  Box start = dem_off_x + x0 - padx with padx = 4 +
// (dem_off_x & 3): a multiple of 4 floats (UTMALDG needs a 16-byte aligned box
// start on B200, scripts/tma_probe.cu), one-column halo on each side.
constexpr int FT_SMW = 136;        // 7 (max padx) + 128 + 1 = 136 floats = 544 B
constexpr int FT_SMH = FT_H + 2;

// land classes as the kill table sees them
enum : uint32_t { LC_NONE = 0, LC_WATER = 1, LC_EVERGREEN_OR_LOW = 2, LC_HIGH = 3 };
// flags byte (bits 24..31) of a big_lut entry
enum : uint32_t { FL_VALID = 1u << 24, FL_CLOUD_VALID = 1u << 25, FL_SHADOW_SENSITIVE = 1u << 26, FL_BIN_SHIFT = 28 };

// big_lut index: kb | c<<3 | cat<<7 | shadowed<<9 | bright<<10, with "shadowed" and "bright" also XORed into bits 3
// and 4 and (fk_lut) the class k1 XORed into bits 2-4 of fmask | k1<<8 | nle<<11: the bits that differ between
// neighbouring pixels then select the bank, the bits that are spatially coherent select the row (44 % of the
// shared wavefronts were bank-conflict replays before, profiles/).  XOR with a function of the upper bits is a
// bijection, the host builds the tables in the same order.
#ifndef PB200_FT_SWIZZLE_BRIGHT
#define PB200_FT_SWIZZLE_BRIGHT 0   // measured: the extra IMAD per pair costs what the fewer bank conflicts give
#endif
constexpr uint32_t BIG_SHADOWED = 0x208u, BIG_BRIGHT = PB200_FT_SWIZZLE_BRIGHT ? 0x410u : 0x400u;
struct FusedTables {               // built on the host per plan, global memory
    uint32_t big_lut[2048];        // [kb | c<<3 | cat<<7 | shadowed<<9 | bright<<10] -> WTR | BWTR<<8 | CONF<<16 | flags<<24
    uint32_t diag_lut[128];        // [code | valid<<5 | not_ocean<<6] -> DIAG value | (k1<<8)<<16
    uint8_t  fk_lut[4096];         // [fmask | k1<<8 | (nir<=1000)<<11] -> kb | c<<3 | (kb is a water class)<<7
    uint8_t  land_lut[256];        // land value -> cat
    uint8_t  kill_lut[128];        // [kb | shadowed<<3 | bright<<4 | cat<<5] -> k2 (optional layers only)
};
static_assert(sizeof(FusedTables) % 16 == 0, "tables are copied with 16-byte loads");

struct ItemDesc { uint32_t tile; uint16_t tx, ty; };

// packed constants of the fast path (kernel parameter -> constant bank)
struct FastParams {
    uint32_t fill_xor[6], fill_or[6];   // x = (raw ^ fill_xor) | fill_or ; x half == 0 <=> raw == fill
    uint32_t fmask_xor4, fmask_or;      // same for the Fmask bytes (4 per word); or-mask on 16-bit halves
    uint32_t m_p1swir1, m_p1nir;        // packed (-threshold), threshold clamped to [1, 32768]
    uint32_t m_p2blue, m_p2swir1, m_p2swir2, m_p2nir;
    uint32_t m_nle, m_lc;               // packed -(1000 + 1), -(lcmask + 1)
    int32_t  awesh_init;                // floor(4*awgt): sign(init - 4*awesh) <=> awesh > awgt
    int32_t  ra[4], rb[4];              // ">=" forms (generic kernel, slow path)
    // strict forms for the fast path (no "- 1": one IMAD less per test):
    //   RN(n/d) > t  <=>  p/q >  sa/sb  <=>  sa*q - p*sb < 0   (sa/sb = largest fraction <= midpoint)
    //   RN(n/d) < t  <=>  p/q <  sa/sb  <=>  p*sb - sa*q < 0   (sa/sb = smallest fraction >= midpoint)
    int32_t  sa[4], nsb[4];             // nsb = -sb for the ">" tests, +sb for the "<" test; sa negated for "<"
    float    kx, ky;                    // 0.5 / dxf, 0.5 / dyf
    float    tan32, e0, cc32;           // float32(tan_thr), 1e-6*|tan_thr| + 1e-30, c*|c| with c = cos_thr
    uint32_t fast_shadow_ok;            // thresholds are finite and |cos_thr| <= 1
    uint32_t any_nofill;                // some band has no fill value (fill_or != 0)
    // ---- FAST8 variant (parameters shaped like the defaults; build_fast_params decides) --------------------------
    uint32_t nfill[6];                  // packed -fill: (raw + nfill) half == 0 <=> raw == fill
    // the rational tests as ONE dp2a each on a per-pixel pack: sign(b0 * lo16 + b1 * hi16) <=> test true
    uint32_t c_wigt;                    // on (gs, dg), dg = S1 - G: unsigned bytes (wigt > 0: both coefficients positive)
    uint32_t c_p1, c_p2;                // on (gs, dg): signed bytes
    uint32_t c_ndvi;                    // on (N, R): signed bytes
    uint32_t c_aw_gd, c_aw_nr;          // 4*awesh terms on the same packs: -2 gs + 8 dg (= -10 G + 6 S1), 6 N
    uint32_t c_aw_b, c_aw_s2;           // -4 B and + S2 on the two-pixel registers (byte 0 for .lo, byte 3 for .hi)
    float    sh_c7;                     // e' = sh_c7 * v + e0 >= 1e-6 (|t1| + |t2|) + e0   (|t1| + |t2| <= v / sqrt 2)
    float    ncc_hi, ncc_lo;            // -(c|c| + 4e-6), -(c|c| - 4e-6)
    uint32_t fast8;                     // every FAST8 precondition holds
};

struct __align__(128) DemHalf { float v[FT_SMH][FT_SMW]; };   // TMA destination: 128-B aligned
constexpr uint32_t DEM_BOX_BYTES = FT_SMH * FT_SMW * sizeof(float);

struct __align__(128) FastSmem {
    DemHalf dem[2];                     // double buffer: the DEM tile of item k + 1 streams in while item k is classified
    uint32_t big_lut[2048];
    uint32_t diag_lut[128];
    uint8_t  fk_lut[4096];
    uint8_t  land_lut[256];
    uint8_t  kill_lut[128];
    TileDev  tile;                      // descriptor of the current tile
    float    sun32[12];                 // SK_* constants of the current tile
    unsigned long long full[2];         // TMA transaction barriers, one per DEM buffer
    unsigned long long empty[2];        // one arrival per warp when it is done with an item's buffer
};

__device__ __forceinline__ int sext_lo(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int sext_hi(uint32_t x) { return ((int)x) >> 16; }


// float32 shortcut of the shadow test (branch-free).  Returns 0x200 (the big_lut
// index bit) when the pixel is CERTAINLY in terrain shadow, 0 when it is
// certainly not, and sets *undecided when a value is too close to a decision
// boundary (or not finite): the caller then runs the exact float64 sequence.
// Error budget in DESIGN.md section 3.
// per-tile float32 constants of the shortcut (FastSmem::sun32), from the float64 sun terms:
enum { SK_SA = 0, SK_CA, SK_SX, SK_SY, SK_SZ, SK_XX, SK_EA, SK_EB, SK_N };
//   kx*sin_az, ky*cos_az, kx*sx, ky*sy, sz, kx*kx (= ky*ky: square pixels, else fast_shadow_ok = 0),
//   1e-6*|kx*sin_az|, 1e-6*|ky*cos_az|   (kx = 0.5/dx, ky = 0.5/dy)
__device__ __forceinline__ uint32_t shadow_fast(float l, float r, float u, float d, const FastParams &F,
                                                const float (&K)[SK_N], bool *undecided) {
    const float a = l - r, b = u - d;                     // -2 g_col, -2 g_row: the reference's own float32 differences
    const float diff = fmaf(a, K[SK_SA], fmaf(b, K[SK_CA], -F.tan32));          // ~ s - tan_thr
    const float e = fmaf(fabsf(a), K[SK_EA], fmaf(fabsf(b), K[SK_EB], F.e0));   // 1e-6 (|t1| + |t2|) + e0
    const float dot = fmaf(a, K[SK_SX], fmaf(b, K[SK_SY], K[SK_SZ]));
    const float v = fmaf(fmaf(a, a, b * b), K[SK_XX], 1.0f);                    // ~ nf^2 = 1 + k^2 (a^2 + b^2)
    const float L = dot * fabsf(dot);                     // t -> t|t| is monotone: x >= c for any sign of c
    const float D = fmaf(-F.cc32, v, L);
    const float eg = 4e-6f * v;
    // shadow      <=> back slope AND NOT low incidence:  diff < -e  and  D < -eg
    // not shadow  <=> not a back slope, or low incidence with x <= 1:  diff > e  or  (D > eg and L < 0.99999 v)
    // NaN / inf make every comparison false -> neither -> undecided (unless the slope test alone decides,
    // which is exact: "not a back slope" never looks at the incidence angle)
    const bool is_shadow = (diff < -e) && (D < -eg);
    const bool not_shadow = (diff > e) || ((D > eg) && (L < 0.99999f * v));
    *undecided = *undecided || !(is_shadow || not_shadow);
    return is_shadow ? BIG_SHADOWED : 0u;
}

// dp2a with signed 16-bit halves and UNSIGNED bytes (IDP.2A.LO.S16.U8): c + lo16(a) * b.byte0 + hi16(a) * b.byte1
__device__ __forceinline__ int dp2a_lo_s16_u8(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// FAST8 flavour of the shortcut for TWO pixels at once on the packed float32 pipe (FFMA2 / FMUL2 / FADD2: one
// instruction, two pixels; scalar operands are broadcast), with every decision read off SIGN BITS instead of float
// compares (no FSETP / PLOP3).  Same quantities and the same error budget as shadow_fast, with
//   * e' = c7 * v + e0 >= e: |t1| + |t2| <= k (|a| + |b|) <= sqrt(2) k sqrt(a^2 + b^2) = sqrt(2 (v - 1)) <= v / sqrt(2);
//   * dot^2 against (c|c| -+ 4e-6) v instead of dot |dot| - c|c| v against -+ 4e-6 v, the sign of dot taken separately
//     (FAST8 requires cos_thr > 0.01: when dot^2 > c^2 v, |dot| > 0.01 and its float32 sign is exact);
//   * no "x <= 1" test: FAST8 requires tan_thr <= -0.005, and a back slope s <= tan_thr < 0 has |n_xy| >= |tan_thr|,
//     so x = dot / nf <= 1 / sqrt(1 + tan_thr^2) < 0.99999 (dot <= cos(zen) - sin(zen) |tan_thr| <= 1).
// is / nt: bit 31 set <=> certainly shadow / certainly not shadow; non-finite inputs make v non-finite, which the
// caller detects on the sum of the four v's.
__device__ __forceinline__ void shadow_fast2(float2 l, float2 r, float2 u, float2 d, const FastParams &F,
                                             const float (&K)[SK_N], uint32_t (&is)[2], uint32_t (&nt)[2], float2 *v_out) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    const float2 a = __ffma2_rn(r, neg1, l), b = __ffma2_rn(d, neg1, u);      // l - r, u - d: exact float32 differences
    const float2 diff = __ffma2_rn(a, make_float2(K[SK_SA], K[SK_SA]),
                                   __ffma2_rn(b, make_float2(K[SK_CA], K[SK_CA]), make_float2(-F.tan32, -F.tan32)));
    const float2 dot = __ffma2_rn(a, make_float2(K[SK_SX], K[SK_SX]),
                                  __ffma2_rn(b, make_float2(K[SK_SY], K[SK_SY]), make_float2(K[SK_SZ], K[SK_SZ])));
    const float2 v = __ffma2_rn(__ffma2_rn(a, a, __fmul2_rn(b, b)), make_float2(K[SK_XX], K[SK_XX]), make_float2(1.0f, 1.0f));
    const float2 e = __ffma2_rn(v, make_float2(F.sh_c7, F.sh_c7), make_float2(F.e0, F.e0));
    const float2 p1 = __fadd2_rn(diff, e);                                    // < 0 <=> diff < -e   (back slope)
    const float2 q2 = __ffma2_rn(diff, neg1, e);                              // < 0 <=> diff >  e   (not a back slope)
    const float2 dd = __fmul2_rn(dot, dot);
    const float2 qlo = __ffma2_rn(v, make_float2(F.ncc_lo, F.ncc_lo), dd);    // < 0 <=> |x| certainly below cos_thr
    const float2 qhi = __ffma2_rn(v, make_float2(F.ncc_hi, F.ncc_hi), dd);    // >= 0 <=> |x| certainly above cos_thr
    const uint32_t p1b[2] = {__float_as_uint(p1.x), __float_as_uint(p1.y)}, q2b[2] = {__float_as_uint(q2.x), __float_as_uint(q2.y)};
    const uint32_t lob[2] = {__float_as_uint(qlo.x), __float_as_uint(qlo.y)}, hib[2] = {__float_as_uint(qhi.x), __float_as_uint(qhi.y)};
    const uint32_t dtb[2] = {__float_as_uint(dot.x), __float_as_uint(dot.y)};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        // shadow: back slope and (|x| < thr, or |x| > thr with dot < 0);  lit: no back slope, or |x| > thr with dot > 0
        is[i] = p1b[i] & (lob[i] | (~hib[i] & dtb[i]));
        nt[i] = q2b[i] | ~(hib[i] | dtb[i]);
    }
    *v_out = v;
}

// The same decision for ONE pixel with scalar FFMA (full rate on B200: two FFMA issue faster than one FFMA2 in a mixed
// instruction stream, scripts/ubench/pipes2.cu)
__device__ __forceinline__ void shadow_fast1(float l, float r, float u, float d, const FastParams &F, const float (&K)[SK_N],
                                             uint32_t *is, uint32_t *nt, float *v_out) {
    const float a = l - r, b = u - d;
    const float diff = fmaf(a, K[SK_SA], fmaf(b, K[SK_CA], -F.tan32));
    const float dot = fmaf(a, K[SK_SX], fmaf(b, K[SK_SY], K[SK_SZ]));
    const float v = fmaf(fmaf(a, a, b * b), K[SK_XX], 1.0f);
    const float e = fmaf(v, F.sh_c7, F.e0);
    const uint32_t p1 = __float_as_uint(diff + e), q2 = __float_as_uint(e - diff);
    const float dd = dot * dot;
    const uint32_t lo = __float_as_uint(fmaf(v, F.ncc_lo, dd)), hi = __float_as_uint(fmaf(v, F.ncc_hi, dd));
    const uint32_t dt = __float_as_uint(dot);
    *is = p1 & (lo | (~hi & dt));
    *nt = q2 | ~(hi | dt);
    *v_out = v;
}

// Shared-memory reads through an explicit 32-bit shared address: the base is computed once per thread
// (an opaque register), instead of being re-derived from SR_CgaCtaId in front of every group of
// register-indexed look-ups (S2R / S2UR + LEA, profiles/).  The volatile forms are for data that changes
// between items (tile descriptor, DEM tile); the tables are read-only after the prologue.
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t lds_tab32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_tab8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
    float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
template <typename T>
__device__ __forceinline__ T *lds_ptr(uint32_t a) {
    unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return reinterpret_cast<T *>(v);
}
#define FS_OFF(member) ((uint32_t)offsetof(FastSmem, member))
#define FS_TILE(member) ((uint32_t)(offsetof(FastSmem, tile) + offsetof(TileDev, member)))

// address of element `pix` of a plane: one IMAD.WIDE (FMA pipe) instead of a 64-bit add pair on the ALU pipe
template <typename T>
__device__ __forceinline__ T *plane_at(T *base, uint32_t pix) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(pix), "n"((int)sizeof(T)), "l"(base));
    return reinterpret_cast<T *>(r);
}

// Rare paths kept out of line so that the hot loop stays small.
__device__ __noinline__ uint32_t diag_pixel_slow(uint32_t B, uint32_t G, uint32_t R, uint32_t N, uint32_t S1,
                                                 uint32_t S2, uint32_t hi, const DevParams &P) {
    // some int16 sum of this pixel wrapped: scalar evaluation with sign normalisation (D:1872-1914); hi selects
    // the pixel of the pair
    const uint32_t sh = hi ? 16u : 0u;
    return diagnostic_tests<false>((int)(short)(B >> sh), (int)(short)(G >> sh), (int)(short)(R >> sh),
                                   (int)(short)(N >> sh), (int)(short)(S1 >> sh), (int)(short)(S2 >> sh), P);
}
__device__ __noinline__ uint32_t shadow_exact(float l, float r, float u, float d, const DevParams &P,
                                              const TileDev &T) {
    SunTerms S;
    S.sx = T.sx; S.sy = T.sy; S.sz = T.sz; S.sin_az = T.sin_az; S.cos_az = T.cos_az;
    const float g_col = __fmul_rn(__fsub_rn(r, l), 0.5f);                     // D:4255
    const float g_row = __fmul_rn(__fsub_rn(d, u), 0.5f);
    return shadow_from_gradient(g_col, g_row, P.dxf, P.dyf, S, P.cos_thr, P.tan_thr) ? 0u : BIG_SHADOWED;
}

// ---------------------------------------------------------------------------
// ALL_GRADED: the tile batch writes all four graded layers (the product's configuration): no pointer tests in the row
// loop.  The lean kernel is also instantiated without it for subsets such as DIAG + WTR (BASELINE configs[0]); a
// run-time test in the one instantiation measured 2.5 % slower on the full product.
// FAST8: parameters shaped like the defaults (FastParams::fast8): fill test as one add-min chain, each rational test
// as one IDP.2A on a per-pixel pack, the terrain-shadow shortcut on the packed float32 pipe with sign-bit decisions.
template <bool OPTIONAL_LAYERS, bool ALL_GRADED = false, bool FAST8 = false>
__global__ void __launch_bounds__(FtGeom<OPTIONAL_LAYERS>::THREADS, 1)
dswx_fused_fast_kernel(const TileDev *__restrict__ tiles, const CUtensorMap *__restrict__ tmaps,
                       const FusedTables *__restrict__ tables, const ItemDesc *__restrict__ items, int n_items,
                       const __grid_constant__ DevParams P, const __grid_constant__ FastParams F) {
    extern __shared__ __align__(128) unsigned char smem_raw[];   // 51 KB: dynamic (every access goes through `sb`)
    FastSmem &s = *reinterpret_cast<FastSmem *>(smem_raw);
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(&s);    // shared address of the block, kept in a register
    asm volatile("mov.b32 %0, %0;" : "+r"(sb));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int FT_THREADS = FtGeom<OPTIONAL_LAYERS>::THREADS, FT_ROWS_PER_WARP = FtGeom<OPTIONAL_LAYERS>::ROWS_PER_WARP;
    const int rgrp = warp;                                // which group of rows of the item

    // ---- tables: once per CTA -------------------------------------------------
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables);
        uint4 *dst = reinterpret_cast<uint4 *>(s.big_lut);
        static_assert(offsetof(FastSmem, kill_lut) - offsetof(FastSmem, big_lut) + 128 == sizeof(FusedTables),
                      "smem table block mirrors FusedTables");
        static_assert(offsetof(FastSmem, big_lut) % 16 == 0, "16-byte table copies");
        for (int i = tid; i < (int)(sizeof(FusedTables) / 16); i += FT_THREADS) dst[i] = __ldg(src + i);
        if (tid == 0) {
            mbar_init(&s.full[0], 1); mbar_init(&s.full[1], 1);
            mbar_init(&s.empty[0], FT_THREADS / 32); mbar_init(&s.empty[1], FT_THREADS / 32);
        }
    }
    uint32_t cur_tile = 0xffffffffu;
    // DEM pipeline state, identical in every thread: bit b = parity of the next phase of full[b] (toggled after each
    // item of a tile with a DEM that used buffer b), bit 2 + b = buffer b has carried a transaction
    uint32_t fstate = 0;
    // counters accumulate in registers across the items of a tile and are flushed once per warp
    // when the CTA moves to another tile (and at the end): valid | cloud-and-valid << 16, not-ocean
    uint32_t acc_vc = 0, acc_nno = 0;
    unsigned long long acc_hist = 0ull;
    auto flush_counters = [&]() {
        // the descriptor in shared memory is still that of the tile being left (it is replaced after the flush)
        unsigned long long *cur_counters = (cur_tile != 0xffffffffu) ? lds_ptr<unsigned long long>(sb + FS_TILE(counters)) : nullptr;
        if (cur_counters != nullptr) {
            const uint32_t wv = __reduce_add_sync(0xffffffffu, acc_vc & 0xffffu);
            const uint32_t wc = __reduce_add_sync(0xffffffffu, acc_vc >> 16);
            const uint32_t wn = __reduce_add_sync(0xffffffffu, acc_nno);
            if (lane == 0) {
                if (wv) atomicAdd(&cur_counters[0], (unsigned long long)wv);
                if (wc) atomicAdd(&cur_counters[1], (unsigned long long)wc);
                if (wn) atomicAdd(&cur_counters[2], (unsigned long long)wn);
            }
            if (OPTIONAL_LAYERS && (P.flags & PF_HISTOGRAM)) {
#pragma unroll
                for (int bin = 0; bin < 9; ++bin) {
                    const uint32_t hv = __reduce_add_sync(0xffffffffu, (uint32_t)(acc_hist >> (7 * bin)) & 127u);
                    if (lane == 0 && hv) atomicAdd(&cur_counters[3 + bin], (unsigned long long)hv);
                }
            }
        }
        acc_vc = 0; acc_nno = 0; acc_hist = 0ull;
    };
    const bool histogram = OPTIONAL_LAYERS && (P.flags & PF_HISTOGRAM) != 0u;   // lean variant: 3 counters only

    // DEM request for the item with ordinal j of this CTA, by warp 0 (all lanes wait, lane 0 issues).  Buffer
    // j & 1 is free when every warp has arrived on empty[] for item j - 2 and its previous transaction has landed.
    auto request_dem = [&](uint32_t j, const ItemDesc &d, int padx) {
        const uint32_t b = j & 1u;
        if (j >= 2u) mbar_wait(&s.empty[b], ((j >> 1) - 1u) & 1u);
        if (fstate & (4u << b)) mbar_wait(&s.full[b], ((fstate >> b) & 1u) ^ 1u);   // its previous transaction
        if (PB200_FT_PRODUCER_WARP == 0 || lane == 0) {
            // generic-proxy reads of the buffer (ordered by the empty barrier) before the async-proxy writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&s.full[b], DEM_BOX_BYTES);
            const int gx = s.tile.dem_off_x + d.tx * FT_W - padx, gy = s.tile.dem_off_y + d.ty * FT_H - 1;
            tma_load_2d(&s.dem[b].v[0][0], &tmaps[d.tile], gx, gy, &s.full[b]);
        }
        if (PB200_FT_PRODUCER_WARP) __syncwarp();
    };

    // input registers of one row of 4 pixels; they persist across items: the first row of the next item of this
    // CTA is requested in the last row of the current one
    uint32_t w[6][2], fm4 = 0u, ld4 = 0xffffffffu, oc4 = 0x01010101u;   // no LAND class / not ocean
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k][0] = w[k][1] = 0u;
    bool row_loaded = false;

    ItemDesc item;
    item.tile = 0xffffffffu; item.tx = 0; item.ty = 0;
    if ((int)blockIdx.x < n_items) item = items[blockIdx.x];
    uint32_t k = 0;                                       // ordinal of the item in this CTA's sequence
#pragma unroll 1
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        // descriptor of the item after this one: loaded a whole item ahead
        ItemDesc next;
        next.tile = 0xffffffffu; next.tx = 0; next.ty = 0;
        if (it + (int)gridDim.x < n_items) next = items[it + gridDim.x];

        const bool fresh_tile = item.tile != cur_tile;    // else: this item's DEM tile was requested during the previous item
        if (fresh_tile) {
            // tile change (rare: a CTA sees ~7 items per HLS tile): the only CTA-wide synchronisation
            __syncthreads();                              // every warp is done with the previous tile
            flush_counters();                             // of the tile this CTA is leaving
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&tiles[item.tile]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&s.tile);
            if (tid < (int)(sizeof(TileDev) / 4)) dst[tid] = __ldg(src + tid);
            if (tid == 64) {
                const TileDev &g = tiles[item.tile];
                const double kx = 0.5 / (double)P.dxf, ky = 0.5 / (double)P.dyf;
                s.sun32[SK_SA] = (float)(kx * g.sin_az); s.sun32[SK_CA] = (float)(ky * g.cos_az);
                s.sun32[SK_SX] = (float)(kx * g.sx); s.sun32[SK_SY] = (float)(ky * g.sy); s.sun32[SK_SZ] = (float)g.sz;
                s.sun32[SK_XX] = (float)(kx * kx);
                s.sun32[SK_EA] = 1e-6f * fabsf(s.sun32[SK_SA]); s.sun32[SK_EB] = 1e-6f * fabsf(s.sun32[SK_CA]);
                if (FAST8) {
                    // "x <= 1 is implied for back slopes" needs a sun whose horizontal component points along
                    // (sin az, cos az), i.e. sin(zenith) >= 0, and a unit sun vector; any other tile (elevation outside
                    // 0..90 degrees, hand-made sun_terms) runs the exact sequence on every pixel: NaN makes v NaN
                    const double hz = g.sx * g.sin_az + g.sy * g.cos_az, n2 = g.sx * g.sx + g.sy * g.sy + g.sz * g.sz;
                    if (!(hz >= 0.0 && fabs(n2 - 1.0) < 1e-9 && fabs(g.sin_az * g.sin_az + g.cos_az * g.cos_az - 1.0) < 1e-9))
                        s.sun32[SK_XX] = __int_as_float(0x7fffffff);
                }
            }
            cur_tile = item.tile;
            ld4 = 0xffffffffu; oc4 = 0x01010101u;         // defaults of a tile without LAND / ocean raster
            __syncthreads();
        }
        const int W = s.tile.width, H = s.tile.height;
        const int x0 = item.tx * FT_W, y0 = item.ty * FT_H;
        const bool has_dem = s.tile.dem != nullptr;
        const bool has_land = s.tile.land != nullptr;
        const bool has_ocean = s.tile.ocean != nullptr;
        const bool has_counters = !OPTIONAL_LAYERS || s.tile.counters != nullptr;   // lean variant: checked by the host
        const int padx = DEM_PADX + (s.tile.dem_off_x & 3);
        const uint32_t buf = k & 1u;
        if (has_dem && (PB200_FT_PRODUCER_WARP ? warp == 0 : tid == 0)) {
            if (fresh_tile) request_dem(k, item, padx);
#if !PB200_FT_LATE_REQUEST
            if (next.tile == cur_tile) request_dem(k + 1u, next, padx);
#endif
        }
        const bool want_shad = OPTIONAL_LAYERS && s.tile.shad != nullptr;
        // the four graded layers present: one test instead of four in the row loop
        const bool all_graded = ALL_GRADED || (s.tile.diag && s.tile.wtr && s.tile.bwtr && s.tile.conf);

        bool dem_ready = !has_dem;

        const int x = x0 + 4 * lane;                      // this lane's 4 pixels
        const int nrows = min(FT_ROWS_PER_WARP, H - (y0 + rgrp * FT_ROWS_PER_WARP));   // warp-uniform
        if (x < W && nrows > 0) {
            // ---- loads: 8 bytes per band, 4 bytes per byte raster.  The input registers of a row are dead
            // after its packed and table stages: the loads of the next row (of the next item, in the last row)
            // are issued there, ahead of the shadow test, the final look-ups and the stores, which hide their latency.
            uint32_t pix = (uint32_t)(y0 + rgrp * FT_ROWS_PER_WARP) * (uint32_t)W + (uint32_t)x;   // < 2^32 checked on the host
#ifdef PB200_EXPERIMENT_FAKE_LOADS
            // issue-rate experiment (never shipped): inputs synthesised from the pixel index, no global loads
#define FT_LOAD_ROW(AT)                                                                       \
    do {                                                                                      \
        _Pragma("unroll") for (int kk = 0; kk < 6; ++kk) {                                    \
            w[kk][0] = ((AT) ^ (0x01230456u * (kk + 1))) & 0x0fff0fffu;                       \
            w[kk][1] = ((AT) ^ (0x06540321u * (kk + 3))) & 0x0fff0fffu;                       \
        }                                                                                     \
        fm4 = ((AT) ^ 0x40a0c060u) & 0xe0e0e0e0u;                                             \
        if (has_land) ld4 = (AT) | 0x80808080u;                                               \
        if (has_ocean) oc4 = 0x01010101u & ((AT) >> 4);                                       \
    } while (0)
#else
#define FT_LOAD_ROW(AT)                                                                                          \
    do {                                                                                                         \
        _Pragma("unroll") for (int kk = 0; kk < 6; ++kk) {                                                       \
            const int2 v = ldg_stream_v2(plane_at(lds_ptr<const int16_t>(sb + FS_TILE(band) + 8 * kk), (AT)));   \
            w[kk][0] = (uint32_t)v.x; w[kk][1] = (uint32_t)v.y;                                                  \
        }                                                                                                        \
        fm4 = ldg_stream_u32(plane_at(lds_ptr<const uint8_t>(sb + FS_TILE(fmask)), (AT)));                       \
        if (has_land) ld4 = ldg_stream_u32(plane_at(lds_ptr<const uint8_t>(sb + FS_TILE(land)), (AT)));          \
        if (has_ocean) oc4 = ldg_stream_u32(plane_at(lds_ptr<const uint8_t>(sb + FS_TILE(ocean)), (AT)));        \
    } while (0)
#endif
            if (!row_loaded) FT_LOAD_ROW(pix);            // not requested by the previous item (first item of a tile, raster edge)
            row_loaded = false;
#pragma unroll 1
            for (int rr = 0; rr < nrows; ++rr, pix += (uint32_t)W) {
                const int ly = rgrp * FT_ROWS_PER_WARP + rr;

                uint32_t idx[4], dgw[2], k1p[2], water_any = 0u;
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    // ================= packed stage: pixels 2p, 2p+1 =================
                    const uint32_t selb = p ? 0x4342u : 0x4140u;              // bytes -> 16-bit halves
                    const uint32_t fmh = __byte_perm(fm4, 0u, selb);          // Fmask values as halves
                    uint32_t xm;
                    if constexpr (FAST8) {
                        // every band has a fill value: (raw - fill) mod 2^16 == 0 <=> raw == fill, min-reduced in one
                        // VIADDMNMX.U16x2 per band                                                    D:2204-2207
                        xm = __byte_perm(fm4 ^ F.fmask_xor4, 0u, selb) | F.fmask_or;
#pragma unroll
                        for (int k = 5; k >= 0; --k) xm = __viaddmin_u16x2(w[k][p], F.nfill[k], xm);   // half == 0 <=> invalid
                    } else {
                        uint32_t xf[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) xf[k] = (w[k][p] ^ F.fill_xor[k]) | F.fill_or[k];   // D:2204-2207 (one LOP3)
                        const uint32_t xfm = __byte_perm(fm4 ^ F.fmask_xor4, 0u, selb) | F.fmask_or;
                        xm = __vimin3_u16x2(__vimin3_u16x2(xf[0], xf[1], xf[2]), xf[3], xf[4]);
                        xm = __vimin3_u16x2(xm, xf[5], xfm);                  // half == 0 <=> pixel invalid
                    }
                    const uint32_t nz = __vminu2(xm, 0x00010001u);            // 1 = valid
                    const uint32_t ob = __vminu2(__byte_perm(oc4, 0u, selb), 0x00010001u);   // D:5245: 0 is masked
                    const uint32_t comb = ob * 2u + nz;
                    const uint32_t B = __vmaxs2(w[0][p], 0x00010001u), G = __vmaxs2(w[1][p], 0x00010001u);   // D:2299
                    const uint32_t R = __vmaxs2(w[2][p], 0x00010001u), N = __vmaxs2(w[3][p], 0x00010001u);
                    const uint32_t S1 = __vmaxs2(w[4][p], 0x00010001u), S2 = __vmaxs2(w[5][p], 0x00010001u);
                    const uint32_t gs = __vadd2(G, S1);                                                      // D:1872
                    const uint32_t gr = __vadd2(G, R), ns = __vadd2(N, S1);                                  // D:1875-1878
                    const uint32_t nrs = __vadd2(N, R);                                                      // D:1884
                    // numerators: FAST8 needs S1 - G only (NDVI runs on the (N, R) pack)
                    const uint32_t gd = FAST8 ? __vsub2(S1, G) : __vsub2(G, S1);
                    const uint32_t nrd = FAST8 ? 0u : __vsub2(N, R);
                    const bool slow = ((gs | gr | ns | nrs) & 0x80008000u) != 0u;      // some int16 sum wrapped
                    const uint32_t T4 = __viaddmax_s16x2(N, F.m_p1nir, __vadd2(S1, F.m_p1swir1));           // < 0 <=> all below
                    const uint32_t T5 = __viaddmax_s16x2(N, F.m_p2nir, __viaddmax_s16x2(S2, F.m_p2swir2,
                                        __viaddmax_s16x2(S1, F.m_p2swir1, __vadd2(B, F.m_p2blue))));
                    const uint32_t tn = __vadd2(N, F.m_nle);                  // sign <=> nir <= 1000      (D:1239)
                    const uint32_t tb = __vadd2(N, F.m_lc);                   // sign <=> !(nir > lcmask)  (D:1354)
                    // fmask value | (nir <= 1000) << 11, and bright << 10, per half
                    const uint32_t fa = fmh | ((tn >> 4) & 0x08000800u);
#if PB200_FT_SWIZZLE_BRIGHT
                    const uint32_t brp = ((~tb >> 15) & 0x00010001u) * BIG_BRIGHT;   // per half: index bit 10 and bank bit 4
#else
                    const uint32_t brp = (~tb >> 5) & 0x04000400u;
#endif

                    uint32_t dcode[2];
#if PB200_FT_PATCH_SLOW
                    // the packed evaluation runs unconditionally (straight-line code the scheduler can interleave
                    // with the look-ups of the previous pair); a pair with a wrapped sum is re-evaluated afterwards
                    {
#else
                    if (!slow) {
#endif
                        bool p2h, p2l;
                        (void)__vibmax_s16x2(ns, gr, &p2h, &p2l);             // pred = mbsrn >= mbsrv: test 2 is its negation
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const bool hi = hh;
                            int x0w, x1w, x2w, x3w, aw = F.awesh_init;
                            if constexpr (FAST8) {
                                // per-pixel packs (lo, hi) = (gs, S1 - G) and (N, R); sign set <=> test true
                                const uint32_t pgd = __byte_perm(gs, gd, hi ? 0x7632 : 0x5410);
                                const uint32_t pnr = __byte_perm(N, R, hi ? 0x7632 : 0x5410);
                                x0w = dp2a_lo_s16_u8(pgd, F.c_wigt, 0);
                                x1w = __dp2a_lo((int)pgd, (int)F.c_p1, 0);
                                x2w = __dp2a_lo((int)pgd, (int)F.c_p2, 0);
                                x3w = __dp2a_lo((int)pnr, (int)F.c_ndvi, 0);
                                // 4*awesh: init - 4B - 10G + 6 N + 6 S1 + S2 < 0 <=> awesh > awgt
                                aw = __dp2a_lo((int)pgd, (int)F.c_aw_gd, aw);
                                aw = __dp2a_lo((int)pnr, (int)F.c_aw_nr, aw);
                                if (hi) {
                                    aw = __dp2a_hi((int)B, (int)F.c_aw_b, aw);
                                    aw = __dp2a_hi((int)S2, (int)F.c_aw_s2, aw);
                                } else {
                                    aw = __dp2a_lo((int)B, (int)F.c_aw_b, aw);
                                    aw = __dp2a_lo((int)S2, (int)F.c_aw_s2, aw);
                                }
                            } else {
                            const int n1 = hi ? sext_hi(gd) : sext_lo(gd);
                            const int q1 = hi ? (int)(gs >> 16) : (int)(gs & 0xffffu);
                            const int n2 = hi ? sext_hi(nrd) : sext_lo(nrd);
                            const int q2 = hi ? (int)(nrs >> 16) : (int)(nrs & 0xffffu);
                            // sign set <=> test true:  sa*q + p*(-sb) < 0 <=> p/q > sa/sb  (two IMADs)
                            x0w = n1 * F.nsb[RB_WIGT] + F.sa[RB_WIGT] * q1;
                            x1w = n1 * F.nsb[RB_P1_MNDWI] + F.sa[RB_P1_MNDWI] * q1;
                            x2w = n1 * F.nsb[RB_P2_MNDWI] + F.sa[RB_P2_MNDWI] * q1;
                            x3w = n2 * F.nsb[RB_P1_NDVI] + F.sa[RB_P1_NDVI] * q2;       // p/q < sa/sb
                            // 4*awesh: init - 4B - 10G + 6*mbsrn + S2 < 0 <=> awesh > awgt
                            if (hi) {
                                aw = __dp2a_hi((int)B, (int)0xFC0000FCu, aw);
                                aw = __dp2a_hi((int)G, (int)0xF60000F6u, aw);
                                aw = __dp2a_hi((int)ns, (int)0x06000006u, aw);
                                aw = __dp2a_hi((int)S2, (int)0x01000001u, aw);
                            } else {
                                aw = __dp2a_lo((int)B, (int)0xFC0000FCu, aw);
                                aw = __dp2a_lo((int)G, (int)0xF60000F6u, aw);
                                aw = __dp2a_lo((int)ns, (int)0x06000006u, aw);
                                aw = __dp2a_lo((int)S2, (int)0x01000001u, aw);
                            }
                            }
                            const uint32_t sh16 = hi ? 0u : 16u;
                            const uint32_t t2w = (hi ? p2h : p2l) ? 0u : 0x80000000u;
                            const uint32_t t4w = (uint32_t)x1w & (uint32_t)x3w & (T4 << sh16);
                            const uint32_t t5w = (uint32_t)x2w & (T5 << sh16);
                            const uint32_t cpx = hi ? (comb >> 16) : (comb & 3u);   // valid | not_ocean << 1
                            uint32_t d = __funnelshift_l(t5w, cpx, 1);
                            d = __funnelshift_l(t4w, d, 1);
                            d = __funnelshift_l((uint32_t)aw, d, 1);
                            d = __funnelshift_l(t2w, d, 1);
                            d = __funnelshift_l((uint32_t)x0w, d, 1);
                            dcode[hh] = d;
                        }
                    }
#if PB200_FT_PATCH_SLOW
                    if (slow) {
#else
                    else {
#endif
#if PB200_FT_PATCH_SLOW
                        // only the pixel whose sums wrapped (usually one of the two)
                        const uint32_t wr = (gs | gr | ns | nrs) & 0x80008000u;
                        if (wr & 0x8000u) dcode[0] = diag_pixel_slow(B, G, R, N, S1, S2, 0u, P) | ((comb & 3u) << 5);
                        if (wr >> 31) dcode[1] = diag_pixel_slow(B, G, R, N, S1, S2, 1u, P) | ((comb >> 16) << 5);
#else
                        dcode[0] = diag_pixel_slow(B, G, R, N, S1, S2, 0u, P) | ((comb & 3u) << 5);
                        dcode[1] = diag_pixel_slow(B, G, R, N, S1, S2, 1u, P) | ((comb >> 16) << 5);
#endif
                    }

                    // ================= per pixel: tables ================================
                    uint32_t dl[2];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool hi = hh;
                        const int j = 2 * p + hh;
                        dl[hh] = lds_tab32(sb + FS_OFF(diag_lut) + 4u * dcode[hh]);   // D:5227-5231, 5245, 5249
                        const uint32_t k1s = dl[hh] >> 16;                    // k1 << 8 | k1 << 2 (bank swizzle)
                        const uint32_t fi = hi ? ((fa >> 16) ^ k1s) : ((fa & 0xffffu) ^ k1s);
                        const uint32_t ev = lds_tab8(sb + FS_OFF(fk_lut) + fi);       // D:1237-1246, 1984-1991, 2081
#if PB200_FT_LAND_DP4A
                        // byte j of ld4 + table base in one IDP.4A (the FMA-side pipe; shift + mask + add are 2 ALU-pipe ops)
                        const uint32_t cat = lds_tab8(__dp4a(ld4, 1u << (8 * j), sb + FS_OFF(land_lut)));
#else
                        const uint32_t cat = lds_tab8(sb + FS_OFF(land_lut) + ((ld4 >> (8 * j)) & 255u));
#endif
                        const uint32_t bri = hi ? (brp >> 16) : (brp & 0xffffu);
                        idx[j] = ((ev & 0x7Fu) ^ bri) + (cat << 7);          // bits 7-8 are still clear: the sum is an OR
                        water_any |= ev;                                      // bit 7: the pixel holds a water class
                    }
                    dgw[p] = __byte_perm(dl[0], dl[1], 0x5410);               // two DIAG values
                    if (OPTIONAL_LAYERS) k1p[p] = __byte_perm(dl[0], dl[1], 0x4743);   // k1 of the two pixels in bytes 0 and 2
                }

                if (has_counters) acc_nno = __dp4a(oc4, 0x01010101u, acc_nno);   // D:5105; no shoreline: 1 per pixel (D:5107)
#if PB200_FT_LATE_REQUEST
                // DEM tile of the next item: requested after this thread's second row, when the other warps have
                // normally left item k - 1 (requested at the top of the item, thread 0 spent half its time waiting
                // for the slowest of them, profiles/)
                if (tid == 0 && has_dem && rr == min(1, nrows - 1) && next.tile == cur_tile) request_dem(k + 1u, next, padx);
#endif
                if (rr + 1 < nrows) {
                    FT_LOAD_ROW(pix + (uint32_t)W);
                } else if (PB200_FT_XITEM_PREFETCH && next.tile == cur_tile) {
                    // last row: first row of this warp in the next item of the CTA (same tile, same planes)
                    const int xn = next.tx * FT_W + 4 * lane, yn = next.ty * FT_H + rgrp * FT_ROWS_PER_WARP;
                    if (xn < W && yn < H) {
                        FT_LOAD_ROW((uint32_t)yn * (uint32_t)W + (uint32_t)xn);
                        row_loaded = true;
                    }
                }

                // ---- terrain shadow: only where it can change the result ----------------
                // (bit 7 of a fk_lut byte = the pixel holds a water class: only those can be masked, D:1340-1343)
                uint32_t shw[4] = {0u, 0u, 0u, 0u};                           // BIG_SHADOWED = in shadow
                if (has_dem && ((FAST8 && PB200_FT_SHADOW_ALWAYS) || want_shad || (water_any & 0x80u))) {
                    if (!dem_ready) {
                        mbar_wait(&s.full[buf], (fstate >> buf) & 1u);
                        dem_ready = true;
                    }
                    // shared address of this lane's first pixel in the middle row of the 3-row window
                    const uint32_t am = sb + FS_OFF(dem) + buf * (uint32_t)sizeof(DemHalf) +
                                        4u * (uint32_t)((ly + 1) * FT_SMW + 4 * lane + padx);
                    constexpr uint32_t RB = 4u * FT_SMW;                      // bytes per row of the DEM tile
                    float m[6], u[4], d[4];
                    if ((padx & 1) == 0) {
                        // even DEM margins (production: 50): 8-byte vectors
                        m[0] = lds_f32(am - 4);
                        const float2 m1 = lds_f32x2(am), m2 = lds_f32x2(am + 8);
                        m[1] = m1.x; m[2] = m1.y; m[3] = m2.x; m[4] = m2.y;
                        const float2 u0 = lds_f32x2(am - RB), d0 = lds_f32x2(am + RB);
                        u[0] = u0.x; u[1] = u0.y; d[0] = d0.x; d[1] = d0.y;
                        m[5] = lds_f32(am + 16);
                        const float2 u1 = lds_f32x2(am - RB + 8), d1 = lds_f32x2(am + RB + 8);
                        u[2] = u1.x; u[3] = u1.y; d[2] = d1.x; d[3] = d1.y;
                    } else {
#pragma unroll
                        for (int j = 0; j < 6; ++j) m[j] = lds_f32(am + 4 * j - 4);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { u[j] = lds_f32(am - RB + 4 * j); d[j] = lds_f32(am + RB + 4 * j); }
                    }
                    float K[SK_N];
                    {
                        const float4 k0 = lds_f32x4(sb + FS_OFF(sun32)), k1 = lds_f32x4(sb + FS_OFF(sun32) + 16);
                        K[0] = k0.x; K[1] = k0.y; K[2] = k0.z; K[3] = k0.w; K[4] = k1.x; K[5] = k1.y; K[6] = k1.z; K[7] = k1.w;
                        static_assert(SK_N == 8, "two 16-byte loads");
                    }
                    bool undecided;
                    if constexpr (FAST8) {
                        uint32_t is[4], nt[4];
                        float2 v01, v23;
#if PB200_FT_SHADOW_SCALAR
                        shadow_fast1(m[0], m[2], u[0], d[0], F, K, &is[0], &nt[0], &v01.x);
                        shadow_fast1(m[1], m[3], u[1], d[1], F, K, &is[1], &nt[1], &v01.y);
                        shadow_fast1(m[2], m[4], u[2], d[2], F, K, &is[2], &nt[2], &v23.x);
                        shadow_fast1(m[3], m[5], u[3], d[3], F, K, &is[3], &nt[3], &v23.y);
                        const uint32_t vsb = __float_as_uint((v01.x + v01.y) + (v23.x + v23.y)) + 0x00800000u;
#else
                        {
                            uint32_t i2[2], n2[2];
                            shadow_fast2(make_float2(m[0], m[1]), make_float2(m[2], m[3]), make_float2(u[0], u[1]),
                                         make_float2(d[0], d[1]), F, K, i2, n2, &v01);
                            is[0] = i2[0]; is[1] = i2[1]; nt[0] = n2[0]; nt[1] = n2[1];
                            shadow_fast2(make_float2(m[2], m[3]), make_float2(m[4], m[5]), make_float2(u[2], u[3]),
                                         make_float2(d[2], d[3]), F, K, i2, n2, &v23);
                            is[2] = i2[0]; is[3] = i2[1]; nt[2] = n2[0]; nt[3] = n2[1];
                        }
                        // bit 31 of `und`: some pixel neither certainly shadowed nor certainly lit, or a non-finite v
                        // (v >= 1; exponent 0xff + 1 carries into bit 31; NaNs from float32 arithmetic are 0x7fffffff)
                        const float2 vs2 = __fadd2_rn(v01, v23);
                        const uint32_t vsb = __float_as_uint(vs2.x + vs2.y) + 0x00800000u;
#endif
                        uint32_t und = vsb;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            und |= ~(is[j] | nt[j]);
                            shw[j] = (uint32_t)((int)is[j] >> 31) & BIG_SHADOWED;
                        }
                        undecided = (und >> 31) != 0u;
                    } else {
                        undecided = (F.fast_shadow_ok == 0u);
#pragma unroll
                        for (int j = 0; j < 4; ++j) shw[j] = shadow_fast(m[j], m[j + 2], u[j], d[j], F, K, &undecided);
                    }
                    if (undecided) {
                        // rare: redo the 4 pixels with the float64 reference sequence
#pragma unroll
                        for (int j = 0; j < 4; ++j) shw[j] = shadow_exact(m[j], m[j + 2], u[j], d[j], P, s.tile);
                    }
                }
                // ---- final look-up (D:1331-1376, 2084-2131, 1727, 1793-1835) ------------------
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = lds_tab32(sb + FS_OFF(big_lut) + 4u * (idx[j] ^ shw[j]));

                // ---- pack and store -------------------------------------------------------
                const uint32_t t01a = __byte_perm(o[0], o[1], 0x5140);
                const uint32_t t23a = __byte_perm(o[2], o[3], 0x5140);
                const uint32_t t01b = __byte_perm(o[0], o[1], 0x7362);
                const uint32_t t23b = __byte_perm(o[2], o[3], 0x7362);
                const uint32_t wtr4 = __byte_perm(t01a, t23a, 0x5410);
                const uint32_t bwtr4 = __byte_perm(t01a, t23a, 0x7632);
                const uint32_t conf4 = __byte_perm(t01b, t23b, 0x5410);
                const uint32_t flag4 = __byte_perm(t01b, t23b, 0x7632);
                if (all_graded) {
                    stg_stream_v2(plane_at(lds_ptr<uint16_t>(sb + FS_TILE(diag)), pix), dgw[0], dgw[1]);
                    stg_stream_u32(plane_at(lds_ptr<uint8_t>(sb + FS_TILE(wtr)), pix), wtr4);
                    stg_stream_u32(plane_at(lds_ptr<uint8_t>(sb + FS_TILE(bwtr)), pix), bwtr4);
                    stg_stream_u32(plane_at(lds_ptr<uint8_t>(sb + FS_TILE(conf)), pix), conf4);
                } else {
                    if (s.tile.diag) stg_stream_v2(s.tile.diag + pix, dgw[0], dgw[1]);
                    if (s.tile.wtr) stg_stream_u32(s.tile.wtr + pix, wtr4);
                    if (s.tile.bwtr) stg_stream_u32(s.tile.bwtr + pix, bwtr4);
                    if (s.tile.conf) stg_stream_u32(s.tile.conf + pix, conf4);
                }

                if (OPTIONAL_LAYERS) {
                    const uint32_t cls_lo = P.cls_lut[0] | (P.cls_lut[1] << 8) | (P.cls_lut[2] << 16) | (P.cls_lut[3] << 24);
                    const uint32_t cls_hi = P.cls_lut[4] | (P.cls_lut[5] << 8) | (P.cls_lut[6] << 16) | (P.cls_lut[7] << 24);
                    uint32_t sel1r = 0, sel2 = 0, c4 = 0, s4 = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // undo the bank swizzle of "bright" (index bit 10 also flipped bit 4 = bit 1 of c)
                        const uint32_t kb = idx[j] & 7u,
                                       c = ((idx[j] >> 3) & 15u) ^ (PB200_FT_SWIZZLE_BRIGHT ? ((idx[j] >> 9) & 2u) : 0u);
                        const uint32_t shb = shw[j] >> 9;
                        // kill_lut index: kb | shadowed<<3 | bright<<4 | cat<<5  (bright = idx bit 10, cat = idx bits 7-8)
                        const uint32_t k2 = s.kill_lut[kb | (shb << 3) | ((idx[j] >> 6) & 0x10u) | ((idx[j] >> 2) & 0x60u)];
                        sel1r |= kb << (4 * j);
                        sel2 |= k2 << (4 * j);
                        c4 |= ((k2 == 7u && !(P.flags & PF_DEFER_SNOW)) ? 255u : c) << (8 * j);    // D:2084 ('cover': after the dilations)
                        s4 |= (shb ^ 1u) << (8 * j);
                    }
                    // k1p[p]: byte 0 = k1 of pixel 2p, byte 2 = k1 of pixel 2p+1
                    const uint32_t sel1 = (k1p[0] & 7u) | ((k1p[0] >> 12) & 0x70u) | ((k1p[1] & 7u) << 8) |
                                          ((k1p[1] >> 4) & 0x7000u);
                    if (s.tile.cloud) stg_stream_u32(s.tile.cloud + pix, c4);
                    if (s.tile.wtr1) stg_stream_u32(s.tile.wtr1 + pix, __byte_perm(cls_lo, cls_hi, sel1));
                    if (s.tile.wtr1r) stg_stream_u32(s.tile.wtr1r + pix, __byte_perm(cls_lo, cls_hi, sel1r));
                    if (s.tile.wtr2) stg_stream_u32(s.tile.wtr2 + pix, __byte_perm(cls_lo, cls_hi, sel2));
                    if (s.tile.shad) stg_stream_u32(s.tile.shad + pix, s4);
                }

                // ---- counters (D:5104-5111) from the flag bytes -----------------------------
                if (has_counters) {
                    // valid in the low half, cloud-and-valid in the high half (a thread sees < 2^16 pixels per tile)
                    acc_vc += __popc(flag4 & 0x01010101u) + (__popc(flag4 & 0x02020202u) << 16);
                    if (histogram) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc_hist += 1ull << (7u * ((flag4 >> (8 * j + 4)) & 15u));
                    }
                }
            }
#undef FT_LOAD_ROW
        }

        // A warp that never looked at the DEM tile (no water pixel, raster edge) must still see the item's
        // transaction land before it arrives: the request for item k + 2 waits for all arrivals of item k, so no
        // warp can get two items ahead and arrive twice in one phase of empty[].
        if (has_dem && !dem_ready) mbar_wait(&s.full[buf], (fstate >> buf) & 1u);
        // this warp is done with the item's DEM buffer (every warp arrives for every item: phases count items)
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[buf]);
        if (has_dem) fstate = (fstate ^ (1u << buf)) | (4u << buf);   // one transaction per item of a tile with a DEM

        // histogram bins are 7 bits wide and an item adds at most 24 pixels per thread: flush every 4 items
        if (histogram && (k & 3u) == 3u) flush_counters();
        item = next;
    }
    // no transaction may be in flight when the CTA exits
    if (warp == 0) {
#pragma unroll
        for (uint32_t b = 0; b < 2u; ++b)
            if (fstate & (4u << b)) mbar_wait(&s.full[b], ((fstate >> b) & 1u) ^ 1u);
    }
    flush_counters();
}

}  // namespace pb200
