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
#define PB200_FT_LAND_DP4A 1       // land_lut address as ONE IDP.4A (byte select + base add) instead of shift, mask, add
#endif
#ifndef PB200_FT_FAST8_WRAPFIX
#define PB200_FT_FAST8_WRAPFIX 1   // FAST8: pixels whose int16 sums wrapped are patched with the IDP.2A forms + sign bookkeeping
#endif
#ifndef PB200_FT_NUMPY1
#define PB200_FT_NUMPY1 1          // the exact shadow sequence honours PF_NUMPY1 (0: A/B builds only)
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
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// arrive from ONE lane of a converged warp, chosen by elect.sync: no lane id needed (S2R + mask + compare otherwise)
__device__ __forceinline__ void mbar_arrive_elected(uint32_t bar_addr) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(bar_addr)
        : "memory");
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
// hooks of pb200_fused_row.inc for a tile descriptor + sun constants in shared memory at FastSmem's offsets from `tb`
#define FT_OUT_PTRS_SHARED(tb, pd, pw, pb, pc) \
    do { pd = lds_ptr<uint16_t>((tb) + FS_TILE(diag)); pw = lds_ptr<uint8_t>((tb) + FS_TILE(wtr)); \
         pb = lds_ptr<uint8_t>((tb) + FS_TILE(bwtr)); pc = lds_ptr<uint8_t>((tb) + FS_TILE(conf)); } while (0)
#define FT_SUN4_SHARED(tb, i) lds_f32x4((tb) + FS_OFF(sun32) + 16u * (i))
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
// FAST8: the four rational tests and 4*awesh of ONE pixel whose int16 sums wrapped (D:1872-1884: numpy adds int16 arrays
// modulo 2^16), from the same IDP.2A forms as the fast path plus sign bookkeeping - no division, no scalar re-evaluation:
//   * a denominator q16 = wrap16(G + S1) < 0: p / q16 > a / b  <=>  (-p) / (-q16) > a / b with -q16 in [2, 32768], the
//     range the bound a / b was derived for, so the test is the sign of -X instead of X (X = 0 stays false);
//   * NDVI is evaluated on the pack (N, R), i.e. with the unwrapped N + R: X(q16) = X(N + R) - sa * 65536, then the same
//     sign flip (a wrapped N + R is always negative);
//   * 4*awesh uses mbsrn = wrap16(N + S1) (D:1878): - 6 * 65536 when it wrapped; the pack (gs, S1 - G) carries the
//     WRAPPED gs, which adds 2 * 65536 when G + S1 wrapped.
// Checked over every (G, S1) and (N, R) in [1, 32767]^2 against IEEE division (pb200_fast8_sweep).
// `hi` selects the pixel of the pair; returns the sign words of tests 1, 4a, 5a, 4b (ndvi) and awesh.
struct Fast8Signs { int x0, x1, x2, x3, aw; };
__device__ __forceinline__ Fast8Signs fast8_signs_wrapped(uint32_t gs, uint32_t gd, uint32_t ns, uint32_t nrs, uint32_t N, uint32_t R,
                                                          uint32_t B, uint32_t S2, uint32_t hi, const FastParams &F) {
    const uint32_t sel = hi ? 0x7632u : 0x5410u;
    const uint32_t pgd = __byte_perm(gs, gd, sel), pnr = __byte_perm(N, R, sel);
    const int mg = hi ? ((int)gs >> 31) : ((int)(gs << 16) >> 31);          // all ones: G + S1 wrapped (q16 < 0)
    const int mn = hi ? ((int)nrs >> 31) : ((int)(nrs << 16) >> 31);        // N + R wrapped
    const int ms = hi ? ((int)ns >> 31) : ((int)(ns << 16) >> 31);          // N + S1 wrapped
    Fast8Signs r;
    r.x0 = (dp2a_lo_s16_u8(pgd, F.c_wigt, 0) ^ mg) - mg;
    r.x1 = (__dp2a_lo((int)pgd, (int)F.c_p1, 0) ^ mg) - mg;
    r.x2 = (__dp2a_lo((int)pgd, (int)F.c_p2, 0) ^ mg) - mg;
    const int x3 = __dp2a_lo((int)pnr, (int)F.c_ndvi, 0) - ((F.sa[RB_P1_NDVI] << 16) & mn);
    r.x3 = (x3 ^ mn) - mn;
    int aw = F.awesh_init;
    aw = __dp2a_lo((int)pgd, (int)F.c_aw_gd, aw);
    aw = __dp2a_lo((int)pnr, (int)F.c_aw_nr, aw);
    const uint32_t b1 = hi ? (B >> 16) : (B & 0xffffu), s1 = hi ? (S2 >> 16) : (S2 & 0xffffu);   // clipped bands: 1 .. 32767
    aw += (int)s1 - 4 * (int)b1;
    r.aw = aw - (mg & 131072) - (ms & 393216);
    return r;
}
// The 5-bit codes of BOTH pixels of a pair (FAST8), out of line and from the six clipped band registers alone - one call
// per pair with a wrapped sum, few registers live across it.  Correct for an unwrapped pixel of the pair too (every mask
// is zero there).  Returns code of the low pixel | code of the high pixel << 16.
__device__ __noinline__ uint32_t diag_pair_wrapped_fast8(uint32_t B, uint32_t G, uint32_t R, uint32_t N, uint32_t S1, uint32_t S2,
                                                         const FastParams &F) {
    const uint32_t gs = __vadd2(G, S1), gr = __vadd2(G, R), ns = __vadd2(N, S1), nrs = __vadd2(N, R), gd = __vsub2(S1, G);
    const uint32_t T4 = __viaddmax_s16x2(N, F.m_p1nir, __vadd2(S1, F.m_p1swir1));
    const uint32_t T5 = __viaddmax_s16x2(N, F.m_p2nir, __viaddmax_s16x2(S2, F.m_p2swir2,
                        __viaddmax_s16x2(S1, F.m_p2swir1, __vadd2(B, F.m_p2blue))));
    bool p2h, p2l;
    (void)__vibmax_s16x2(ns, gr, &p2h, &p2l);                 // pred = mbsrn >= mbsrv on the wrapped int16 values (D:1896)
    uint32_t out = 0u;
#pragma unroll
    for (uint32_t hi = 0u; hi < 2u; ++hi) {
        const Fast8Signs r = fast8_signs_wrapped(gs, gd, ns, nrs, N, R, B, S2, hi, F);
        const uint32_t sh16 = hi ? 0u : 16u;
        const uint32_t t2w = (hi ? p2h : p2l) ? 0u : 0x80000000u;
        const uint32_t t4w = (uint32_t)r.x1 & (uint32_t)r.x3 & (T4 << sh16);
        const uint32_t t5w = (uint32_t)r.x2 & (T5 << sh16);
        uint32_t d = __funnelshift_l(t5w, 0u, 1);
        d = __funnelshift_l(t4w, d, 1);
        d = __funnelshift_l((uint32_t)r.aw, d, 1);
        d = __funnelshift_l(t2w, d, 1);
        d = __funnelshift_l((uint32_t)r.x0, d, 1);
        out |= d << (16u * hi);
    }
    return out;
}
__device__ __forceinline__ uint32_t shadow_exact1(float l, float r, float u, float d, const DevParams &P, const SunTerms &S) {
    const float g_col = __fmul_rn(__fsub_rn(r, l), 0.5f);                     // D:4255
    const float g_row = __fmul_rn(__fsub_rn(d, u), 0.5f);
    return shadow_from_gradient(g_col, g_row, P.dxf, P.dyf, S, P.cos_thr, P.tan_thr, PB200_FT_NUMPY1 && (P.flags & PF_NUMPY1) != 0u) ? 0u : BIG_SHADOWED;
}
__device__ __noinline__ uint32_t shadow_exact(float l, float r, float u, float d, const DevParams &P,
                                              const TileDev &T) {
    SunTerms S;
    S.sx = T.sx; S.sy = T.sy; S.sz = T.sz; S.sin_az = T.sin_az; S.cos_az = T.cos_az;
    return shadow_exact1(l, r, u, d, P, S);
}
// The exact sequence for the FOUR pixels of a lane, out of line with two register arguments: the DEM window is re-read
// from the item's tile in shared memory (`am` = shared address of the lane's first pixel in the middle row, as in the row
// body), the sun terms from the tile descriptor at `sb`.  Nothing but `am`, `sb` and the result crosses the call - the row
// loop's registers stay out of the rare path (a 4 x 6-argument call sequence cost 20-40 bytes of spills IN the loop and,
// with a larger callee, 6 % of the throughput, profiles/).  Bit j of the result: pixel j is in shadow.
// bits of the four pixels from the DEM window at `am` with the given sun terms
__device__ __forceinline__ uint32_t shadow_exact4_bits(uint32_t am, const SunTerms &S, const DevParams &P) {
    constexpr uint32_t RB = 4u * FT_SMW;
    uint32_t bits = 0u;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const uint32_t a = am + 4u * (uint32_t)j;
        const float l = lds_f32(a - 4u), r = lds_f32(a + 4u), u = lds_f32(a - RB), d = lds_f32(a + RB);
        bits |= (shadow_exact1(l, r, u, d, P, S) ? 1u : 0u) << j;
    }
    return bits;
}
// the same with the tile descriptor in GLOBAL memory (dswx_fused_stream_dyn_kernel)
__device__ __noinline__ uint32_t shadow_exact4_global(uint32_t am, const TileDev *g, const DevParams &P) {
    SunTerms S;
    S.sx = __ldg(&g->sx); S.sy = __ldg(&g->sy); S.sz = __ldg(&g->sz); S.sin_az = __ldg(&g->sin_az); S.cos_az = __ldg(&g->cos_az);
    return shadow_exact4_bits(am, S, P);
}
__device__ __noinline__ uint32_t shadow_exact4(uint32_t am, uint32_t sb, const DevParams &P) {
    SunTerms S;
    {
        const uint32_t at = sb + (uint32_t)(offsetof(FastSmem, tile) + offsetof(TileDev, sx));
        double v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[i]) : "r"(at + 8u * (uint32_t)i));
        S.sx = v[0]; S.sy = v[1]; S.sz = v[2]; S.sin_az = v[3]; S.cos_az = v[4];
    }
    return shadow_exact4_bits(am, S, P);
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

                // hooks of the shared row body (pb200_fused_row.inc).  Mid-row, when the row's input registers are dead:
                // request the DEM tile of the next item (after this thread's second row, when the other warps have
                // normally left item k - 1: requested at the top of the item, thread 0 spent half its time waiting for the
                // slowest of them), then issue the loads of the next row - of the warp's first row in the next item of
                // the CTA (same tile, same planes) when this is the last one
#define FT_DEM_WAIT() mbar_wait(&s.full[buf], (fstate >> buf) & 1u)
#define FT_ROW_MIDPOINT() do { \
                    if (PB200_FT_LATE_REQUEST && tid == 0 && has_dem && rr == min(1, nrows - 1) && next.tile == cur_tile) request_dem(k + 1u, next, padx); \
                if (rr + 1 < nrows) { \
                    FT_LOAD_ROW(pix + (uint32_t)W); \
                } else if (PB200_FT_XITEM_PREFETCH && next.tile == cur_tile) { \
                    const int xn = next.tx * FT_W + 4 * lane, yn = next.ty * FT_H + rgrp * FT_ROWS_PER_WARP; \
                    if (xn < W && yn < H) { \
                        FT_LOAD_ROW((uint32_t)yn * (uint32_t)W + (uint32_t)xn); \
                        row_loaded = true; \
                    } \
                } \
                } while (0)
#define FT_PADX() padx
#define FT_OUT_PTRS(pd, pw, pb, pc) FT_OUT_PTRS_SHARED(sb, pd, pw, pb, pc)
#define FT_SUN4(i) FT_SUN4_SHARED(sb, i)
#define FT_EXACT4(am) shadow_exact4(am, sb, P)
#define FT_DEM_BASE (sb + FS_OFF(dem) + buf * (uint32_t)sizeof(DemHalf))
#include "pb200_fused_row.inc"
#undef FT_DEM_BASE
#undef FT_EXACT4
#undef FT_SUN4
#undef FT_OUT_PTRS
#undef FT_PADX
#undef FT_ROW_MIDPOINT
#undef FT_DEM_WAIT
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
