// pb200_api.cu - C ABI of libproteus_b200.so (see include/proteus_b200.h).
//
// Host side of the B200 DSWx-HLS classification path: parameter derivation
// (exact rational bounds, integer thresholds, look-up tables), tile
// descriptors + TMA tensor maps, kernel launches, and the strip-pipelined
// host-buffer entry point.  No CPU implementation of the pixel math lives
// here: every compute entry point launches a CUDA kernel or fails.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/proteus_b200.h"
#include "pb200_kernels.cuh"
#include "pb200_fused.cuh"
#include "pb200_stream.cuh"
#include "pb200_cover.cuh"
#include "pb200_landcover.cuh"
#include "pb200_comm.cuh"
#include "pb200_hillshade.cuh"
#include "pb200_sweep.cuh"

using namespace pb200;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
static int fail_cuda(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " - " + cudaGetErrorString(e);
    return (int)e > 0 ? (int)e : 999;
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call); \
    } while (0)

extern "C" int pb200_version(void) { return PB200_ABI_VERSION; }
extern "C" const char *pb200_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------
// exact rational bound for  float64(n)/float64(d) {>,<} t,  n, d int16
// ---------------------------------------------------------------------------
// RN(n/d) > t  <=>  n/d > M, M = midpoint(t, succ(t))   (n/d is never a
// midpoint: it would need a 54-bit significand, |d| < 2^16).  Among all p/q
// with |p| <= 32768, 1 <= q <= 32768 the smallest one above M is a/b, so
// n/d > M <=> p/q >= a/b <=> p*b >= a*q.  Mirror image for "<".
typedef __int128 i128;

// value = mant * 2^exp2, mant odd or zero
static void decompose(double v, long long *mant, int *exp2) {
    if (v == 0.0) { *mant = 0; *exp2 = 0; return; }
    int e;
    const double f = std::frexp(v, &e);          // v = f * 2^e, 0.5 <= |f| < 1
    *mant = (long long)std::ldexp(f, 53);        // exact: 53-bit integer
    *exp2 = e - 53;
}

// floor(M * q) for M = (m1*2^e1 + m2*2^e2) / 2 and 1 <= q <= 32768
static long long floor_mid_times(long long m1, int e1, long long m2, int e2, long long q) {
    const int k = std::min(e1, e2);
    const i128 sum = ((i128)m1 << (e1 - k)) + ((i128)m2 << (e2 - k));    // shifts <= 2 for neighbours / zero
    const i128 prod = sum * (i128)q;                                     // |prod| < 2^72
    const int sh = -(k - 1);                                             // M*q = prod / 2^sh
    if (sh <= 0) return (long long)(prod << (-sh));
    if (sh >= 120) return prod < 0 ? -1 : 0;
    return (long long)(prod >> sh);                                      // arithmetic shift = floor
}

extern "C" int pb200_ratio_bound(double t, int is_less, int32_t *a_out, int32_t *b_out) {
    if (!a_out || !b_out) return fail(PB200_E_INVALID_ARG, "pb200_ratio_bound: null output");
    const int never_a = is_less ? -1 : 1, always_a = is_less ? 1 : -1;
    if (std::isnan(t)) { *a_out = never_a; *b_out = 0; return 0; }
    if (t >= 1048576.0) { *a_out = is_less ? always_a : never_a; *b_out = 0; return 0; }
    if (t <= -1048576.0) { *a_out = is_less ? never_a : always_a; *b_out = 0; return 0; }
    const double nb = is_less ? std::nextafter(t, -INFINITY) : std::nextafter(t, INFINITY);
    long long m1, m2;
    int e1, e2;
    decompose(t, &m1, &e1);
    decompose(nb, &m2, &e2);
    if (m1 == 0) e1 = e2;
    if (m2 == 0) e2 = e1;
    long long best_a = 0, best_b = 0;
    for (long long q = 1; q <= 32768; ++q) {
        long long a;
        if (!is_less) {
            a = floor_mid_times(m1, e1, m2, e2, q) + 1;       // smallest integer a with a/q > M
            if (a > 32768) continue;
            if (a < -32768) a = -32768;
            if (best_b == 0 || a * best_b < best_a * q) { best_a = a; best_b = q; }
        } else {
            // largest integer a with a/q < M:  a = ceil(M q) - 1 = -floor(-M q) - 1
            a = -floor_mid_times(-m1, e1, -m2, e2, q) - 1;
            if (a < -32768) continue;
            if (a > 32768) a = 32768;
            if (best_b == 0 || a * best_b > best_a * q) { best_a = a; best_b = q; }
        }
    }
    if (best_b == 0) { *a_out = never_a; *b_out = 0; return 0; }
    *a_out = (int32_t)best_a;
    *b_out = (int32_t)best_b;
    return 0;
}

// Strict companion of pb200_ratio_bound for the fast kernel:
//   is_less = 0:  RN(n/d) > t  <=>  p/q > a/b   with a/b = LARGEST fraction <= midpoint M
//   is_less = 1:  RN(n/d) < t  <=>  p/q < a/b   with a/b = SMALLEST fraction >= midpoint M-
// (p/q is never equal to a midpoint, so "<= M" / ">= M" only matter when M is itself such a
// fraction, which cannot happen; the test then needs no "- 1" term.)  b == 0: always / never as in
// pb200_ratio_bound, with a = -1 (">" always), +1 (">" never), +1 ("<" always), -1 ("<" never).
static int ratio_bound_strict(double t, int is_less, int32_t *a_out, int32_t *b_out) {
    const int never_a = is_less ? -1 : 1, always_a = is_less ? 1 : -1;
    if (std::isnan(t)) { *a_out = never_a; *b_out = 0; return 0; }
    if (t >= 1048576.0) { *a_out = is_less ? always_a : never_a; *b_out = 0; return 0; }
    if (t <= -1048576.0) { *a_out = is_less ? never_a : always_a; *b_out = 0; return 0; }
    const double nb = is_less ? std::nextafter(t, -INFINITY) : std::nextafter(t, INFINITY);
    long long m1, m2;
    int e1, e2;
    decompose(t, &m1, &e1);
    decompose(nb, &m2, &e2);
    if (m1 == 0) e1 = e2;
    if (m2 == 0) e2 = e1;
    long long best_a = 0, best_b = 0;
    for (long long q = 1; q <= 32768; ++q) {
        long long a;
        if (!is_less) {
            a = floor_mid_times(m1, e1, m2, e2, q);                // largest integer a with a/q <= M
            if (a < -32768) continue;
            if (a > 32768) a = 32768;
            if (best_b == 0 || a * best_b > best_a * q) { best_a = a; best_b = q; }
        } else {
            a = -floor_mid_times(-m1, e1, -m2, e2, q);             // smallest integer a with a/q >= M-
            if (a > 32768) continue;
            if (a < -32768) a = -32768;
            if (best_b == 0 || a * best_b < best_a * q) { best_a = a; best_b = q; }
        }
    }
    if (best_b == 0) { *a_out = always_a; *b_out = 0; return 0; }   // no fraction on that side: every p/q passes
    *a_out = (int32_t)best_a;
    *b_out = (int32_t)best_b;
    return 0;
}

// ---------------------------------------------------------------------------
// angle thresholds in the cosine / tangent domain (libm flavour)
// ---------------------------------------------------------------------------
static inline long long dkey(double x) {            // order-preserving double -> int64
    long long k;
    std::memcpy(&k, &x, 8);
    return k < 0 ? (long long)0x8000000000000000ULL - k : k;
}
static inline double dunkey(long long k) {
    long long b = k < 0 ? (long long)0x8000000000000000ULL - k : k;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
static const double RAD2DEG = 180.0 / 3.14159265358979323846;

// smallest x in [-1, 1] with degrees(acos(x)) <= max_inc; 2.0 if none
static double derive_cos_threshold(double max_inc) {
    auto pred = [&](double x) { return std::acos(x) * RAD2DEG <= max_inc; };
    if (!pred(1.0)) return 2.0;
    if (pred(-1.0)) return -1.0;
    long long lo = dkey(-1.0), hi = dkey(1.0);      // pred(lo) false, pred(hi) true
    while (lo + 1 < hi) {
        const long long mid = (lo >> 1) + (hi >> 1) + (lo & hi & 1);    // no overflow
        if (pred(dunkey(mid))) hi = mid; else lo = mid;
    }
    return dunkey(hi);
}
// largest s with degrees(atan(s)) <= min_slope; +inf if all, NaN if none
static double derive_tan_threshold(double min_slope) {
    auto pred = [&](double s) { return std::atan(s) * RAD2DEG <= min_slope; };
    if (pred(INFINITY)) return INFINITY;
    if (!pred(-INFINITY)) return std::numeric_limits<double>::quiet_NaN();
    long long lo = dkey(-INFINITY), hi = dkey(INFINITY);   // pred(lo) true, pred(hi) false
    while (lo + 1 < hi) {
        const long long mid = (lo >> 1) + (hi >> 1) + (lo & hi & 1);    // no overflow
        if (pred(dunkey(mid))) lo = mid; else hi = mid;
    }
    return dunkey(lo);
}

// float32 flavours (numpy 1.x promotion: arccos / arctan / degrees of float32 arrays, degrees = x * (180 / pi) in float32)
static uint32_t fkey(float x) { uint32_t b; std::memcpy(&b, &x, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
static float funkey(uint32_t k) { const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float x; std::memcpy(&x, &b, 4); return x; }
static const float RAD2DEG_F = (float)RAD2DEG;
static double derive_cos_threshold_f32(double max_inc) {
    const float lim = (float)max_inc;
    auto pred = [&](float x) { return std::acos(x) * RAD2DEG_F <= lim; };
    if (!pred(1.0f)) return 2.0;
    if (pred(-1.0f)) return -1.0;
    uint32_t lo = fkey(-1.0f), hi = fkey(1.0f);
    while (lo + 1 < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (pred(funkey(mid))) hi = mid; else lo = mid;
    }
    return (double)funkey(hi);
}
static double derive_tan_threshold_f32(double min_slope) {
    const float lim = (float)min_slope;
    auto pred = [&](float s) { return std::atan(s) * RAD2DEG_F <= lim; };
    if (pred(INFINITY)) return INFINITY;
    if (!pred(-INFINITY)) return std::numeric_limits<double>::quiet_NaN();
    uint32_t lo = fkey(-INFINITY), hi = fkey(INFINITY);
    while (lo + 1 < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (pred(funkey(mid))) lo = mid; else hi = mid;
    }
    return (double)funkey(lo);
}

extern "C" int pb200_angle_thresholds(const pb200_params *p, double *cos_inc, double *tan_slope) {
    if (!p || !cos_inc || !tan_slope) return fail(PB200_E_INVALID_ARG, "pb200_angle_thresholds: null argument");
    // NaN in cos_inc_threshold = "derive both"; otherwise both are taken as given
    // (NaN is a legitimate tan_slope_threshold: "no pixel is a back slope").
    if (std::isnan(p->cos_inc_threshold)) {
        *cos_inc = p->numpy1_promotion ? derive_cos_threshold_f32(p->max_sun_local_inc_angle) : derive_cos_threshold(p->max_sun_local_inc_angle);
        *tan_slope = p->numpy1_promotion ? derive_tan_threshold_f32(p->min_slope_angle) : derive_tan_threshold(p->min_slope_angle);
    } else {
        *cos_inc = p->cos_inc_threshold;
        *tan_slope = p->tan_slope_threshold;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// parameter derivation
// ---------------------------------------------------------------------------
extern "C" int pb200_params_default(pb200_params *p) {
    if (!p) return fail(PB200_E_INVALID_ARG, "pb200_params_default: null");
    std::memset(p, 0, sizeof(*p));
    p->th.wigt = 0.124; p->th.awgt = 0.0;
    p->th.pswt_1_mndwi = -0.44; p->th.pswt_1_nir = 1500; p->th.pswt_1_swir1 = 900; p->th.pswt_1_ndvi = 0.7;
    p->th.pswt_2_mndwi = -0.5; p->th.pswt_2_blue = 1000; p->th.pswt_2_nir = 2500;
    p->th.pswt_2_swir1 = 3000; p->th.pswt_2_swir2 = 1000;
    p->th.lcmask_nir = 1200;
    for (int k = 0; k < 6; ++k) p->band_fill[k] = -9999;
    p->fmask_fill = 255;
    p->adjacent_mode = PB200_ADJ_MASK;
    p->apply_aerosol_class_remapping = 1;
    const int l3[] = {224, 160, 96}, l5[] = {224, 192, 160, 128, 96};
    for (int v : l3) p->aerosol_class_bits[v] |= (1u << 0) | (1u << 2);
    for (int v : l5) p->aerosol_class_bits[v] |= (1u << 3) | (1u << 4);
    p->min_slope_angle = -5; p->max_sun_local_inc_angle = 40;
    p->cos_inc_threshold = p->tan_slope_threshold = std::numeric_limits<double>::quiet_NaN();
    p->pixel_spacing_x = p->pixel_spacing_y = 30;
    p->collapse_wtr_classes = 1;
    p->class_histogram = 0;
    return 0;
}

static int32_t clamp_i32(double v, double lo, double hi) { return (int32_t)std::max(lo, std::min(hi, v)); }
// integer x:  x < t <=> x < ceil(t);  NaN -> never
static int32_t lt_threshold(double t) { return std::isnan(t) ? -40000 : clamp_i32(std::ceil(t), -40000, 40000); }
// integer x:  x > t <=> x > floor(t); NaN -> never
static int32_t gt_threshold(double t) { return std::isnan(t) ? 40000 : clamp_i32(std::floor(t), -40000, 40000); }

static int derive_thresholds(const pb200_thresholds &th, DevParams *D) {
    int rc;
    if ((rc = pb200_ratio_bound(th.wigt, 0, &D->r_a[RB_WIGT], &D->r_b[RB_WIGT]))) return rc;
    if ((rc = pb200_ratio_bound(th.pswt_1_mndwi, 0, &D->r_a[RB_P1_MNDWI], &D->r_b[RB_P1_MNDWI]))) return rc;
    if ((rc = pb200_ratio_bound(th.pswt_2_mndwi, 0, &D->r_a[RB_P2_MNDWI], &D->r_b[RB_P2_MNDWI]))) return rc;
    if ((rc = pb200_ratio_bound(th.pswt_1_ndvi, 1, &D->r_a[RB_P1_NDVI], &D->r_b[RB_P1_NDVI]))) return rc;
    // awesh > awgt <=> 4*awesh > 4*awgt <=> (int)4*awesh > floor(4*awgt)
    D->awesh4_thr = std::isnan(th.awgt) ? (1 << 30)
                                        : clamp_i32(std::floor(4.0 * th.awgt), -(double)(1 << 30), (double)(1 << 30));
    D->p1_swir1 = lt_threshold(th.pswt_1_swir1);
    D->p1_nir = lt_threshold(th.pswt_1_nir);
    D->p2_blue = lt_threshold(th.pswt_2_blue);
    D->p2_swir1 = lt_threshold(th.pswt_2_swir1);
    D->p2_swir2 = lt_threshold(th.pswt_2_swir2);
    D->p2_nir = lt_threshold(th.pswt_2_nir);
    D->lc_nir = gt_threshold(th.lcmask_nir);
    return 0;
}

static int derive_params(const pb200_params *p, DevParams *D, bool fused) {
    std::memset(D, 0, sizeof(*D));
    if (p->adjacent_mode != PB200_ADJ_MASK && p->adjacent_mode != PB200_ADJ_IGNORE &&
        p->adjacent_mode != PB200_ADJ_COVER)
        return fail(PB200_E_BAD_MODE, "ERROR mask adjacent to cloud/cloud-shadow mode: %d", p->adjacent_mode);
    if (fused && p->adjacent_mode == PB200_ADJ_COVER)
        return fail(PB200_E_UNSUPPORTED,
                    "mask_adjacent_to_cloud_mode 'cover' (masked dilation, D:2055-2078) is not part of the fused pass");
    int rc = derive_thresholds(p->th, D);
    if (rc) return rc;
    for (int k = 0; k < 6; ++k) D->band_fill[k] = p->band_fill[k];
    D->fmask_fill = p->fmask_fill;
    D->flags = (p->apply_aerosol_class_remapping ? PF_AEROSOL : 0u) | (p->collapse_wtr_classes ? PF_COLLAPSE : 0u) |
               (p->class_histogram ? PF_HISTOGRAM : 0u) | (p->defer_snow ? PF_DEFER_SNOW : 0u) |
               (p->numpy1_promotion ? PF_NUMPY1 : 0u);
    D->dxf = (float)p->pixel_spacing_x;
    D->dyf = -std::fabs((float)p->pixel_spacing_y);
    double c, t;
    pb200_angle_thresholds(p, &c, &t);
    D->cos_thr = c;
    D->tan_thr = t;
    const int mode = p->adjacent_mode == PB200_ADJ_MASK ? 0 : 1;
    for (uint32_t v = 0; v < 256; ++v) {
        uint32_t e = preliminary_cloud(v, mode);
        if ((v & 16u) && !p->defer_snow) e |= 8u;
        e |= (uint32_t)(p->aerosol_class_bits[v] & 0x1Du) << 4;
        if ((int)v == p->fmask_fill) e |= 0x8000u;
        D->fmask_lut[v] = (uint16_t)e;
    }
    for (uint32_t d = 0; d < 32; ++d) {
        const uint32_t rep = (d & 1u) + ((d >> 1) & 1u) * 10u + ((d >> 2) & 1u) * 100u + ((d >> 3) & 1u) * 1000u +
                             ((d >> 4) & 1u) * 10000u;
        D->diag_lut[d] = rep | (interpreted_class(d) << 16);
    }
    const bool collapse = p->collapse_wtr_classes != 0;
    for (uint32_t k = 0; k < 8; ++k) {
        const uint32_t w2 = k < 5u ? k : (k == 5u ? 255u : 248u + k);
        D->cls_lut[k] = (uint8_t)(collapse ? collapse_class(w2) : w2);
        for (uint32_t c4 = 0; c4 < 16; ++c4) {
            uint32_t wtr = cloud_masking(w2, c4);
            const uint32_t bw = binary_water(wtr);
            const uint32_t cf = confidence(w2, c4);
            // D:2084 (cloud[wtr2 == 255] = 255) comes AFTER the dilations of the 'cover' branch (D:2057-2078), which test
            // cloud == 0 on the preliminary values: with the snow bit deferred the fill pixels keep them too
            const uint32_t cl = (w2 == 255u && !p->defer_snow) ? 255u : c4;
            if (collapse) wtr = collapse_class(wtr);
            D->out_lut[k * 16 + c4] = wtr | (bw << 8) | (cf << 16) | (cl << 24);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// tables and packed constants of the fast kernel (pb200_fused.cuh)
// ---------------------------------------------------------------------------
static uint32_t land_category(uint32_t land) {          // D:1133-1207
    if (land == 200u) return LC_WATER;
    if (land == 201u || land < 100u) return LC_EVERGREEN_OR_LOW;
    if (land < 200u) return LC_HIGH;
    return LC_NONE;
}
static uint32_t kill_class(uint32_t kb, bool shadowed, bool bright, uint32_t cat) {   // D:1331-1376
    const bool water = kb >= 1u && kb <= 4u, psw = kb == 3u || kb == 4u;
    const bool kill = (shadowed && water && cat != LC_WATER) || (cat == LC_EVERGREEN_OR_LOW && bright && psw) ||
                      (cat == LC_HIGH && water);
    return kill ? 0u : kb;
}

static void build_fused_tables(const pb200_params *p, const DevParams &D, FusedTables *T) {
    std::memset(T, 0, sizeof(*T));
    for (uint32_t idx = 0; idx < 128; ++idx) {
        const uint32_t d = idx & 31u;
        const bool valid = (idx >> 5) & 1u, not_ocean = (idx >> 6) & 1u;
        const uint32_t rep = valid ? (D.diag_lut[d] & 0xffffu) : 65535u;                 // D:5227, D:5231
        const uint32_t k1 = !valid ? 7u : (!not_ocean ? 6u : (D.diag_lut[d] >> 16));     // D:5229, 5245, 5249
        // k1 at bits 8-10 of the fk_lut index, and again at bits 2-4: XORed into the index it spreads the classes
        // of neighbouring pixels over the shared-memory banks (FK_SWIZZLE in pb200_fused.cuh)
        T->diag_lut[idx] = rep | (((k1 << 8) | (k1 << 2)) << 16);
    }
    for (uint32_t v = 0; v < 256; ++v) T->land_lut[v] = (uint8_t)land_category(v);
    const bool aerosol_on = p->apply_aerosol_class_remapping != 0;
    for (uint32_t idx = 0; idx < 4096; ++idx) {
        const uint32_t fm = idx & 255u, k1 = (idx >> 8) & 7u, nle = idx >> 11;
        const uint32_t fe = D.fmask_lut[fm];
        const uint32_t cprelim = fe & 7u, snow = (fe >> 3) & 1u, aero = (fe >> 4) & 0x1Fu;
        const bool remap = aerosol_on && nle && k1 <= 4u && ((aero >> k1) & 1u);        // D:1237-1239
        const uint32_t kb = remap ? 1u : k1;
        const uint32_t c = cprelim | (remap ? 8u : 0u) | (snow << 1);                    // D:1246, D:2081
        const uint32_t water = (kb >= 1u && kb <= 4u) ? 0x80u : 0u;                       // can be masked by terrain shadow
        T->fk_lut[idx ^ (k1 << 2)] = (uint8_t)(kb | (c << 3) | water);
    }
    for (uint32_t idx = 0; idx < 128; ++idx)
        T->kill_lut[idx] = (uint8_t)kill_class(idx & 7u, (idx >> 3) & 1u, (idx >> 4) & 1u, (idx >> 5) & 3u);
    for (uint32_t idx = 0; idx < 2048; ++idx) {
        const uint32_t kb = idx & 7u, c = (idx >> 3) & 15u, cat = (idx >> 7) & 3u;
        const bool shadowed = (idx >> 9) & 1u, bright = (idx >> 10) & 1u;
        const uint32_t k2 = kill_class(kb, shadowed, bright, cat);
        const uint32_t k2_lit = kill_class(kb, false, bright, cat), k2_dark = kill_class(kb, true, bright, cat);
        const uint32_t out = D.out_lut[k2 * 16 + c] & 0x00ffffffu;                       // WTR | BWTR | CONF
        // flags: valid / valid-and-preliminary-cloud (D:5104-5111), shadow sensitivity, histogram bin
        const bool valid = kb <= 4u;
        const bool cv = valid && (c & 5u) != 0u;
        const uint32_t w2 = k2 < 5u ? k2 : (k2 == 5u ? 255u : 248u + k2);
        const uint32_t wtr = cloud_masking(w2, c);                                       // uncollapsed WTR class
        const uint32_t bin = wtr < 5u ? wtr : wtr - 247u;                                // 252..255 -> 5..8
        uint32_t flags = (valid ? 1u : 0u) | (cv ? 2u : 0u) | (k2_lit != k2_dark ? 4u : 0u) | (bin << 4);
        // physical slot: "shadowed" and "bright" also flip bits 3 and 4 (the kernel XORs 0x208 / 0x410 in)
        T->big_lut[idx ^ (shadowed ? (BIG_SHADOWED & 31u) : 0u) ^ (bright ? (BIG_BRIGHT & 31u) : 0u)] = out | (flags << 24);
    }
}

static uint32_t pack_neg(int v) { const uint32_t h = (uint32_t)(-v) & 0xffffu; return h | (h << 16); }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void build_fast_params(const pb200_params *p, const DevParams &D, FastParams *F) {
    std::memset(F, 0, sizeof(*F));
    for (int k = 0; k < 6; ++k) {
        const int f = p->band_fill[k];
        if (f == PB200_NO_FILL || f < -32768 || f > 32767) {
            F->fill_xor[k] = 0u; F->fill_or[k] = 0xffffffffu;          // never equal
            F->any_nofill = 1u;
        } else {
            const uint32_t h = (uint32_t)f & 0xffffu;
            F->fill_xor[k] = h | (h << 16); F->fill_or[k] = 0u;
        }
    }
    // clipped reflectances lie in [1, 32767]: "x < t" only needs t in [1, 32768]
    F->m_p1swir1 = pack_neg(clampi(D.p1_swir1, 1, 32768));
    F->m_p1nir = pack_neg(clampi(D.p1_nir, 1, 32768));
    F->m_p2blue = pack_neg(clampi(D.p2_blue, 1, 32768));
    F->m_p2swir1 = pack_neg(clampi(D.p2_swir1, 1, 32768));
    F->m_p2swir2 = pack_neg(clampi(D.p2_swir2, 1, 32768));
    F->m_p2nir = pack_neg(clampi(D.p2_nir, 1, 32768));
    F->m_nle = pack_neg(1001);                                          // nir <= 1000.0  (D:46, D:1239)
    F->m_lc = pack_neg(clampi(D.lc_nir, 0, 32767) + 1);                 // nir > lcmask_nir
    // |4*awesh| < 2^20: clamp so that init - 4*awesh cannot overflow
    F->awesh_init = clampi(D.awesh4_thr, -(1 << 24), 1 << 24);
    for (int i = 0; i < 4; ++i) { F->ra[i] = D.r_a[i]; F->rb[i] = D.r_b[i]; }
    {
        const double thr[4] = {p->th.wigt, p->th.pswt_1_mndwi, p->th.pswt_2_mndwi, p->th.pswt_1_ndvi};
        for (int i = 0; i < 4; ++i) {
            const int less = (i == RB_P1_NDVI);
            int32_t a, b;
            ratio_bound_strict(thr[i], less, &a, &b);
            // ">": sign(a*q - p*b);  "<": sign(p*b - a*q).  With b == 0 the value is +-a*q (q >= 1).
            F->sa[i] = less ? -a : a;
            F->nsb[i] = less ? b : -b;
        }
    }
    if (p->fmask_fill >= 0 && p->fmask_fill <= 255) {
        F->fmask_xor4 = (uint32_t)p->fmask_fill * 0x01010101u; F->fmask_or = 0u;
    } else {
        F->fmask_xor4 = 0u; F->fmask_or = 0x00010001u;                  // never equal
    }
    F->kx = 0.5f / D.dxf;
    F->ky = 0.5f / D.dyf;
    const bool ok = std::isfinite(D.cos_thr) && std::fabs(D.cos_thr) <= 1.0 && std::isfinite(D.tan_thr) &&
                    std::fabs(D.tan_thr) < 1e6 && std::isfinite(F->kx) && std::isfinite(F->ky) &&
                    std::fabs(D.dxf) == std::fabs(D.dyf);     // square pixels: the shortcut shares kx^2 = ky^2
    F->fast_shadow_ok = ok ? 1u : 0u;
    F->tan32 = ok ? (float)D.tan_thr : 0.0f;
    F->e0 = 1e-6f * std::fabs(F->tan32) + 1e-30f;
    F->cc32 = ok ? (float)(D.cos_thr * std::fabs(D.cos_thr)) : 0.0f;

    // ---- FAST8: the kernel variant for parameters shaped like the defaults (pb200_fused.cuh) ------------------------
    bool f8 = ok && !F->any_nofill;
    for (int k = 0; k < 6; ++k) {
        const uint32_t h = (uint32_t)(-p->band_fill[k]) & 0xffffu;
        F->nfill[k] = h | (h << 16);
    }
    // rational tests as one dp2a on a per-pixel pack.  x = sa*q + nsb*n < 0 <=> test true (strict forms above).
    //   MNDWI: q = gs = G + S1, n = G - S1 = -dg  ->  sa * gs + (-nsb) * dg        on the pack (gs, dg)
    //   NDVI:  q = N + R, n = N - R               ->  (nsb + sa) * N + (sa - nsb) * R   on the pack (N, R)
    auto bytes = [](int b0, int b1) { return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8); };
    auto fits_s8 = [](int v) { return v >= -128 && v <= 127; };
    auto fits_u8 = [](int v) { return v >= 0 && v <= 255; };
    {
        const int q0 = F->sa[RB_WIGT], d0 = -F->nsb[RB_WIGT];
        f8 = f8 && fits_u8(q0) && fits_u8(d0);
        F->c_wigt = bytes(q0, d0);
        const int q1 = F->sa[RB_P1_MNDWI], d1 = -F->nsb[RB_P1_MNDWI];
        f8 = f8 && fits_s8(q1) && fits_s8(d1);
        F->c_p1 = bytes(q1, d1);
        const int q2 = F->sa[RB_P2_MNDWI], d2 = -F->nsb[RB_P2_MNDWI];
        f8 = f8 && fits_s8(q2) && fits_s8(d2);
        F->c_p2 = bytes(q2, d2);
        const int cn = F->nsb[RB_P1_NDVI] + F->sa[RB_P1_NDVI], cr = F->sa[RB_P1_NDVI] - F->nsb[RB_P1_NDVI];
        f8 = f8 && fits_s8(cn) && fits_s8(cr);
        F->c_ndvi = bytes(cn, cr);
        F->c_aw_gd = bytes(-2, 8);      // -2 gs + 8 dg = -10 G + 6 S1
        F->c_aw_nr = bytes(6, 0);       // 6 N
        F->c_aw_b = 0xFC0000FCu;        // -4 B
        F->c_aw_s2 = 0x01000001u;       // + S2
    }
    // shadow shortcut on sign bits: needs cos_thr > 0 (dot^2 form) and tan_thr < 0 with a margin ("x <= 1" is implied)
    f8 = f8 && D.cos_thr > 0.01 && D.cos_thr <= 1.0 && D.tan_thr <= -0.005;
    F->sh_c7 = 7.08e-7f;
    {
        const double cc = D.cos_thr * D.cos_thr;
        F->ncc_hi = -(float)(cc + 4e-6);
        F->ncc_lo = -(float)(cc - 4e-6);
    }
    F->fast8 = f8 ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Arena {                  // grow-only device scratch, reused by every pb200_classify_host call
    unsigned char *base = nullptr;
    size_t cap = 0, used = 0;
    void *take(size_t bytes) {
        const size_t a = (used + 255) & ~(size_t)255;
        if (a + bytes > cap) return nullptr;
        used = a + bytes;
        return base + a;
    }
};

struct HostPipe {               // device mirror of one host tile (pb200_classify_host)
    size_t cap_px = 0, cap_dem = 0;
    int16_t *band[6] = {};
    uint8_t *fmask = nullptr, *land = nullptr, *ocean = nullptr;
    float *dem = nullptr;
    uint16_t *diag = nullptr;
    uint8_t *u8out[8] = {};
    unsigned long long *counters = nullptr;
    FusedTables *tables = nullptr;      // uploaded once per pb200_classify_host call
    Arena arena;                        // strip descriptors, tensor maps, item list
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    // ancillary rasters left on the device by the last successful call (PB200_HOST_REUSE_ANCILLARY)
    bool pending = false;               // enqueued with PB200_HOST_ASYNC and not yet waited for
    pb200_tile pending_geom = {};       // what pb200_host_wait records as resident once the work has completed
    bool anc_valid = false, anc_dem = false, anc_land = false, anc_ocean = false;
    int anc_h = 0, anc_w = 0, anc_dem_rows = 0, anc_dem_pitch = 0, anc_dem_off_y = 0, anc_dem_off_x = 0;
};

struct pb200_ctx {
    int device = 0;
    int sm_count = 0;
    bool fast_ready = false;
    int fast_ctas_per_sm = 1;        // lean variant
    int fast_ctas_per_sm_full = 1;   // variant with the optional layers
    EncodeTiledFn encode = nullptr;
    cudaMemPool_t pool = nullptr;    // private stream-ordered pool: plan descriptors, tensor maps, item lists
    cudaStream_t s_plan = nullptr;   // private non-blocking stream: uploads of pb200_plan_create
    Comm comm;                       // NCCL communicator of a row-stripped raster (pb200_comm_init), else empty
    HostPipe pipe[2];                // two host tiles in flight (PB200_HOST_SLOT1)
    std::mutex mu;
};

// tile groups of a plan: which kernel runs them
enum { G_FAST = 0, G_VEC = 1, G_GENERIC = 2, N_GROUPS = 3 };

struct pb200_plan {
    pb200_ctx *ctx = nullptr;
    DevParams P;
    FastParams F;
    TileDev *d_tiles[N_GROUPS] = {};                 // [G_FAST] packed-SIMD persistent kernel,
    CUtensorMap *d_maps[N_GROUPS] = {};              // [G_VEC] / [G_GENERIC] dswx_fused_kernel<true/false>
    int n[N_GROUPS] = {};
    int max_ctas[N_GROUPS] = {};
    FusedTables *d_tables = nullptr;                 // G_FAST only
    bool owns_tables = false;
    ItemDesc *d_items = nullptr;
    TileSlot *d_slots = nullptr;                     // stream_dyn: one record per fast tile (descriptor + sun constants + row info)
    int n_items = 0;
    bool fast_optional = false;                      // some fast tile wants WTR-1 / WTR-2 / CLOUD / SHAD
    bool fast_all_graded = true;                     // every fast tile writes DIAG, WTR, BWTR and CONF
    bool no_fast8 = false;                           // PB200_NO_FAST8=1: run the general variant (A/B, tests)
    int n_fast_seen = 0;
    bool stream = false;                             // every fast tile meets the preconditions of dswx_fused_stream_kernel
    bool stream_dyn = false;                         // ... and it runs as dswx_fused_stream_dyn_kernel (128 x 48 items, 12-row boxes)
    int maps_per_tile = 1;                           // tensor maps per fast tile: 1 (DEM) or ST_MAPS
    bool stream_ordered = false;                     // allocated with cudaMallocAsync
    bool from_arena = false;                         // allocated from a caller-owned arena: nothing to free
    // per input tile: where it went (for launching one tile of the plan on its own)
    std::vector<int> tile_group, tile_slot, tile_ctas, item_start, item_end;
};

extern "C" int pb200_ctx_create(int device, pb200_ctx **out) {
    if (!out) return fail(PB200_E_INVALID_ARG, "pb200_ctx_create: null output");
    *out = nullptr;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(PB200_E_INVALID_ARG, "pb200_ctx_create: device %d of %d", device, n);
    CK(cudaSetDevice(device));
    CK(cudaFree(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(PB200_E_UNSUPPORTED, "pb200_ctx_create: device %d is sm_%d%d; this library is sm_100a only", device,
                    prop.major, prop.minor);
    pb200_ctx *c = new pb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete c;
        return fail(PB200_E_NO_DRIVER_API, "cuTensorMapEncodeTiled not available from the driver");
    }
    c->encode = (EncodeTiledFn)fn;
    {
        // descriptors are allocated stream-ordered from a pool that belongs to this context (the device's default
        // pool, shared with every other cudaMallocAsync user of the process, is left alone); it keeps its memory
        // between calls, so that after the first plan no call reaches the OS allocator
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t pe = cudaMemPoolCreate(&c->pool, &props);
        if (pe == cudaSuccess) {
            unsigned long long keep = ~0ull;
            pe = cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        if (pe == cudaSuccess) pe = cudaStreamCreateWithFlags(&c->s_plan, cudaStreamNonBlocking);
        if (pe != cudaSuccess) {
            if (c->pool) cudaMemPoolDestroy(c->pool);
            delete c;
            return fail_cuda(pe, "pb200_ctx_create: memory pool / stream");
        }
    }
    *out = c;
    return 0;
}

static void pipe_free(HostPipe &p) {
    for (auto &b : p.band) { cudaFree(b); b = nullptr; }
    cudaFree(p.fmask); cudaFree(p.land); cudaFree(p.ocean); cudaFree(p.dem); cudaFree(p.diag);
    p.fmask = p.land = p.ocean = nullptr; p.dem = nullptr; p.diag = nullptr;
    for (auto &b : p.u8out) { cudaFree(b); b = nullptr; }
    p.cap_px = p.cap_dem = 0;
}

extern "C" int pb200_ctx_destroy(pb200_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    for (HostPipe &p : ctx->pipe) {
        if (p.s_out) cudaStreamSynchronize(p.s_out);
        if (p.s_k) cudaStreamSynchronize(p.s_k);
        pipe_free(p);
        cudaFree(p.counters);
        cudaFree(p.tables);
        cudaFree(p.arena.base);
        if (p.s_in) cudaStreamDestroy(p.s_in);
        if (p.s_k) cudaStreamDestroy(p.s_k);
        if (p.s_out) cudaStreamDestroy(p.s_out);
        for (auto e : p.ev_in) cudaEventDestroy(e);
        for (auto e : p.ev_k) cudaEventDestroy(e);
    }
    if (ctx->comm.comm) pb200_comm_destroy(ctx);
    if (ctx->s_plan) cudaStreamDestroy(ctx->s_plan);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
    return 0;
}

// ---------------------------------------------------------------------------
// tile descriptors
// ---------------------------------------------------------------------------
static bool aligned(const void *p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; }

static void sun_terms(const pb200_tile &t, TileDev *d) {
    if (!std::isnan(t.sun_terms[0])) {
        d->sx = t.sun_terms[0]; d->sy = t.sun_terms[1]; d->sz = t.sun_terms[2];
        d->sin_az = t.sun_terms[3]; d->cos_az = t.sun_terms[4];
        return;
    }
    const double deg2rad = 3.14159265358979323846 / 180.0;          // np.radians: x * (pi / 180)
    const double az = t.sun_azimuth * deg2rad;                       // D:4245
    const double zen = (90 - t.sun_elevation) * deg2rad;             // D:4246-4247
    d->sx = std::sin(az) * std::sin(zen);                            // D:4250-4252
    d->sy = std::cos(az) * std::sin(zen);
    d->sz = std::cos(zen);
    d->sin_az = std::sin(az);                                        // D:4276-4277
    d->cos_az = std::cos(az);
}

static int make_tile_dev(pb200_ctx *ctx, const pb200_tile &t, int index, TileDev *d, CUtensorMap *map) {
    std::memset(d, 0, sizeof(*d));
    if (t.height <= 0 || t.width <= 0)
        return fail(PB200_E_INVALID_ARG, "tile %d: empty raster %d x %d", index, t.height, t.width);
    for (int k = 0; k < 6; ++k) {
        if (!t.band[k]) return fail(PB200_E_INVALID_ARG, "tile %d: band %d is NULL", index, k);
        if (!aligned(t.band[k], 2)) return fail(PB200_E_ALIGNMENT, "tile %d: band %d not 2-byte aligned", index, k);
        d->band[k] = t.band[k];
    }
    if (!t.fmask) return fail(PB200_E_INVALID_ARG, "tile %d: fmask is NULL", index);
    d->fmask = t.fmask; d->dem = t.dem; d->land = t.land; d->ocean = t.ocean;
    d->diag = t.diag; d->wtr1 = t.wtr1; d->wtr1r = t.wtr1_remapped; d->wtr2 = t.wtr2; d->cloud = t.cloud;
    d->shad = t.shad; d->wtr = t.wtr; d->bwtr = t.bwtr; d->conf = t.conf;
    d->counters = (unsigned long long *)t.counters;
    d->height = t.height; d->width = t.width;
    d->tiles_x = (t.width + TW - 1) / TW;
    d->n_ctas = d->tiles_x * ((t.height + TH - 1) / TH);
    if (!aligned(t.diag, 2) || !aligned(t.counters, 8))
        return fail(PB200_E_ALIGNMENT, "tile %d: diag / counters misaligned", index);
    if (t.shad && !t.dem) return fail(PB200_E_INVALID_ARG, "tile %d: SHAD output requested without a DEM", index);
    if (t.dem) {
        if (!aligned(t.dem, 4)) return fail(PB200_E_ALIGNMENT, "tile %d: dem not 4-byte aligned", index);
        if (t.dem_off_y < 1 || t.dem_off_x < 1 || t.dem_off_y + t.height + 1 > t.dem_rows ||
            t.dem_off_x + t.width + 1 > t.dem_pitch)
            return fail(PB200_E_INVALID_ARG,
                        "tile %d: the DEM (%d rows x %d, offset %d,%d) must cover the %d x %d tile plus one element on "
                        "every side",
                        index, t.dem_rows, t.dem_pitch, t.dem_off_y, t.dem_off_x, t.height, t.width);
        d->dem_pitch = t.dem_pitch; d->dem_rows = t.dem_rows; d->dem_off_y = t.dem_off_y; d->dem_off_x = t.dem_off_x;
        sun_terms(t, d);
    }
    // vector path: 8-B band loads, 4-B byte-raster loads/stores, 8-B DIAG stores on every row
    bool vec = (t.width % 4) == 0 && aligned(t.fmask, 4) && aligned(t.land, 4) && aligned(t.ocean, 4) &&
               aligned(t.diag, 8) && aligned(t.wtr1, 4) && aligned(t.wtr1_remapped, 4) && aligned(t.wtr2, 4) &&
               aligned(t.cloud, 4) && aligned(t.shad, 4) && aligned(t.wtr, 4) && aligned(t.bwtr, 4) &&
               aligned(t.conf, 4);
    for (int k = 0; k < 6; ++k) vec = vec && aligned(t.band[k], 8);
    if (vec) d->flags |= TF_VEC;
    // TMA needs a 16-B aligned base and a row pitch that is a multiple of 16 B
    if (t.dem && aligned(t.dem, 16) && (t.dem_pitch % 4) == 0) {
        const cuuint64_t gdim[2] = {(cuuint64_t)t.dem_pitch, (cuuint64_t)t.dem_rows};
        const cuuint64_t gstr[1] = {(cuuint64_t)t.dem_pitch * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)SMW, (cuuint32_t)SMH};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)t.dem, gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) d->flags |= TF_TMA;
    }
    return 0;
}

static int plan_build(pb200_ctx *ctx, const pb200_tile *tiles, int n_tiles, const pb200_params *params,
                      pb200_plan *pl, cudaStream_t stream, bool stream_ordered,
                      FusedTables *shared_tables = nullptr, Arena *arena = nullptr) {
    if (!ctx || !tiles || n_tiles <= 0 || !params)
        return fail(PB200_E_INVALID_ARG, "classify: null argument or n_tiles <= 0");
    if (n_tiles > 65535) return fail(PB200_E_INVALID_ARG, "classify: at most 65535 tiles per launch");
    int rc = derive_params(params, &pl->P, true);
    if (rc) return rc;
    build_fast_params(params, pl->P, &pl->F);
    pl->ctx = ctx;
    pl->stream_ordered = stream_ordered;
    {
        const char *e = std::getenv("PB200_NO_FAST8");
        pl->no_fast8 = e && *e && *e != '0';
    }
    std::vector<TileDev> td[N_GROUPS];
    std::vector<CUtensorMap> tm[N_GROUPS];
    std::vector<ItemDesc> items;
    bool stream_ok = true;
    {
        const char *e = std::getenv("PB200_NO_STREAM");
        if (e && *e && *e != '0') stream_ok = false;
    }
    for (int i = 0; i < n_tiles; ++i) {
        TileDev d;
        alignas(64) CUtensorMap m;
        std::memset(&m, 0, sizeof(m));
        rc = make_tile_dev(ctx, tiles[i], i, &d, &m);
        if (rc) return rc;
        const pb200_tile &t = tiles[i];
        bool fast = (d.flags & TF_VEC) && (d.dem == nullptr || (d.flags & TF_TMA)) &&
                    (uint64_t)d.height * (uint64_t)d.width < 0xfff00000ull && d.width <= 65000 * FT_W &&
                    d.height <= 65000 * std::min(std::min(FT_H, ST_H), SD_H);
        (void)t;
        // the first DEM box of a row of items must not start left of the DEM array
        if (d.dem) fast = fast && d.dem_off_x >= DEM_PADX + (d.dem_off_x & 3);
        const int g = fast ? G_FAST : ((d.flags & TF_VEC) ? G_VEC : G_GENERIC);
        pl->tile_group.push_back(g);
        pl->tile_slot.push_back((int)td[g].size());
        pl->tile_ctas.push_back(d.n_ctas);
        if (fast) {
            // TMA-fed variant (pb200_stream.cuh): 4-row super-rows need height % 4 == 0 and 16-byte aligned planes
            bool st = (d.height % 4) == 0 && d.width >= 36 && aligned(d.fmask, 16) && aligned(d.land, 16) && aligned(d.ocean, 16);
            for (int k = 0; k < 6; ++k) st = st && aligned(d.band[k], 16);
            if (!st) stream_ok = false;
            pl->n_fast_seen++;
            // the lean kernel variant writes any subset of the four graded layers and assumes the counters;
            // anything else -> full variant
            if (d.wtr1 || d.wtr1r || d.wtr2 || d.cloud || d.shad || params->class_histogram || !d.counters)
                pl->fast_optional = true;
            if (!d.diag || !d.wtr || !d.bwtr || !d.conf) pl->fast_all_graded = false;
        }
        td[g].push_back(d);
        tm[g].push_back(m);
        pl->max_ctas[g] = std::max(pl->max_ctas[g], d.n_ctas);
    }
    // TMA-fed variant: lean product configuration only (all four graded layers + counters on every fast tile)
    {
        // the lean product configuration (graded layers + counters on every fast tile): rows handed to whichever warp is
        // free, deep ring (default; any subset of the graded layers), or row w of every chunk to warp w
        // (PB200_STREAM_DYNAMIC=0; all four graded layers only)
        const bool eligible = stream_ok && pl->n_fast_seen > 0 && !pl->fast_optional;
        const char *e = std::getenv("PB200_STREAM_DYNAMIC");
        pl->stream_dyn = eligible && !(e && e[0] == '0');
        pl->stream = pl->stream_dyn || (eligible && pl->fast_all_graded);
    }
    // items and DEM boxes of the fast tiles, in the geometry of the kernel that will run them
    {
        const int item_h = pl->stream_dyn ? SD_H : pl->stream ? ST_H : FT_H;
        for (int i = 0; i < n_tiles; ++i) {
            pl->item_start.push_back((int)items.size());
            if (pl->tile_group[i] == G_FAST) {
                const uint32_t slot = (uint32_t)pl->tile_slot[i];
                const TileDev &d = td[G_FAST][slot];
                if (d.dem) {
                    // the fast kernels stage the DEM with their own box shape: item width + pad, item height + 2 halo rows
                    const cuuint64_t gdim[2] = {(cuuint64_t)d.dem_pitch, (cuuint64_t)d.dem_rows};
                    const cuuint64_t gstr[1] = {(cuuint64_t)d.dem_pitch * sizeof(float)};
                    const cuuint32_t box[2] = {(cuuint32_t)FT_SMW, (cuuint32_t)(item_h + 2)};
                    const cuuint32_t estr[2] = {1, 1};
                    const CUresult r = ctx->encode(&tm[G_FAST][slot], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)d.dem, gdim, gstr,
                                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) return fail(PB200_E_INVALID_ARG, "tile %d: cuTensorMapEncodeTiled failed (%d)", i, (int)r);
                }
                const int ntx = (d.width + FT_W - 1) / FT_W, nty = (d.height + item_h - 1) / item_h;
                for (int ty = 0; ty < nty; ++ty)
                    for (int tx = 0; tx < ntx; ++tx) items.push_back(ItemDesc{slot, (uint16_t)tx, (uint16_t)ty});
            }
            pl->item_end.push_back((int)items.size());
        }
    }
    if (pl->stream) {
        // ten tensor maps per fast tile: DEM, six bands, Fmask, LAND, ocean (4-row super-rows, see pb200_stream.cuh)
        std::vector<CUtensorMap> all(td[G_FAST].size() * ST_MAPS);
        std::memset(all.data(), 0, all.size() * sizeof(CUtensorMap));
        for (size_t i = 0; i < td[G_FAST].size(); ++i) {
            const TileDev &d = td[G_FAST][i];
            CUtensorMap *m = &all[i * ST_MAPS];
            m[SM_DEM] = tm[G_FAST][i];
            const cuuint64_t gdim[2] = {(cuuint64_t)4 * d.width, (cuuint64_t)d.height / 4};
            const cuuint32_t estr[2] = {1, 1};
            auto enc = [&](CUtensorMap *dst, CUtensorMapDataType dt, const void *base, size_t esz, int box_w) -> bool {
                if (!base) return true;
                const cuuint64_t gstr[1] = {(cuuint64_t)4 * d.width * esz};
                const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)(pl->stream_dyn ? SD_CH : ST_WARPS)};
                return ctx->encode(dst, dt, 2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            };
            bool ok = true;
            for (int k = 0; k < 6; ++k) ok = ok && enc(&m[SM_BAND0 + k], CU_TENSOR_MAP_DATA_TYPE_UINT16, d.band[k], 2, ST_BAND_W);
            ok = ok && enc(&m[SM_FMASK], CU_TENSOR_MAP_DATA_TYPE_UINT8, d.fmask, 1, ST_BYTE_W);
            ok = ok && enc(&m[SM_LAND], CU_TENSOR_MAP_DATA_TYPE_UINT8, d.land, 1, ST_BYTE_W);
            ok = ok && enc(&m[SM_OCEAN], CU_TENSOR_MAP_DATA_TYPE_UINT8, d.ocean, 1, ST_BYTE_W);
            if (!ok) { pl->stream = false; break; }     // (cannot happen after the checks above; the items would have the wrong height)
        }
        if (pl->stream) {
            tm[G_FAST].swap(all);
            pl->maps_per_tile = ST_MAPS;
        }
    }
    CK(cudaSetDevice(ctx->device));
    pl->from_arena = arena != nullptr;
    auto dev_alloc = [&](void **ptr, size_t bytes) -> cudaError_t {
        if (arena) {
            *ptr = arena->take(bytes);
            return *ptr ? cudaSuccess : cudaErrorMemoryAllocation;
        }
        return stream_ordered ? cudaMallocFromPoolAsync(ptr, bytes, ctx->pool, stream) : cudaMalloc(ptr, bytes);
    };
    for (int g = 0; g < N_GROUPS; ++g) {
        pl->n[g] = (int)td[g].size();
        if (!pl->n[g]) continue;
        const size_t bt = td[g].size() * sizeof(TileDev), bm = tm[g].size() * sizeof(CUtensorMap);
        CK(dev_alloc((void **)&pl->d_tiles[g], bt));
        CK(dev_alloc((void **)&pl->d_maps[g], bm));
        // pageable source: the copy is staged before the call returns
        CK(cudaMemcpyAsync(pl->d_tiles[g], td[g].data(), bt, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(pl->d_maps[g], tm[g].data(), bm, cudaMemcpyHostToDevice, stream));
    }
    if (pl->n[G_FAST]) {
        pl->n_items = (int)items.size();
        // (one spare element: the dynamic kernel copies the aligned 16 bytes around an item)
        items.push_back(ItemDesc{0u, 0, 0});
        CK(dev_alloc((void **)&pl->d_items, items.size() * sizeof(ItemDesc)));
        CK(cudaMemcpyAsync(pl->d_items, items.data(), items.size() * sizeof(ItemDesc), cudaMemcpyHostToDevice, stream));
        if (pl->stream_dyn) {
            // per-tile records of dswx_fused_stream_dyn_kernel: descriptor, float32 sun constants of the shadow shortcut
            // (the arithmetic of the other fast kernels' per-tile prologue, done once here), first-look vector of a row
            std::vector<TileSlot> slots(td[G_FAST].size());
            const bool f8 = pl->F.fast8 != 0u && !pl->no_fast8;
            for (size_t i = 0; i < slots.size(); ++i) {
                TileSlot &sl = slots[i];
                std::memset(&sl, 0, sizeof(sl));
                sl.tile = td[G_FAST][i];
                const TileDev &g = sl.tile;
                const uint32_t tsf = (g.dem ? TSF_DEM : 0u) | (g.land ? TSF_LAND : 0u) | (g.ocean ? TSF_OCEAN : 0u);
                sl.tile.pad_ = tsf;
                const double kx = 0.5 / (double)pl->P.dxf, ky = 0.5 / (double)pl->P.dyf;
                float *K = sl.sun32;
                K[SK_SA] = (float)(kx * g.sin_az); K[SK_CA] = (float)(ky * g.cos_az);
                K[SK_SX] = (float)(kx * g.sx); K[SK_SY] = (float)(ky * g.sy); K[SK_SZ] = (float)g.sz;
                K[SK_XX] = (float)(kx * kx);
                K[SK_EA] = 1e-6f * std::fabs(K[SK_SA]); K[SK_EB] = 1e-6f * std::fabs(K[SK_CA]);
                if (f8) {
                    // tiles whose sun vector breaks the preconditions of the sign-bit shortcut run the exact sequence
                    const double hz = g.sx * g.sin_az + g.sy * g.cos_az, n2 = g.sx * g.sx + g.sy * g.sy + g.sz * g.sz;
                    if (!(hz >= 0.0 && std::fabs(n2 - 1.0) < 1e-9 && std::fabs(g.sin_az * g.sin_az + g.cos_az * g.cos_az - 1.0) < 1e-9)) {
                        const uint32_t nanbits = 0x7fffffffu;
                        std::memcpy(&K[SK_XX], &nanbits, 4);
                    }
                }
                sl.row_info = make_uint4((uint32_t)g.width, (uint32_t)g.height, (uint32_t)(DEM_PADX + (g.dem_off_x & 3)),
                                         tsf | ((uint32_t)i << 3));
                sl.out_ptrs[0] = (unsigned long long)(uintptr_t)g.diag; sl.out_ptrs[1] = (unsigned long long)(uintptr_t)g.wtr;
                sl.out_ptrs[2] = (unsigned long long)(uintptr_t)g.bwtr; sl.out_ptrs[3] = (unsigned long long)(uintptr_t)g.conf;
            }
            CK(dev_alloc((void **)&pl->d_slots, slots.size() * sizeof(TileSlot)));
            CK(cudaMemcpyAsync(pl->d_slots, slots.data(), slots.size() * sizeof(TileSlot), cudaMemcpyHostToDevice, stream));
        }
        if (shared_tables) {
            pl->d_tables = shared_tables;
        } else {
            FusedTables T;
            build_fused_tables(params, pl->P, &T);
            CK(dev_alloc((void **)&pl->d_tables, sizeof(FusedTables)));
            pl->owns_tables = true;
            CK(cudaMemcpyAsync(pl->d_tables, &T, sizeof(T), cudaMemcpyHostToDevice, stream));
        }
    }
    return 0;
}

// FastSmem (double-buffered DEM tile + tables, 51 KB) is dynamic shared memory: above the 48 KB static limit
constexpr size_t FAST_DYN_SMEM = sizeof(FastSmem);

static int fast_kernel_setup(pb200_ctx *ctx) {
    if (ctx->fast_ready) return 0;
    if (FAST_DYN_SMEM) {
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
        CK(cudaFuncSetAttribute(dswx_fused_fast_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)FAST_DYN_SMEM));
    }
    CK(cudaFuncSetAttribute(dswx_fused_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)));
    CK(cudaFuncSetAttribute(dswx_fused_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)));
    CK(cudaFuncSetAttribute(dswx_fused_stream_dyn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamDynSmem)));
    CK(cudaFuncSetAttribute(dswx_fused_stream_dyn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamDynSmem)));
    CK(cudaFuncSetAttribute(dswx_fused_stream_dyn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamDynSmem)));
    CK(cudaFuncSetAttribute(dswx_fused_stream_dyn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamDynSmem)));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dswx_fused_fast_kernel<false>, FtGeom<false>::THREADS, FAST_DYN_SMEM));
    ctx->fast_ctas_per_sm = nb > 0 ? nb : 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, dswx_fused_fast_kernel<true>, FtGeom<true>::THREADS, FAST_DYN_SMEM));
    ctx->fast_ctas_per_sm_full = nb > 0 ? nb : 1;
    ctx->fast_ready = true;
    return 0;
}

// the fast kernel over `n` items: variant by what the tiles write and by the shape of the parameters
static void launch_fast(pb200_plan *pl, const ItemDesc *it, int n, cudaStream_t stream) {
    const pb200_ctx *ctx = pl->ctx;
    const int grid = std::min(n, ctx->sm_count * (pl->fast_optional ? ctx->fast_ctas_per_sm_full : ctx->fast_ctas_per_sm));
    const bool f8 = pl->F.fast8 != 0u && !pl->no_fast8;
    if (pl->stream) {
        const int g1 = std::min(n, ctx->sm_count);            // 219 KB of shared memory: one CTA per SM
        if (pl->stream_dyn) {
#define PB200_LAUNCH_DYN(F8, GRADED)                                                                       \
    dswx_fused_stream_dyn_kernel<F8, GRADED><<<g1, ST_THREADS, sizeof(StreamDynSmem), stream>>>(           \
        pl->d_slots, pl->d_maps[G_FAST], pl->d_tables, it, n, pl->P, pl->F)
            if (pl->fast_all_graded) { if (f8) PB200_LAUNCH_DYN(true, true); else PB200_LAUNCH_DYN(false, true); }
            else { if (f8) PB200_LAUNCH_DYN(true, false); else PB200_LAUNCH_DYN(false, false); }
#undef PB200_LAUNCH_DYN
            return;
        }
        if (f8)
            dswx_fused_stream_kernel<true><<<g1, ST_THREADS, sizeof(StreamSmem), stream>>>(
                pl->d_tiles[G_FAST], pl->d_maps[G_FAST], pl->d_tables, it, n, pl->P, pl->F);
        else
            dswx_fused_stream_kernel<false><<<g1, ST_THREADS, sizeof(StreamSmem), stream>>>(
                pl->d_tiles[G_FAST], pl->d_maps[G_FAST], pl->d_tables, it, n, pl->P, pl->F);
        return;
    }
#define PB200_LAUNCH_FAST(OPT, GRADED, F8)                                                                         \
    dswx_fused_fast_kernel<OPT, GRADED, F8><<<grid, FtGeom<OPT>::THREADS, FAST_DYN_SMEM, stream>>>(                 \
        pl->d_tiles[G_FAST], pl->d_maps[G_FAST], pl->d_tables, it, n, pl->P, pl->F)
    if (pl->fast_optional) { if (f8) PB200_LAUNCH_FAST(true, false, true); else PB200_LAUNCH_FAST(true, false, false); }
    else if (pl->fast_all_graded) { if (f8) PB200_LAUNCH_FAST(false, true, true); else PB200_LAUNCH_FAST(false, true, false); }
    else { if (f8) PB200_LAUNCH_FAST(false, false, true); else PB200_LAUNCH_FAST(false, false, false); }
#undef PB200_LAUNCH_FAST
}

static int plan_launch(pb200_plan *pl, cudaStream_t stream) {
    if (pl->n[G_FAST]) {
        int rc = fast_kernel_setup(pl->ctx);
        if (rc) return rc;
        launch_fast(pl, pl->d_items, pl->n_items, stream);
    }
    if (pl->n[G_VEC]) {
        dim3 grid(pl->max_ctas[G_VEC], pl->n[G_VEC]);
        dswx_fused_kernel<true><<<grid, NTHREADS, 0, stream>>>(pl->d_tiles[G_VEC], pl->d_maps[G_VEC], pl->P);
    }
    if (pl->n[G_GENERIC]) {
        dim3 grid(pl->max_ctas[G_GENERIC], pl->n[G_GENERIC]);
        dswx_fused_kernel<false><<<grid, NTHREADS, 0, stream>>>(pl->d_tiles[G_GENERIC], pl->d_maps[G_GENERIC], pl->P);
    }
    CK(cudaGetLastError());
    return 0;
}

// launch only input tile `i` of the plan (host pipeline: one strip at a time)
static int plan_launch_tile(pb200_plan *pl, int i, cudaStream_t stream) {
    const int g = pl->tile_group[i], slot = pl->tile_slot[i];
    if (g == G_FAST) {
        int rc = fast_kernel_setup(pl->ctx);
        if (rc) return rc;
        launch_fast(pl, pl->d_items + pl->item_start[i], pl->item_end[i] - pl->item_start[i], stream);
    } else if (g == G_VEC) {
        dswx_fused_kernel<true><<<dim3(pl->tile_ctas[i], 1), NTHREADS, 0, stream>>>(pl->d_tiles[g] + slot,
                                                                                    pl->d_maps[g] + slot, pl->P);
    } else {
        dswx_fused_kernel<false><<<dim3(pl->tile_ctas[i], 1), NTHREADS, 0, stream>>>(pl->d_tiles[g] + slot,
                                                                                     pl->d_maps[g] + slot, pl->P);
    }
    CK(cudaGetLastError());
    return 0;
}

static void plan_release(pb200_plan *pl, cudaStream_t stream) {
    auto dev_free = [&](void *ptr) {
        if (!ptr || pl->from_arena) return;
        if (pl->stream_ordered) cudaFreeAsync(ptr, stream); else cudaFree(ptr);
    };
    for (int g = 0; g < N_GROUPS; ++g) {
        dev_free(pl->d_tiles[g]);
        dev_free(pl->d_maps[g]);
        pl->d_tiles[g] = nullptr;
        pl->d_maps[g] = nullptr;
    }
    dev_free(pl->d_items);
    pl->d_items = nullptr;
    dev_free(pl->d_slots);
    pl->d_slots = nullptr;
    if (pl->owns_tables) dev_free(pl->d_tables);
    pl->d_tables = nullptr;
}

extern "C" int pb200_classify(pb200_ctx *ctx, const pb200_tile *tiles, int n_tiles, const pb200_params *params,
                              void *stream) {
    pb200_plan pl;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = plan_build(ctx, tiles, n_tiles, params, &pl, st, true);
    if (rc == 0) rc = plan_launch(&pl, st);
    plan_release(&pl, st);
    return rc;
}

extern "C" int pb200_plan_create(pb200_ctx *ctx, const pb200_tile *tiles, int n_tiles, const pb200_params *params,
                                 pb200_plan **out) {
    if (!out) return fail(PB200_E_INVALID_ARG, "pb200_plan_create: null output");
    *out = nullptr;
    pb200_plan *pl = new pb200_plan();
    // stream-ordered allocations from the context's memory pool (release threshold: never): after the first plan a
    // create / destroy pair costs no cudaMalloc / cudaFree (those took ~1 ms per pair).  Uploads run on the context's
    // own non-blocking stream (never the legacy NULL stream, which would serialise with every blocking stream of the
    // host application and cannot be captured); the plan is complete when this call returns.
    if (!ctx) { delete pl; return fail(PB200_E_INVALID_ARG, "pb200_plan_create: null context"); }
    std::lock_guard<std::mutex> lock(ctx->mu);
    int rc = plan_build(ctx, tiles, n_tiles, params, pl, ctx->s_plan, true);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(ctx->s_plan);
        if (e != cudaSuccess) rc = fail_cuda(e, "plan upload");
    }
    if (rc) {
        plan_release(pl, ctx->s_plan);
        delete pl;
        return rc;
    }
    *out = pl;
    return 0;
}
// which kernels a plan launches: bit 0 = dswx_fused_fast_kernel, bit 1 = dswx_fused_stream_kernel (TMA-fed), bit 2 =
// FAST8 flavour of either, bit 3 = dswx_fused_kernel (generic path) for some tile
extern "C" int pb200_plan_kernels(const pb200_plan *plan, int *mask) {
    if (!plan || !mask) return fail(PB200_E_INVALID_ARG, "pb200_plan_kernels: null argument");
    int m = 0;
    if (plan->n[G_FAST]) m |= plan->stream ? 2 : 1;
    if (plan->n[G_FAST] && plan->stream_dyn) m |= 16;
    if (plan->n[G_FAST] && plan->F.fast8 != 0u && !plan->no_fast8) m |= 4;
    if (plan->n[G_VEC] || plan->n[G_GENERIC]) m |= 8;
    *mask = m;
    return 0;
}
extern "C" int pb200_plan_run(pb200_plan *plan, void *stream) {
    if (!plan) return fail(PB200_E_INVALID_ARG, "pb200_plan_run: null plan");
    return plan_launch(plan, (cudaStream_t)stream);
}
extern "C" int pb200_plan_destroy(pb200_plan *plan) {
    if (!plan) return 0;
    cudaSetDevice(plan->ctx->device);
    // the plan may still be running on any stream: wait for the device (what cudaFree used to do implicitly), then
    // hand the buffers back to the pool
    cudaDeviceSynchronize();
    plan_release(plan, plan->ctx->s_plan);
    delete plan;
    return 0;
}

// ---------------------------------------------------------------------------
// host-buffer entry point: H2D / kernel / D2H overlapped over row strips
// ---------------------------------------------------------------------------
extern "C" int pb200_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(PB200_E_INVALID_ARG, "pb200_host_alloc: null output");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}
extern "C" int pb200_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}

static int pipe_reserve(HostPipe &p, size_t px, size_t dem_elems) {
    if (!p.s_in) {
        CK(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
        CK(cudaMalloc((void **)&p.counters, PB200_N_COUNTERS * sizeof(unsigned long long)));
        CK(cudaMalloc((void **)&p.tables, sizeof(FusedTables)));
    }
    if (px > p.cap_px) {
        // capacity and pointers are cleared BEFORE anything is freed: a failed cudaMalloc below leaves an empty
        // mirror (the next call reallocates), never freed pointers behind a stale capacity
        p.cap_px = 0;
        p.anc_valid = false;
        for (auto &b : p.band) { cudaFree(b); b = nullptr; }
        cudaFree(p.fmask); cudaFree(p.land); cudaFree(p.ocean); cudaFree(p.diag);
        p.fmask = p.land = p.ocean = nullptr; p.diag = nullptr;
        for (auto &b : p.u8out) { cudaFree(b); b = nullptr; }
        for (auto &b : p.band) CK(cudaMalloc((void **)&b, px * 2));
        CK(cudaMalloc((void **)&p.fmask, px));
        CK(cudaMalloc((void **)&p.land, px));
        CK(cudaMalloc((void **)&p.ocean, px));
        CK(cudaMalloc((void **)&p.diag, px * 2));
        for (auto &b : p.u8out) CK(cudaMalloc((void **)&b, px));
        p.cap_px = px;
    }
    if (dem_elems > p.cap_dem) {
        p.cap_dem = 0;
        p.anc_valid = false;
        cudaFree(p.dem);
        p.dem = nullptr;
        CK(cudaMalloc((void **)&p.dem, dem_elems * 4));
        p.cap_dem = dem_elems;
    }
    return 0;
}

// error path of the host pipeline: asynchronous copies from / into the caller's buffers may be in flight on the three
// streams; none may outlive the call
static void pipe_drain(HostPipe &p) {
    if (p.s_in) cudaStreamSynchronize(p.s_in);
    if (p.s_k) cudaStreamSynchronize(p.s_k);
    if (p.s_out) cudaStreamSynchronize(p.s_out);
}
#define CKP(call)                                             \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) {                              \
            pipe_drain(p);                                    \
            return fail_cuda(e_, #call);                      \
        }                                                     \
    } while (0)

// complete the tile enqueued on pipe `p` (outputs in host memory) and record its ancillary rasters as resident
static int host_wait_locked(pb200_ctx *ctx, HostPipe &p) {
    if (!p.pending) return 0;
    p.pending = false;
    CK(cudaSetDevice(ctx->device));
    CKP(cudaStreamSynchronize(p.s_out));
    CKP(cudaStreamSynchronize(p.s_k));
    const pb200_tile *ht = &p.pending_geom;
    p.anc_valid = true;
    p.anc_h = ht->height; p.anc_w = ht->width;
    p.anc_dem = ht->dem != nullptr; p.anc_land = ht->land != nullptr; p.anc_ocean = ht->ocean != nullptr;
    p.anc_dem_rows = ht->dem_rows; p.anc_dem_pitch = ht->dem_pitch;
    p.anc_dem_off_y = ht->dem_off_y; p.anc_dem_off_x = ht->dem_off_x;
    return 0;
}

extern "C" int pb200_host_wait(pb200_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot > 1) return fail(PB200_E_INVALID_ARG, "pb200_host_wait: bad argument");
    std::lock_guard<std::mutex> lock(ctx->mu);
    return host_wait_locked(ctx, ctx->pipe[slot]);
}

extern "C" int pb200_classify_host(pb200_ctx *ctx, const pb200_tile *ht, const pb200_params *params,
                                   int strip_rows) {
    return pb200_classify_host_ex(ctx, ht, params, strip_rows, 0);
}

extern "C" int pb200_classify_host_ex(pb200_ctx *ctx, const pb200_tile *ht, const pb200_params *params,
                                      int strip_rows, int flags) {
    if (!ctx || !ht || !params) return fail(PB200_E_INVALID_ARG, "pb200_classify_host: null argument");
    if (flags & ~(PB200_HOST_REUSE_ANCILLARY | PB200_HOST_ASYNC | PB200_HOST_SLOT1))
        return fail(PB200_E_INVALID_ARG, "pb200_classify_host_ex: unknown flag");
    const bool reuse = (flags & PB200_HOST_REUSE_ANCILLARY) != 0, async = (flags & PB200_HOST_ASYNC) != 0;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const int H = ht->height, W = ht->width;
    if (H <= 0 || W <= 0) return fail(PB200_E_INVALID_ARG, "pb200_classify_host: empty raster");
    if (strip_rows <= 0) strip_rows = 33 * TH;      // 1056 rows = 11 items of the fast kernel: 4-5 strips per HLS tile
    strip_rows = ((strip_rows + TH - 1) / TH) * TH;
    const size_t px = (size_t)H * W;
    const size_t dem_elems = ht->dem ? (size_t)ht->dem_rows * ht->dem_pitch : 0;
    HostPipe &p = ctx->pipe[(flags & PB200_HOST_SLOT1) ? 1 : 0];
    if (p.pending) {                                      // the slot's previous tile: complete it first
        int rcw = host_wait_locked(ctx, p);
        if (rcw) return rcw;
    }
    if (reuse) {
        const bool same = p.anc_valid && p.anc_h == H && p.anc_w == W && p.anc_dem == (ht->dem != nullptr) &&
                          p.anc_land == (ht->land != nullptr) && p.anc_ocean == (ht->ocean != nullptr) &&
                          (!ht->dem || (p.anc_dem_rows == ht->dem_rows && p.anc_dem_pitch == ht->dem_pitch &&
                                        p.anc_dem_off_y == ht->dem_off_y && p.anc_dem_off_x == ht->dem_off_x));
        if (!same)
            return fail(PB200_E_INVALID_ARG, "pb200_classify_host_ex: PB200_HOST_REUSE_ANCILLARY needs a previous "
                        "successful call on this context with the same tile size, DEM geometry and rasters present");
    }
    p.anc_valid = false;                                  // until this call has succeeded
    int rc = pipe_reserve(p, px, dem_elems);
    if (rc) return rc;
    // PB200_PIPE_TRACE=1: print where the time of this call went (host stages, H2D / kernel / D2H spans)
    static const bool trace = std::getenv("PB200_PIPE_TRACE") != nullptr;
    const auto t_enter = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    cudaEvent_t tr[4] = {};
    if (trace)
        for (auto &e : tr) CK(cudaEventCreate(&e));
    // strip boundaries: strip_rows each, the tail of the raster in shorter strips (down to 8 * TH rows) so that the
    // kernel + D2H of the last strip - the only part of the pipeline that nothing overlaps - stays short
    std::vector<int> edge;
    edge.push_back(0);
    for (int r = strip_rows; r < H; r += strip_rows) edge.push_back(r);
    edge.push_back(H);
    for (int min_rows = 8 * TH; ; ) {
        const int n = (int)edge.size(), last = edge[n - 1] - edge[n - 2];
        if (last <= min_rows + TH) break;
        int cut = edge[n - 2] + ((last / 2 + TH - 1) / TH) * TH;
        if (cut >= H) break;
        edge.insert(edge.end() - 1, cut);
    }
    const int n_strips = (int)edge.size() - 1;
    while ((int)p.ev_in.size() < n_strips) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p.ev_in.push_back(e);
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p.ev_k.push_back(e);
    }
    DevParams P;
    rc = derive_params(params, &P, true);             // argument checks before anything is in flight
    if (rc) return rc;
    // ---- H2D of every strip first: the copy engine starts while the host still builds tables and plan ----------
    if (trace) { CKP(cudaEventRecord(tr[0], p.s_in)); }
    {
        int dem_copied = 0;                 // DEM rows [.., dem_copied) are on the device
        bool first_dem = true;
        for (int sidx = 0; sidx < n_strips; ++sidx) {
            const int r0 = edge[sidx], r1 = edge[sidx + 1], nr = r1 - r0;
            const size_t off = (size_t)r0 * W, cnt = (size_t)nr * W;
            for (int k = 0; k < 6; ++k)
                CKP(cudaMemcpyAsync(p.band[k] + off, ht->band[k] + off, cnt * 2, cudaMemcpyHostToDevice, p.s_in));
            CKP(cudaMemcpyAsync(p.fmask + off, ht->fmask + off, cnt, cudaMemcpyHostToDevice, p.s_in));
            if (ht->land && !reuse) CKP(cudaMemcpyAsync(p.land + off, ht->land + off, cnt, cudaMemcpyHostToDevice, p.s_in));
            if (ht->ocean && !reuse) CKP(cudaMemcpyAsync(p.ocean + off, ht->ocean + off, cnt, cudaMemcpyHostToDevice, p.s_in));
            if (ht->dem && !reuse) {
                int d0 = ht->dem_off_y + r0 - 1, d1 = ht->dem_off_y + r1 + 1;     // rows the strip's stencil reads
                if (!first_dem) d0 = std::max(d0, dem_copied);
                d0 = std::max(d0, 0);
                d1 = std::min(d1, ht->dem_rows);
                if (d1 > d0)
                    CKP(cudaMemcpyAsync(p.dem + (size_t)d0 * ht->dem_pitch, ht->dem + (size_t)d0 * ht->dem_pitch,
                                       (size_t)(d1 - d0) * ht->dem_pitch * 4, cudaMemcpyHostToDevice, p.s_in));
                dem_copied = std::max(dem_copied, d1);
                first_dem = false;
            }
            CKP(cudaEventRecord(p.ev_in[sidx], p.s_in));
        }
        if (trace) { CKP(cudaEventRecord(tr[1], p.s_in)); }
    }
    const double ms_h2d_enqueued = ms_since(t_enter);
    // validate once with a whole-tile descriptor on the device mirror
    pb200_tile dt = *ht;
    for (int k = 0; k < 6; ++k) dt.band[k] = p.band[k];
    dt.fmask = p.fmask;
    dt.land = ht->land ? p.land : nullptr;
    dt.ocean = ht->ocean ? p.ocean : nullptr;
    dt.dem = ht->dem ? p.dem : nullptr;
    dt.diag = ht->diag ? p.diag : nullptr;
    uint8_t *const host_u8[8] = {ht->wtr1, ht->wtr1_remapped, ht->wtr2, ht->cloud, ht->shad, ht->wtr, ht->bwtr, ht->conf};
    uint8_t **dev_u8_field[8] = {&dt.wtr1, &dt.wtr1_remapped, &dt.wtr2, &dt.cloud, &dt.shad, &dt.wtr, &dt.bwtr, &dt.conf};
    for (int i = 0; i < 8; ++i) *dev_u8_field[i] = host_u8[i] ? p.u8out[i] : nullptr;
    dt.counters = (uint64_t *)p.counters;             // always counted on the device (lean kernel variant); copied back on request
    if (dt.counters) CKP(cudaMemsetAsync(p.counters, 0, PB200_N_COUNTERS * sizeof(unsigned long long), p.s_k));
    {
        FusedTables T;
        build_fused_tables(params, P, &T);
        CKP(cudaMemcpyAsync(p.tables, &T, sizeof(T), cudaMemcpyHostToDevice, p.s_k));
    }

    // descriptors of ALL strips in one plan, uploaded before the pipeline starts: inside the
    // loop the host only enqueues asynchronous work and never waits for the GPU
    std::vector<pb200_tile> strips(n_strips);
    for (int sidx = 0; sidx < n_strips; ++sidx) {
        const int r0 = edge[sidx], r1 = edge[sidx + 1];
        const size_t off = (size_t)r0 * W;
        pb200_tile &st = strips[sidx];
        st = dt;
        st.height = r1 - r0;
        for (int k = 0; k < 6; ++k) st.band[k] = dt.band[k] + off;
        st.fmask = dt.fmask + off;
        if (st.land) st.land = dt.land + off;
        if (st.ocean) st.ocean = dt.ocean + off;
        st.dem_off_y = dt.dem_off_y + r0;
        if (st.diag) st.diag = dt.diag + off;
        uint8_t **sf[8] = {&st.wtr1, &st.wtr1_remapped, &st.wtr2, &st.cloud, &st.shad, &st.wtr, &st.bwtr, &st.conf};
        for (int i = 0; i < 8; ++i)
            if (*sf[i]) *sf[i] += off;
    }
    // worst case per strip: descriptor + tensor map + one item per FT_W x FT_H pixels
    {
        const size_t items_max = (size_t)((W + FT_W - 1) / FT_W) * (size_t)((H + SD_H - 1) / SD_H + n_strips);
        const size_t need = (size_t)n_strips * (sizeof(TileDev) + sizeof(TileSlot) + ST_MAPS * sizeof(CUtensorMap) + 1024) +
                            items_max * sizeof(ItemDesc) + 8192;
        if (need > p.arena.cap) {
            CKP(cudaStreamSynchronize(p.s_k));
            cudaFree(p.arena.base);
            p.arena.base = nullptr;
            p.arena.cap = 0;
            CKP(cudaMalloc((void **)&p.arena.base, need * 2));
            p.arena.cap = need * 2;
        }
        p.arena.used = 0;
    }
    pb200_plan plan;
    rc = plan_build(ctx, strips.data(), n_strips, params, &plan, p.s_k, true, p.tables, &p.arena);
    if (rc) {
        plan_release(&plan, p.s_k);
        pipe_drain(p);                                // the copies read the caller's buffers: none may outlive the call
        return rc;
    }

    const double ms_prologue = ms_since(t_enter);
    for (int sidx = 0; sidx < n_strips; ++sidx) {
        const int r0 = edge[sidx], r1 = edge[sidx + 1], nr = r1 - r0;
        const size_t off = (size_t)r0 * W, cnt = (size_t)nr * W;
        // ---- kernel on the strip -------------------------------------------
        CKP(cudaStreamWaitEvent(p.s_k, p.ev_in[sidx], 0));
        rc = plan_launch_tile(&plan, sidx, p.s_k);
        if (rc) { pipe_drain(p); plan_release(&plan, p.s_k); return rc; }
        CKP(cudaEventRecord(p.ev_k[sidx], p.s_k));
        // ---- D2H --------------------------------------------------------------
        CKP(cudaStreamWaitEvent(p.s_out, p.ev_k[sidx], 0));
        if (ht->diag) CKP(cudaMemcpyAsync(ht->diag + off, p.diag + off, cnt * 2, cudaMemcpyDeviceToHost, p.s_out));
        for (int i = 0; i < 8; ++i)
            if (host_u8[i])
                CKP(cudaMemcpyAsync(host_u8[i] + off, p.u8out[i] + off, cnt, cudaMemcpyDeviceToHost, p.s_out));
    }
    const double ms_enqueued = ms_since(t_enter);
    if (trace) { CKP(cudaEventRecord(tr[2], p.s_k)); }
    plan_release(&plan, p.s_k);
    if (ht->counters)
        CKP(cudaMemcpyAsync(ht->counters, p.counters, PB200_N_COUNTERS * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, p.s_out));
    if (trace) { CKP(cudaEventRecord(tr[3], p.s_out)); }
    p.pending_geom = *ht;
    p.pending = true;
    if (async) {
        // two tiles in flight: the caller enqueues the next tile on the other slot before it waits for this one
        if (trace)
            for (auto &e : tr) cudaEventDestroy(e);
        return 0;
    }
    rc = host_wait_locked(ctx, p);
    if (rc) return rc;
    if (trace) {
        float h2d = 0, k_end = 0, out_end = 0;
        cudaEventElapsedTime(&h2d, tr[0], tr[1]);
        cudaEventElapsedTime(&k_end, tr[0], tr[2]);
        cudaEventElapsedTime(&out_end, tr[0], tr[3]);
        std::fprintf(stderr, "[pb200 pipe] H2D enqueued at %.3f ms, tables + plan ready at %.3f ms, all work enqueued at %.3f ms, "
                     "call %.3f ms | from first H2D: H2D done %.3f, last kernel done %.3f, last D2H done %.3f ms (%d strips)\n",
                     ms_h2d_enqueued, ms_prologue, ms_enqueued, ms_since(t_enter), h2d, k_end, out_end, n_strips);
        for (auto &e : tr) cudaEventDestroy(e);
    }
    return 0;
}

// ---------------------------------------------------------------------------
// function-granular entry points
// ---------------------------------------------------------------------------
static int grid_for(pb200_ctx *ctx, long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = (long long)ctx->sm_count * 16;
    return (int)std::max(1LL, std::min(g, cap));
}
#define REQUIRE(cond, ...) \
    do { if (!(cond)) return fail(PB200_E_INVALID_ARG, __VA_ARGS__); } while (0)
#define ENTER(ctx)                                         \
    REQUIRE(ctx != nullptr, "%s: null context", __func__); \
    CK(cudaSetDevice(ctx->device));                        \
    cudaStream_t st = (cudaStream_t)stream
#define LEAVE() CK(cudaGetLastError()); return 0
#define EMPTY_OK(n) do { if ((n) == 0) return 0; } while (0)

extern "C" int pb200_invalid_and_clip(pb200_ctx *ctx, const int16_t *const raw[6], const uint8_t *fmask,
                                      const pb200_params *params, int64_t n, int16_t *const clipped[6],
                                      uint8_t *invalid, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(raw && params && n >= 0, "pb200_invalid_and_clip: bad argument");
    if (n == 0) return 0;
    DevParams P;
    std::memset(&P, 0, sizeof(P));
    for (int k = 0; k < 6; ++k) P.band_fill[k] = params->band_fill[k];
    P.fmask_fill = params->fmask_fill;
    BandPtrs in; BandOutPtrs out;
    for (int k = 0; k < 6; ++k) {
        REQUIRE(raw[k], "pb200_invalid_and_clip: band %d is NULL", k);
        in.p[k] = raw[k];
        out.p[k] = clipped ? clipped[k] : nullptr;
    }
    invalid_and_clip_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(in, fmask, out, invalid, n, P);
    LEAVE();
}

extern "C" int pb200_diagnostic_tests(pb200_ctx *ctx, const int16_t *const band[6], const pb200_thresholds *th,
                                      int64_t n, uint16_t *diag, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(band && th && diag && n >= 0, "pb200_diagnostic_tests: bad argument");
    if (n == 0) return 0;
    DevParams P;
    std::memset(&P, 0, sizeof(P));
    int rc = derive_thresholds(*th, &P);
    if (rc) return rc;
    BandPtrs in;
    for (int k = 0; k < 6; ++k) {
        REQUIRE(band[k], "pb200_diagnostic_tests: band %d is NULL", k);
        in.p[k] = band[k];
    }
    diagnostic_tests_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(in, diag, n, P);
    LEAVE();
}

extern "C" int pb200_diagnostic_tests_f32(pb200_ctx *ctx, const float *const band[6], const pb200_thresholds *th,
                                          int64_t n, uint16_t *diag, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(band && th && diag && n >= 0, "pb200_diagnostic_tests_f32: bad argument");
    BandPtrsF in;
    for (int k = 0; k < 6; ++k) {
        REQUIRE(band[k], "pb200_diagnostic_tests_f32: band %d is NULL", k);
        in.p[k] = band[k];
    }
    // numpy compares a float32 array with a Python scalar in float32 (NEP 50 weak scalar; same under 1.23)
    ThresholdsF T{(float)th->wigt, (float)th->awgt, (float)th->pswt_1_mndwi, (float)th->pswt_1_nir,
                  (float)th->pswt_1_swir1, (float)th->pswt_1_ndvi, (float)th->pswt_2_mndwi, (float)th->pswt_2_blue,
                  (float)th->pswt_2_nir, (float)th->pswt_2_swir1, (float)th->pswt_2_swir2};
    diagnostic_tests_f32_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(in, diag, n, T);
    LEAVE();
}

extern "C" int pb200_interpreted_layer(pb200_ctx *ctx, const uint16_t *diag, int64_t n, uint8_t *wtr1, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(diag && wtr1 && n >= 0, "pb200_interpreted_layer: bad argument");
    if (n == 0) return 0;
    interpreted_layer_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(diag, wtr1, n);
    LEAVE();
}

extern "C" int pb200_binary_representation(pb200_ctx *ctx, const uint16_t *diag, int64_t n, uint16_t *out,
                                           void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(diag && out && n >= 0, "pb200_binary_representation: bad argument");
    if (n == 0) return 0;
    binary_representation_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(diag, out, n);
    LEAVE();
}

static int check_mode(int mode) {
    if (mode != PB200_ADJ_MASK && mode != PB200_ADJ_IGNORE && mode != PB200_ADJ_COVER)
        return fail(PB200_E_BAD_MODE, "ERROR mask adjacent to cloud/cloud-shadow mode: %d", mode);
    return 0;
}

extern "C" int pb200_preliminary_cloud(pb200_ctx *ctx, const uint8_t *fmask, int mode, int64_t n, uint8_t *cloud,
                                       void *stream) {
    ENTER(ctx);
    int rc = check_mode(mode);
    if (rc) return rc;
    EMPTY_OK(n);
    REQUIRE(fmask && cloud && n >= 0, "pb200_preliminary_cloud: bad argument");
    if (n == 0) return 0;
    preliminary_cloud_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(fmask, cloud, mode == PB200_ADJ_MASK ? 0 : 1, n);
    LEAVE();
}

extern "C" int pb200_aerosol_remap(pb200_ctx *ctx, uint8_t *wtr1, const int16_t *nir, uint8_t *cloud,
                                   const uint8_t *fmask, const uint8_t bits[256], int64_t n, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr1 && nir && cloud && fmask && bits && n >= 0, "pb200_aerosol_remap: bad argument");
    if (n == 0) return 0;
    AerosolBits ab;
    for (int i = 0; i < 64; ++i)
        ab.w[i] = (uint32_t)(bits[4 * i] & 0x1D) | ((uint32_t)(bits[4 * i + 1] & 0x1D) << 8) |
                  ((uint32_t)(bits[4 * i + 2] & 0x1D) << 16) | ((uint32_t)(bits[4 * i + 3] & 0x1D) << 24);
    aerosol_remap_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr1, nir, cloud, fmask, ab, n);
    LEAVE();
}

extern "C" int pb200_landcover_shadow_masks(pb200_ctx *ctx, const uint8_t *wtr1, const int16_t *nir,
                                            const uint8_t *land, const uint8_t *shad, double lcmask_nir, int64_t n,
                                            uint8_t *wtr2, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr1 && wtr2 && n >= 0, "pb200_landcover_shadow_masks: bad argument");
    REQUIRE(nir || !land, "pb200_landcover_shadow_masks: nir is required with a land-cover raster");
    if (n == 0) return 0;
    landcover_shadow_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr1, nir, land, shad, gt_threshold(lcmask_nir),
                                                                    wtr2, n);
    LEAVE();
}

extern "C" int pb200_snow_to_cloud(pb200_ctx *ctx, const uint8_t *wtr2, uint8_t *cloud, const uint8_t *fmask,
                                   int mode, int64_t n, void *stream) {
    ENTER(ctx);
    int rc = check_mode(mode);
    if (rc) return rc;
    EMPTY_OK(n);
    if (mode == PB200_ADJ_COVER)
        return fail(PB200_E_UNSUPPORTED, "pb200_snow_to_cloud: 'cover' dilation (D:2055-2078) not implemented yet");
    REQUIRE(wtr2 && cloud && fmask && n >= 0, "pb200_snow_to_cloud: bad argument");
    if (n == 0) return 0;
    snow_to_cloud_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr2, cloud, fmask, n);
    LEAVE();
}

// `iterations` masked dilation steps from `a` (input) ping-ponging with `b`, MD_HALO steps per launch; returns the
// buffer that holds the result (never the input when iterations >= 1)
static uint8_t *run_masked_dilation(const uint8_t *a, uint8_t *b, uint8_t *c, const uint8_t *mask, int rows, int cols,
                                    int iterations, cudaStream_t st) {
    dim3 grid((cols + MD_TW - 1) / MD_TW, (rows + MD_TH - 1) / MD_TH);
    const uint8_t *src = a;
    uint8_t *dst = b;
    while (iterations > 0) {
        const int k = std::min(iterations, MD_HALO);
        masked_dilation_tiled_kernel<<<grid, 256, 0, st>>>(src, mask, dst, rows, cols, k);
        iterations -= k;
        src = dst;
        dst = (dst == b) ? c : b;
    }
    return const_cast<uint8_t *>(src);
}

extern "C" int pb200_masked_dilation(pb200_ctx *ctx, const uint8_t *in, const uint8_t *mask, int rows, int cols,
                                     int iterations, uint8_t *out, uint8_t *scratch, void *stream) {
    ENTER(ctx);
    REQUIRE(rows >= 0 && cols >= 0 && iterations >= 1, "pb200_masked_dilation: bad size / iterations");
    if ((long long)rows * cols == 0) return 0;
    REQUIRE(in && mask && out && scratch && in != out && scratch != out && scratch != in,
            "pb200_masked_dilation: bad argument");
    const size_t n = (size_t)rows * cols;
    // arrange the ping-pong so that the last launch writes `out`
    const int launches = (iterations + MD_HALO - 1) / MD_HALO;
    uint8_t *first = (launches % 2) ? out : scratch, *second = (launches % 2) ? scratch : out;
    uint8_t *res = run_masked_dilation(in, first, second, mask, rows, cols, iterations, st);
    if (res != out) CK(cudaMemcpyAsync(out, res, n, cudaMemcpyDeviceToDevice, st));
    LEAVE();
}

extern "C" int pb200_snow_to_cloud_cover(pb200_ctx *ctx, const uint8_t *wtr2, uint8_t *cloud, const uint8_t *fmask,
                                         int rows, int cols, uint8_t *scratch, void *stream) {
    ENTER(ctx);
    REQUIRE(rows >= 0 && cols >= 0, "pb200_snow_to_cloud_cover: bad size");
    const long long n = (long long)rows * cols;
    if (n == 0) return 0;
    REQUIRE(wtr2 && cloud && fmask && scratch, "pb200_snow_to_cloud_cover: bad argument");
    uint8_t *A = scratch, *B = scratch + n, *M = scratch + 2 * n, *S = scratch + 3 * n;
    const int g = grid_for(ctx, n, 256);
    cover_init_kernel<<<g, 256, 0, st>>>(fmask, cloud, A, M, n);                      // snow, area
    uint8_t *snow = run_masked_dilation(A, S, B, M, rows, cols, 10, st);               // D:2060 (one launch -> S)
    cover_mid_kernel<<<g, 256, 0, st>>>(snow, cloud, wtr2, M, A, n);                   // area2, not_masked
    uint8_t *nm = run_masked_dilation(A, B, nullptr, M, rows, cols, 7, st);            // D:2075 (one launch -> B)
    cover_final_kernel<<<g, 256, 0, st>>>(snow, nm, wtr2, cloud, n);
    LEAVE();
}

extern "C" int pb200_cover_tail(pb200_ctx *ctx, uint8_t *wtr2, const uint8_t *cloud, int64_t n, uint8_t *wtr,
                                uint8_t *bwtr, uint8_t *conf, uint8_t *wtr1, uint8_t *wtr1_remapped, int collapse,
                                void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr2 && cloud && n >= 0, "pb200_cover_tail: bad argument");
    if (n == 0) return 0;
    cover_tail_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr2, cloud, wtr, bwtr, conf, wtr1, wtr1_remapped, collapse, n);
    LEAVE();
}

extern "C" int pb200_cloud_masking(pb200_ctx *ctx, const uint8_t *wtr2, const uint8_t *cloud, int64_t n,
                                   uint8_t *wtr, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr2 && cloud && wtr && n >= 0, "pb200_cloud_masking: bad argument");
    if (n == 0) return 0;
    cloud_masking_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr2, cloud, wtr, n);
    LEAVE();
}
extern "C" int pb200_binary_water(pb200_ctx *ctx, const uint8_t *wtr, int64_t n, uint8_t *bwtr, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr && bwtr && n >= 0, "pb200_binary_water: bad argument");
    if (n == 0) return 0;
    binary_water_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr, bwtr, n);
    LEAVE();
}
extern "C" int pb200_confidence(pb200_ctx *ctx, const uint8_t *wtr2, const uint8_t *cloud, int64_t n, uint8_t *conf,
                                void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(wtr2 && cloud && conf && n >= 0, "pb200_confidence: bad argument");
    if (n == 0) return 0;
    confidence_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(wtr2, cloud, conf, n);
    LEAVE();
}
extern "C" int pb200_collapse(pb200_ctx *ctx, const uint8_t *layer, int64_t n, uint8_t *out, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(layer && out && n >= 0, "pb200_collapse: bad argument");
    if (n == 0) return 0;
    collapse_kernel<<<grid_for(ctx, n, 256), 256, 0, st>>>(layer, out, n);
    LEAVE();
}

extern "C" int pb200_browse_table(int collapse_wtr_classes, int exclude_psw_aggressive, int set_not_water_to_nodata,
                                  int set_cloud_to_nodata, int set_snow_to_nodata, int set_ocean_masked_to_nodata,
                                  uint8_t table[256]) {
    if (!table) return fail(PB200_E_INVALID_ARG, "pb200_browse_table: null table");
    for (int v = 0; v < 256; ++v) {
        int x = v;
        if (exclude_psw_aggressive && x == 4) x = 0;                 // D:3107-3110
        if (collapse_wtr_classes) {                                   // D:3112-3113, table D:201-213
            switch (x) {
                case 0: case 1: case 252: case 253: case 254: case 255: break;
                case 2: x = 1; break;
                case 3: case 4: x = 2; break;
                default: x = 255;
            }
        }
        if (set_not_water_to_nodata && x == 0) x = 255;               // D:3115-3116
        if (set_cloud_to_nodata && x == 253) x = 255;                 // D:3118-3119
        if (set_snow_to_nodata && x == 252) x = 255;                  // D:3121-3122
        if (set_ocean_masked_to_nodata && x == 254) x = 255;          // D:3124-3125
        table[v] = (uint8_t)x;
    }
    return 0;
}

extern "C" int pb200_byte_table(pb200_ctx *ctx, const uint8_t *in, int64_t n, const uint8_t table[256], uint8_t *out,
                                void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(in && out && table && n >= 0, "pb200_byte_table: bad argument");
    if (n == 0) return 0;
    ByteTable T;
    std::memcpy(T.v, table, 256);
    byte_table_kernel<<<grid_for(ctx, (n + 15) / 16, 256), 256, 0, st>>>(in, out, n, T);
    LEAVE();
}

extern "C" int pb200_scale_offset(pb200_ctx *ctx, const int16_t *band, int64_t n, double scale, double offset,
                                  const uint8_t *invalid, float *out, void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(band && out && n >= 0, "pb200_scale_offset: bad argument");
    if (n == 0) return 0;
    // Python-float scalars are "weak" in numpy: both are rounded to float32 before the float32 array operation
    scale_offset_kernel<<<grid_for(ctx, (n + 7) / 8, 256), 256, 0, st>>>(band, invalid, out, n, (float)scale, (float)offset);
    LEAVE();
}

extern "C" int pb200_histogram_u8(pb200_ctx *ctx, const uint8_t *image, int64_t n, unsigned long long *counts,
                                  void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(image && counts && n >= 0, "pb200_histogram_u8: bad argument");
    if (n == 0) return 0;
    // one CTA per SM (128 KB of lane-private counters each); a CTA counts at most 2^32 - 1 pixels per bin
    CK(cudaFuncSetAttribute(histogram_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HIST_SMEM_BYTES));
    const long long work = (n + 16 * 32 * HIST_WARPS - 1) / (16 * 32 * HIST_WARPS);
    const int grid = (int)std::max<long long>(1, std::min<long long>(work, ctx->sm_count));
    histogram_u8_kernel<<<grid, 32 * HIST_WARPS, HIST_SMEM_BYTES, st>>>(image, n, counts);
    LEAVE();
}

// _compute_otsu_threshold (D:1638-1684) from exact per-value counts, every float64 operation in numpy's order.
extern "C" int pb200_otsu_threshold(const unsigned long long counts[256], int is_normalized, double *threshold) {
    if (!counts || !threshold) return fail(PB200_E_INVALID_ARG, "pb200_otsu_threshold: null argument");
    int vmin = -1, vmax = -1;
    for (int v = 0; v < 256; ++v)
        if (counts[v]) { if (vmin < 0) vmin = v; vmax = v; }
    if (vmin < 0) return fail(PB200_E_INVALID_ARG, "pb200_otsu_threshold: empty image (np.histogram raises ValueError)");
    constexpr int NB = 256;
    // np.histogram(image, bins=256): outer edges (min, max), widened by 0.5 when equal; edges = linspace(first, last, 257)
    double first = (double)vmin, last = (double)vmax;
    if (vmin == vmax) { first -= 0.5; last += 0.5; }
    double edges[NB + 1];
    const double step = (last - first) / (double)NB;
    for (int i = 0; i <= NB; ++i) edges[i] = (double)i * step + first;
    edges[NB] = last;
    // uniform-bin fast path of np.histogram: scaled index, then the two edge corrections
    long long hist_i[NB] = {};
    const double norm_denom = last - first;
    for (int v = vmin; v <= vmax; ++v) {
        if (!counts[v]) continue;
        const double a = (double)v;
        const double f = ((a - first) / norm_denom) * (double)NB;
        long long idx = (long long)f;
        if (idx == NB) idx -= 1;
        if (a < edges[idx]) idx -= 1;
        if (a >= edges[idx + 1] && idx != NB - 1) idx += 1;
        hist_i[idx] += (long long)counts[v];
    }
    double hist[NB], mids[NB];
    long long hmax = 0;
    for (int i = 0; i < NB; ++i) hmax = std::max(hmax, hist_i[i]);
    for (int i = 0; i < NB; ++i) {
        hist[i] = is_normalized ? (double)hist_i[i] / (double)hmax : (double)hist_i[i];      // D:1667
        mids[i] = (edges[i] + edges[i + 1]) / 2.0;                                            // D:1670
    }
    // cumulative sums, forward and backward (np.cumsum is a sequential accumulation)
    double w1[NB], w2[NB], m1[NB], m2[NB];
    long long w1_i[NB], w2_i[NB];
    {
        double acc = 0.0, accm = 0.0; long long acci = 0;
        for (int i = 0; i < NB; ++i) {
            acc += hist[i]; acci += hist_i[i]; accm += hist[i] * mids[i];
            w1[i] = acc; w1_i[i] = acci; m1[i] = accm / acc;                                  // D:1673, 1677
        }
        acc = 0.0; accm = 0.0; acci = 0;
        for (int i = NB - 1; i >= 0; --i) {
            acc += hist[i]; acci += hist_i[i]; accm += hist[i] * mids[i];
            w2[i] = acc; w2_i[i] = acci; m2[i] = accm / acc;                                  // D:1674, 1679
        }
    }
    // D:1681-1684: first maximum of the inter-class variance; np.argmax returns the first NaN if there is one
    int best = 0;
    double best_v = 0.0;
    for (int i = 0; i < NB - 1; ++i) {
        const double d = m1[i] - m2[i + 1];
        // not normalised: weight1 * weight2 is an int64 product in numpy before it meets the float64 factor
        const double ww = is_normalized ? w1[i] * w2[i + 1] : (double)(w1_i[i] * w2_i[i + 1]);
        const double v = ww * (d * d);
        if (std::isnan(v)) { best = i; break; }
        if (i == 0 || v > best_v) { best = i; best_v = v; }
    }
    *threshold = mids[best];
    return 0;
}

extern "C" int pb200_greater_than_u8(pb200_ctx *ctx, const uint8_t *image, int64_t n, double threshold, uint8_t *out,
                                     void *stream) {
    ENTER(ctx);
    EMPTY_OK(n);
    REQUIRE(image && out && n >= 0, "pb200_greater_than_u8: bad argument");
    REQUIRE(threshold == threshold, "pb200_greater_than_u8: NaN threshold");
    if (n == 0) return 0;
    // x > t for integer x: x >= floor(t) + 1
    const double lim = std::floor(threshold) + 1.0;
    const int limit = lim < 0.0 ? 0 : (lim > 256.0 ? 256 : (int)lim);
    greater_than_u8_kernel<<<grid_for(ctx, (n + 15) / 16, 256), 256, 0, st>>>(image, out, n, limit);
    LEAVE();
}

extern "C" int pb200_shadow(pb200_ctx *ctx, const float *dem, int rows, int cols, double sun_azimuth,
                            double sun_elevation, const double *terms, const pb200_params *params, uint8_t *out,
                            void *stream) {
    ENTER(ctx);
    REQUIRE(dem && out && params, "pb200_shadow: null argument");
    // np.gradient needs at least 2 samples along every axis (ValueError otherwise)
    REQUIRE(rows >= 2 && cols >= 2, "pb200_shadow: the DEM must be at least 2 x 2 (np.gradient, D:4255)");
    DevParams P;
    std::memset(&P, 0, sizeof(P));
    P.dxf = (float)params->pixel_spacing_x;
    P.dyf = -std::fabs((float)params->pixel_spacing_y);
    P.flags = params->numpy1_promotion ? PF_NUMPY1 : 0u;
    pb200_angle_thresholds(params, &P.cos_thr, &P.tan_thr);
    pb200_tile t;
    std::memset(&t, 0, sizeof(t));
    t.sun_azimuth = sun_azimuth;
    t.sun_elevation = sun_elevation;
    t.sun_terms[0] = std::numeric_limits<double>::quiet_NaN();
    if (terms)
        for (int i = 0; i < 5; ++i) t.sun_terms[i] = terms[i];
    TileDev d;
    sun_terms(t, &d);
    SunTerms S{d.sx, d.sy, d.sz, d.sin_az, d.cos_az};
    dim3 block(64, 4), grid((cols + 63) / 64, (rows + 3) / 4);
    shadow_kernel<<<grid, block, 0, st>>>(dem, rows, cols, out, S, P);
    LEAVE();
}

extern "C" int pb200_shadow_f64(pb200_ctx *ctx, const double *dem, int rows, int cols, double sun_azimuth,
                                double sun_elevation, const double *terms, const pb200_params *params, uint8_t *out,
                                void *stream) {
    ENTER(ctx);
    REQUIRE(dem && out && params, "pb200_shadow_f64: null argument");
    REQUIRE(rows >= 2 && cols >= 2, "pb200_shadow_f64: the DEM must be at least 2 x 2 (np.gradient, D:4255)");
    double cos_thr, tan_thr;
    pb200_angle_thresholds(params, &cos_thr, &tan_thr);
    pb200_tile t;
    std::memset(&t, 0, sizeof(t));
    t.sun_azimuth = sun_azimuth;
    t.sun_elevation = sun_elevation;
    t.sun_terms[0] = std::numeric_limits<double>::quiet_NaN();
    if (terms)
        for (int i = 0; i < 5; ++i) t.sun_terms[i] = terms[i];
    TileDev d;
    sun_terms(t, &d);
    SunTerms S{d.sx, d.sy, d.sz, d.sin_az, d.cos_az};
    dim3 block(64, 4), grid((cols + 63) / 64, (rows + 3) / 4);
    shadow_f64_kernel<<<grid, block, 0, st>>>(dem, rows, cols, out, S, params->pixel_spacing_x, -std::fabs(params->pixel_spacing_y),
                                              cos_thr, tan_thr);
    LEAVE();
}

extern "C" int pb200_landcover_aggregate(pb200_ctx *ctx, const uint8_t *worldcover, const uint8_t *copernicus, int rows,
                                         int cols, const uint8_t forest[256], int year_offset,
                                         const int32_t thresholds[4], uint8_t *land, void *stream) {
    ENTER(ctx);
    REQUIRE(rows >= 0 && cols >= 0, "pb200_landcover_aggregate: bad size");
    if ((long long)rows * cols == 0) return 0;
    REQUIRE(worldcover && copernicus && forest && thresholds && land, "pb200_landcover_aggregate: null argument");
    REQUIRE(year_offset >= 0 && year_offset <= 99, "pb200_landcover_aggregate: year offset %d outside 0..99 (D:253-257)",
            year_offset);
    LandParams L;
    L.thr_tree = thresholds[0]; L.thr_low = thresholds[1]; L.thr_high = thresholds[2]; L.thr_water = thresholds[3];
    L.cls_tree = 201u; L.cls_low = (uint32_t)year_offset; L.cls_high = 100u + (uint32_t)year_offset; L.cls_water = 200u;
    for (int i = 0; i < 8; ++i) L.forest_bits[i] = 0u;
    for (int v = 0; v < 256; ++v)
        if (forest[v]) L.forest_bits[v >> 5] |= 1u << (v & 31);
    const bool vec = (cols % 4) == 0 && aligned(worldcover, 4) && aligned(copernicus, 4) && aligned(land, 4);
    dim3 block(32, 8), grid(((cols + 3) / 4 + 31) / 32, (rows + 7) / 8);
    grid.y = (unsigned)std::max(1, std::min((int)grid.y, (ctx->sm_count * 8 + (int)grid.x - 1) / (int)grid.x));
    if (vec) landcover_aggregate_kernel<true><<<grid, block, 0, st>>>(worldcover, copernicus, land, rows, cols, L);
    else landcover_aggregate_kernel<false><<<grid, block, 0, st>>>(worldcover, copernicus, land, rows, cols, L);
    LEAVE();
}

extern "C" int pb200_hillshade(pb200_ctx *ctx, const float *dem, int rows, int cols, double sun_azimuth, double sun_elevation,
                               double ewres, double nsres, uint8_t *out, unsigned long long *counts, void *stream) {
    ENTER(ctx);
    REQUIRE(rows >= 0 && cols >= 0, "pb200_hillshade: bad size");
    if ((long long)rows * cols == 0) return 0;
    REQUIRE(dem && out, "pb200_hillshade: null argument");
    REQUIRE(ewres != 0.0 && nsres != 0.0, "pb200_hillshade: zero pixel spacing");
    // GDALCreateHillshadeData with gdaldem's defaults: z = 1, scale = 1, Horn (divisor 8)
    const double deg2rad = 3.14159265358979323846 / 180.0;
    const double z_scaled = 1.0 / 8.0;
    HillParams H;
    H.inv_ewres = 1.0 / ewres;
    H.inv_nsres = 1.0 / nsres;
    const double cos_alt_z = std::cos(sun_elevation * deg2rad) * z_scaled;
    H.sin_alt_254 = 254.0 * std::sin(sun_elevation * deg2rad);
    H.cos_az_cos_alt_z_254 = 254.0 * (std::cos(sun_azimuth * deg2rad) * cos_alt_z);
    H.sin_az_cos_alt_z_254 = 254.0 * (std::sin(sun_azimuth * deg2rad) * cos_alt_z);
    H.square_z = z_scaled * z_scaled;
    dim3 grid((cols + TW - 1) / TW, (rows + TH - 1) / TH);
    alignas(64) CUtensorMap map;
    std::memset(&map, 0, sizeof(map));
    bool tma = aligned(dem, 16) && (cols % 4) == 0 && cols >= SMW && rows >= SMH;
    if (tma) {
        const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        const cuuint64_t gstr[1] = {(cuuint64_t)cols * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)SMW, (cuuint32_t)SMH};
        const cuuint32_t estr[2] = {1, 1};
        tma = ctx->encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)dem, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (tma) hillshade_hist_kernel<true><<<grid, NTHREADS, 0, st>>>(dem, map, rows, cols, out, counts, H);
    else hillshade_hist_kernel<false><<<grid, NTHREADS, 0, st>>>(dem, map, rows, cols, out, counts, H);
    LEAVE();
}

extern "C" int pb200_shadow_sweep(pb200_ctx *ctx, const pb200_params *params, double sun_azimuth, double sun_elevation,
                                  const double *terms, int mode, uint64_t seed, uint64_t n_samples, uint64_t counts[8]) {
    REQUIRE(ctx && params && counts, "pb200_shadow_sweep: null argument");
    REQUIRE(mode >= 0 && mode <= 3, "pb200_shadow_sweep: mode 0..3");
    CK(cudaSetDevice(ctx->device));
    DevParams P;
    pb200_params p = *params;
    if (p.adjacent_mode == PB200_ADJ_COVER) p.adjacent_mode = PB200_ADJ_MASK;
    int rc = derive_params(&p, &P, true);
    if (rc) return rc;
    FastParams F;
    build_fast_params(&p, P, &F);
    pb200_tile t;
    std::memset(&t, 0, sizeof(t));
    t.sun_azimuth = sun_azimuth;
    t.sun_elevation = sun_elevation;
    t.sun_terms[0] = std::numeric_limits<double>::quiet_NaN();
    if (terms)
        for (int i = 0; i < 5; ++i) t.sun_terms[i] = terms[i];
    TileDev d;
    std::memset(&d, 0, sizeof(d));
    sun_terms(t, &d);
    SweepCounts *dc = nullptr;
    CK(cudaMalloc((void **)&dc, sizeof(SweepCounts)));
    CK(cudaMemset(dc, 0, sizeof(SweepCounts)));
    const int blocks = ctx->sm_count * 8, threads = 256;
    const unsigned long long per_thread = (n_samples + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
    shadow_sweep_kernel<<<blocks, threads>>>(P, F, d, mode, seed, per_thread, dc);
    static_assert(sizeof(SweepCounts) == 8 * sizeof(uint64_t), "counts[8]");
    cudaError_t e = cudaMemcpy(counts, dc, sizeof(SweepCounts), cudaMemcpyDeviceToHost);
    cudaFree(dc);
    if (e != cudaSuccess) return fail_cuda(e, "pb200_shadow_sweep");
    return 0;
}

extern "C" int pb200_fast8_sweep(pb200_ctx *ctx, const pb200_params *params, uint64_t counts[6]) {
    REQUIRE(ctx && params && counts, "pb200_fast8_sweep: null argument");
    CK(cudaSetDevice(ctx->device));
    DevParams P;
    pb200_params p = *params;
    if (p.adjacent_mode == PB200_ADJ_COVER) p.adjacent_mode = PB200_ADJ_MASK;
    int rc = derive_params(&p, &P, true);
    if (rc) return rc;
    FastParams F;
    build_fast_params(&p, P, &F);
    if (!F.fast8)
        return fail(PB200_E_UNSUPPORTED, "pb200_fast8_sweep: these parameters do not select the FAST8 kernel variant");
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **)&d, 6 * 8));
    CK(cudaMemset(d, 0, 6 * 8));
    fast8_sweep_kernel<<<32767, 256>>>(F, p.th.wigt, p.th.pswt_1_mndwi, p.th.pswt_2_mndwi, p.th.pswt_1_ndvi, p.th.awgt, d);
    cudaError_t e = cudaMemcpy(counts, d, 6 * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(e, "pb200_fast8_sweep");
    return 0;
}

extern "C" int pb200_ratio_sweep(pb200_ctx *ctx, double t, int is_less, uint64_t *mismatches) {
    REQUIRE(ctx && mismatches, "pb200_ratio_sweep: null argument");
    CK(cudaSetDevice(ctx->device));
    int32_t a, b;
    int rc = pb200_ratio_bound(t, is_less, &a, &b);
    if (rc) return rc;
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **)&d, 8));
    CK(cudaMemset(d, 0, 8));
    ratio_sweep_kernel<<<65536, 256>>>(a, b, t, is_less, d);
    cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail_cuda(e, "pb200_ratio_sweep");
    return 0;
}

// ---------------------------------------------------------------------------
// multi-GPU: NCCL DEM halo exchange of a row-stripped raster
// ---------------------------------------------------------------------------
static int fail_nccl(int r, const char *what) {
    NcclApi *api = nccl_api();
    g_err = std::string(what) + ": NCCL error " + std::to_string(r) + " - " +
            ((api->handle && api->GetErrorString) ? api->GetErrorString(r) : "?");
    return PB200_E_NCCL;
}
#define CKN(call)                                          \
    do {                                                   \
        int r_ = (call);                                   \
        if (r_ != NCCL_SUCCESS) return fail_nccl(r_, #call); \
    } while (0)
static int need_nccl(NcclApi **out) {
    NcclApi *api = nccl_api();
    if (!api->handle) return fail(PB200_E_NCCL, "%s", api->error.c_str());
    *out = api;
    return 0;
}

extern "C" int pb200_comm_unique_id(uint8_t id[PB200_COMM_ID_BYTES]) {
    if (!id) return fail(PB200_E_INVALID_ARG, "pb200_comm_unique_id: null output");
    NcclApi *api;
    int rc = need_nccl(&api);
    if (rc) return rc;
    static_assert(sizeof(NcclUniqueId) == PB200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclUniqueId u;
    CKN(api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return 0;
}

extern "C" int pb200_comm_init(pb200_ctx *ctx, const uint8_t id[PB200_COMM_ID_BYTES], int rank, int nranks) {
    if (!ctx || !id) return fail(PB200_E_INVALID_ARG, "pb200_comm_init: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks)
        return fail(PB200_E_INVALID_ARG, "pb200_comm_init: rank %d of %d", rank, nranks);
    if (ctx->comm.comm) return fail(PB200_E_INVALID_ARG, "pb200_comm_init: the context already has a communicator");
    NcclApi *api;
    int rc = need_nccl(&api);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    NcclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    NcclComm c = nullptr;
    CKN(api->CommInitRank(&c, nranks, u, rank));
    ctx->comm.comm = c;
    ctx->comm.rank = rank;
    ctx->comm.nranks = nranks;
    return 0;
}

extern "C" int pb200_comm_destroy(pb200_ctx *ctx) {
    if (!ctx || !ctx->comm.comm) return 0;
    NcclApi *api;
    int rc = need_nccl(&api);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    NcclComm c = ctx->comm.comm;
    ctx->comm = Comm();
    CKN(api->CommDestroy(c));
    return 0;
}

extern "C" int pb200_halo_exchange_dem(pb200_ctx *ctx, float *dem_ext, int n_rows, int pitch, void *stream) {
    if (!ctx || !dem_ext || n_rows < 1 || pitch < 1)
        return fail(PB200_E_INVALID_ARG, "pb200_halo_exchange_dem: bad argument");
    if (!ctx->comm.comm) return fail(PB200_E_INVALID_ARG, "pb200_halo_exchange_dem: pb200_comm_init has not been called");
    NcclApi *api;
    int rc = need_nccl(&api);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int rank = ctx->comm.rank, n = ctx->comm.nranks;
    if (n == 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    NcclComm c = ctx->comm.comm;
    const size_t cnt = (size_t)pitch;
    float *first = dem_ext + (size_t)pitch, *last = dem_ext + (size_t)n_rows * pitch;   // own rows 1 and n_rows
    float *above = dem_ext, *below = dem_ext + (size_t)(n_rows + 1) * pitch;            // halo rows 0 and n_rows + 1
    CKN(api->GroupStart());
    int r = NCCL_SUCCESS;
    if (rank > 0) {
        if (r == NCCL_SUCCESS) r = api->Send(first, cnt, NCCL_FLOAT32, rank - 1, c, st);
        if (r == NCCL_SUCCESS) r = api->Recv(above, cnt, NCCL_FLOAT32, rank - 1, c, st);
    }
    if (rank < n - 1) {
        if (r == NCCL_SUCCESS) r = api->Send(last, cnt, NCCL_FLOAT32, rank + 1, c, st);
        if (r == NCCL_SUCCESS) r = api->Recv(below, cnt, NCCL_FLOAT32, rank + 1, c, st);
    }
    const int re = api->GroupEnd();
    if (r != NCCL_SUCCESS) return fail_nccl(r, "ncclSend / ncclRecv");
    if (re != NCCL_SUCCESS) return fail_nccl(re, "ncclGroupEnd");
    return 0;
}

extern "C" int pb200_comm_allreduce_u64(pb200_ctx *ctx, uint64_t *values, int n, void *stream) {
    if (!ctx || !values || n < 1) return fail(PB200_E_INVALID_ARG, "pb200_comm_allreduce_u64: bad argument");
    if (!ctx->comm.comm) return fail(PB200_E_INVALID_ARG, "pb200_comm_allreduce_u64: pb200_comm_init has not been called");
    NcclApi *api;
    int rc = need_nccl(&api);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm.nranks == 1) return 0;
    CKN(api->AllReduce(values, values, (size_t)n, NCCL_UINT64, NCCL_SUM, ctx->comm.comm, (cudaStream_t)stream));
    return 0;
}
