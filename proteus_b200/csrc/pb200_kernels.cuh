// pb200_kernels.cuh - sm_100a kernels of the DSWx-HLS classification path.
//
//  dswx_fused_kernel<VEC>   SURVEY K1: D:2203-2209, 2298-2299, 5088-5111,
//                           5161-5171, 5225-5286, 5358, 5368, 2688-2689 in
//                           one pass over a batch of tiles.
//  function-granular kernels: one per reference function (parity tests and
//                           the rebindable drop-in surface).
#pragma once
#include <cuda.h>
#include "pb200_device.cuh"

namespace pb200 {

constexpr int TW = 128;            // pixels per CTA tile row  (32 lanes x 4 px)
constexpr int TH = 32;             // rows per CTA tile        (8 warps x 4 rows)
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int ROWS_PER_WARP = TH / NWARPS;
// The TMA box must START on a 16-byte boundary of the DEM row (measured on B200:
// UTMALDG traps with "illegal instruction" when the inner coordinate is not a
// multiple of 4 floats, scripts/tma_probe.cu).  The tile's first DEM column is
// dem_off_x + x0 with x0 % 128 == 0 and dem_off_x = 50 in production, so the box
// starts DEM_PADX + (dem_off_x & 3) columns to the left of the tile: 4..7 columns.
constexpr int DEM_PADX = 4;
constexpr int SMW = TW + 2 * DEM_PADX;   // 136 floats = 544 B >= 7 + 128 + 1; multiple of 16 B (TMA box rule)
constexpr int SMH = TH + 2;
constexpr int N_CNT = 12;

struct __align__(128) FusedSmem {
    float dem[SMH][SMW];                 // 18 496 B, TMA destination
    uint32_t out_lut[128];
    uint32_t diag_lut[32];
    uint16_t fmask_lut[256];
    uint8_t  bin_lut[128];
    unsigned long long mbar;
    unsigned int cnt[N_CNT];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- TMA + mbarrier (PTX; SASS: UTMALDG / SYNCS) ---------------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Polls of a waiting warp take issue slots from the working warps of its SM sub-partition (profiles/: 12 of 125
// executed thread-instructions per pixel were YIELD / TRYWAIT / BRA of waiting consumers): the optional
// suspend-time hint lets the hardware park the thread until the phase completes or the time is over.
#ifndef PB200_MBAR_HINT_NS
#define PB200_MBAR_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
#if PB200_MBAR_HINT_NS > 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_addr),
        "r"(parity), "r"((uint32_t)PB200_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
// the box load alone; the caller has acquired the tensor map (tma_acquire_map) since it was written
__device__ __forceinline__ void tma_load_2d_acquired(void *dst, const CUtensorMap *map, int c0, int c1,
                                                     unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_acquire_map(const CUtensorMap *map) {
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
}
// wait of a thread with nothing else to do (a producer): sleep between polls instead of spinning - a spinning warp
// competes for the issue slots of the working warps of its SM sub-partition (measured: 15 of 126 executed
// thread-instructions per pixel were the producer's polls)
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, uint32_t parity, unsigned ns) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done) __nanosleep(ns);
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1,
                                            unsigned long long *bar) {
    // The tensor maps live in GLOBAL memory (one per tile of the batch, written
    // by the host with cudaMemcpy): the tensormap proxy must acquire them before
    // first use in every CTA (CUDA programming guide, "tensor map in global
    // memory"); without the fence UTMALDG traps with an illegal instruction.
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// K1: fused classification.  grid = (max CTAs per tile, n_tiles), block = 256.
// Each lane owns 4 consecutive pixels of a row; a warp owns 4 rows of the
// 128 x 32 CTA tile.  Band values arrive as one 8-byte load per band (VEC) -
// row pitch 3660 * 2 B is 8-B but not 16-B aligned, so 8 B is the widest load
// that stays aligned on every row - and the DEM tile with its halo as one TMA
// 2-D box into shared memory.
// ---------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(NTHREADS, 3)
dswx_fused_kernel(const TileDev *__restrict__ tiles, const CUtensorMap *__restrict__ tmaps,
                  const __grid_constant__ DevParams P) {
    __shared__ FusedSmem s;
    const TileDev &T = tiles[blockIdx.y];
    if ((int)blockIdx.x >= T.n_ctas) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = T.width, H = T.height;
    const int tile_x = blockIdx.x % T.tiles_x, tile_y = blockIdx.x / T.tiles_x;
    const int x0 = tile_x * TW, y0 = tile_y * TH;
    const bool has_dem = T.dem != nullptr;
    const bool has_land = T.land != nullptr;
    const bool has_ocean = T.ocean != nullptr;

    // ---- stage the DEM tile (+halo) ---------------------------------------
    const int padx = DEM_PADX + (T.dem_off_x & 3);
    const int dem_x0 = T.dem_off_x + x0 - padx;          // multiple of 4 -> 16-B aligned box start
    const int dem_y0 = T.dem_off_y + y0 - 1;
    const bool use_tma = has_dem && (T.flags & TF_TMA);
    if (use_tma && tid == 0) {
        mbar_init(&s.mbar, 1);
        mbar_expect_tx(&s.mbar, (uint32_t)sizeof(s.dem));
        tma_load_2d(&s.dem[0][0], &tmaps[blockIdx.y], dem_x0, dem_y0, &s.mbar);
    }
    if (has_dem && !use_tma) {
        const float *__restrict__ dem = T.dem;
        for (int i = tid; i < SMH * SMW; i += NTHREADS) {
            const int r = i / SMW, c = i - r * SMW;
            const int gy = dem_y0 + r, gx = dem_x0 + c;
            float v = 0.0f;
            if (gy >= 0 && gy < T.dem_rows && gx >= 0 && gx < T.dem_pitch)
                v = __ldg(dem + (size_t)gy * T.dem_pitch + gx);
            s.dem[r][c] = v;
        }
    }
    // ---- tables into shared memory ----------------------------------------
    if (tid < 128) {
        s.out_lut[tid] = P.out_lut[tid];
        s.fmask_lut[tid] = P.fmask_lut[tid];
        s.fmask_lut[tid + 128] = P.fmask_lut[tid + 128];
    } else if (tid < 160) {
        s.diag_lut[tid - 128] = P.diag_lut[tid - 128];
    } else if (tid < 160 + N_CNT) {
        s.cnt[tid - 160] = 0u;
    }
    if (P.flags & PF_HISTOGRAM) {
        // bin of the UNCOLLAPSED WTR value, recomputed from (k2, c)
        if (tid < 128) {
            const uint32_t k2 = tid >> 4, c = tid & 15;
            const uint32_t w = (k2 == 7u) ? 255u : cloud_masking(expand_class(k2), c);
            s.bin_lut[tid] = (uint8_t)(w < 5u ? w : w - 247u);     // 252..255 -> 5..8
        }
    }
    __syncthreads();

    const uint32_t cls_lo = P.cls_lut[0] | (P.cls_lut[1] << 8) | (P.cls_lut[2] << 16) | (P.cls_lut[3] << 24);
    const uint32_t cls_hi = P.cls_lut[4] | (P.cls_lut[5] << 8) | (P.cls_lut[6] << 16) | (P.cls_lut[7] << 24);
    const bool aerosol_on = P.flags & PF_AEROSOL;
    const bool histogram = P.flags & PF_HISTOGRAM;
    SunTerms S;
    S.sx = T.sx; S.sy = T.sy; S.sz = T.sz; S.sin_az = T.sin_az; S.cos_az = T.cos_az;

    uint32_t acc_valid = 0, acc_cv = 0, acc_nno = 0;
    unsigned long long acc_hist = 0ull;
    bool dem_ready = !use_tma;

    const int x = x0 + lane * 4;
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        const int ly = warp * ROWS_PER_WARP + rr;
        const int y = y0 + ly;
        if (y >= H) break;                       // warp-uniform
        if (x >= W) continue;                    // whole quad outside (VEC: W % 4 == 0)
        const size_t pix = (size_t)y * W + x;
        const int nvalid = VEC ? 4 : min(4, W - x);

        // ---- loads ---------------------------------------------------------
        int b[6][4];
        uint32_t fm[4], ld[4], oc[4];
        if (VEC) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int2 v = ldg_stream_v2(T.band[k] + pix);
                b[k][0] = (int)(short)(v.x & 0xffff);
                b[k][1] = v.x >> 16;
                b[k][2] = (int)(short)(v.y & 0xffff);
                b[k][3] = v.y >> 16;
            }
            const uint32_t f4 = ldg_stream_u32(T.fmask + pix);
            const uint32_t l4 = has_land ? ldg_stream_u32(T.land + pix) : 0xffffffffu;
            const uint32_t o4 = has_ocean ? ldg_stream_u32(T.ocean + pix) : 0x01010101u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                fm[j] = (f4 >> (8 * j)) & 255u;
                ld[j] = (l4 >> (8 * j)) & 255u;
                oc[j] = (o4 >> (8 * j)) & 255u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool in = j < nvalid;
#pragma unroll
                for (int k = 0; k < 6; ++k) b[k][j] = in ? (int)__ldg(T.band[k] + pix + j) : 1;
                fm[j] = in ? (uint32_t)__ldg(T.fmask + pix + j) : 0u;
                ld[j] = (in && has_land) ? (uint32_t)__ldg(T.land + pix + j) : 255u;
                oc[j] = (in && has_ocean) ? (uint32_t)__ldg(T.ocean + pix + j) : 1u;
            }
        }

        // ---- terrain shadow (needs the staged DEM tile) ---------------------
        uint32_t sh[4] = {1u, 1u, 1u, 1u};
        if (has_dem) {
            if (!dem_ready) {
                mbar_wait(&s.mbar, 0);
                dem_ready = true;
            }
            const int sc = lane * 4 + padx;              // smem column of pixel 0
            float m[6], u[4], d[4];
#pragma unroll
            for (int j = 0; j < 6; ++j) m[j] = s.dem[ly + 1][sc - 1 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u[j] = s.dem[ly][sc + j];
                d[j] = s.dem[ly + 2][sc + j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // np.gradient interior: (f[i+1] - f[i-1]) / 2.0 in float32 (D:4255)
                const float g_col = __fmul_rn(__fsub_rn(m[j + 2], m[j]), 0.5f);
                const float g_row = __fmul_rn(__fsub_rn(d[j], u[j]), 0.5f);
                sh[j] = shadow_from_gradient(g_col, g_row, P.dxf, P.dyf, S, P.cos_thr, P.tan_thr, (P.flags & PF_NUMPY1) != 0u);
            }
        }

        // ---- per-pixel classification ---------------------------------------
        uint32_t o[4], dg[4], k1[4], k1r[4], k2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t fe = s.fmask_lut[fm[j]];
            bool inv = (fe >> 15) != 0u;                                   // D:2204 (Fmask)
#pragma unroll
            for (int k = 0; k < 6; ++k) inv |= (b[k][j] == P.band_fill[k]);   // D:2204-2207
            const int B = max(b[0][j], 1), G = max(b[1][j], 1), R = max(b[2][j], 1);   // D:2299
            const int N = max(b[3][j], 1), S1 = max(b[4][j], 1), S2 = max(b[5][j], 1);
            const uint32_t d = diagnostic_tests<false>(B, G, R, N, S1, S2, P);
            const uint32_t dl = s.diag_lut[d];
            dg[j] = inv ? 65535u : (dl & 0xffffu);                         // D:5227, 5231
            const bool ocean_masked = has_ocean && oc[j] == 0u;
            const uint32_t ka = inv ? 7u : (ocean_masked ? 6u : (dl >> 16));   // D:5229, 5245, 5249
            k1[j] = ka;

            // coverage counters (D:5104-5111) use the PRELIMINARY cloud layer
            const uint32_t cprelim = fe & 7u;
            const bool valid = !inv && !ocean_masked;
            if (!VEC && j >= nvalid) {
                // lanes past the raster edge contribute nothing
            } else {
                acc_valid += valid;
                acc_cv += (valid && cprelim != 0u);
                acc_nno += oc[j];                      // no shoreline: 1 per pixel (D:5107)
            }

            // aerosol remapping (D:1237-1246)
            const bool remap = aerosol_on && (N <= 1000) && ((fe >> (4u + ka)) & 1u);
            const uint32_t kb = remap ? 1u : ka;
            k1r[j] = kb;
            uint32_t c = cprelim | (remap ? 8u : 0u);

            // land cover + terrain shadow (D:1331-1376) as a class bit mask
            uint32_t km = 0u;
            if (has_dem && sh[j] == 0u && (!has_land || ld[j] != 200u)) km = 0x1Eu;
            if (has_land) {
                if ((ld[j] == 201u || ld[j] < 100u) && N > P.lc_nir) km |= 0x18u;
                if (ld[j] >= 100u && ld[j] < 200u) km |= 0x1Eu;
            }
            const uint32_t kc = ((km >> kb) & 1u) ? 0u : kb;
            k2[j] = kc;

            c += (fe >> 2) & 2u;                                           // D:2081 snow bit
            const uint32_t idx = kc * 16u + c;
            o[j] = s.out_lut[idx];                                         // D:2084-2131, 1727, 1793-1835
            if (histogram && (VEC || j < nvalid)) acc_hist += 1ull << (6u * s.bin_lut[idx]);
        }

        // ---- stores ----------------------------------------------------------
        const uint32_t sel1 = k1[0] | (k1[1] << 4) | (k1[2] << 8) | (k1[3] << 12);
        const uint32_t sel1r = k1r[0] | (k1r[1] << 4) | (k1r[2] << 8) | (k1r[3] << 12);
        const uint32_t sel2 = k2[0] | (k2[1] << 4) | (k2[2] << 8) | (k2[3] << 12);
        const uint32_t w1_4 = __byte_perm(cls_lo, cls_hi, sel1);
        const uint32_t w1r_4 = __byte_perm(cls_lo, cls_hi, sel1r);
        const uint32_t w2_4 = __byte_perm(cls_lo, cls_hi, sel2);
        const uint32_t t01a = __byte_perm(o[0], o[1], 0x5140);   // [o0.b0 o1.b0 o0.b1 o1.b1]
        const uint32_t t23a = __byte_perm(o[2], o[3], 0x5140);
        const uint32_t t01b = __byte_perm(o[0], o[1], 0x7362);   // [o0.b2 o1.b2 o0.b3 o1.b3]
        const uint32_t t23b = __byte_perm(o[2], o[3], 0x7362);
        const uint32_t wtr_4 = __byte_perm(t01a, t23a, 0x5410);
        const uint32_t bwtr_4 = __byte_perm(t01a, t23a, 0x7632);
        const uint32_t conf_4 = __byte_perm(t01b, t23b, 0x5410);
        const uint32_t cloud_4 = __byte_perm(t01b, t23b, 0x7632);
        const uint32_t shad_4 = sh[0] | (sh[1] << 8) | (sh[2] << 16) | (sh[3] << 24);
        const uint32_t diag_lo = dg[0] | (dg[1] << 16), diag_hi = dg[2] | (dg[3] << 16);
        if (VEC) {
            if (T.diag) stg_stream_v2(T.diag + pix, diag_lo, diag_hi);
            if (T.wtr) stg_stream_u32(T.wtr + pix, wtr_4);
            if (T.bwtr) stg_stream_u32(T.bwtr + pix, bwtr_4);
            if (T.conf) stg_stream_u32(T.conf + pix, conf_4);
            if (T.cloud) stg_stream_u32(T.cloud + pix, cloud_4);
            if (T.wtr1) stg_stream_u32(T.wtr1 + pix, w1_4);
            if (T.wtr1r) stg_stream_u32(T.wtr1r + pix, w1r_4);
            if (T.wtr2) stg_stream_u32(T.wtr2 + pix, w2_4);
            if (T.shad) stg_stream_u32(T.shad + pix, shad_4);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nvalid) {
                    if (T.diag) T.diag[pix + j] = (uint16_t)dg[j];
                    if (T.wtr) T.wtr[pix + j] = (uint8_t)(wtr_4 >> (8 * j));
                    if (T.bwtr) T.bwtr[pix + j] = (uint8_t)(bwtr_4 >> (8 * j));
                    if (T.conf) T.conf[pix + j] = (uint8_t)(conf_4 >> (8 * j));
                    if (T.cloud) T.cloud[pix + j] = (uint8_t)(cloud_4 >> (8 * j));
                    if (T.wtr1) T.wtr1[pix + j] = (uint8_t)(w1_4 >> (8 * j));
                    if (T.wtr1r) T.wtr1r[pix + j] = (uint8_t)(w1r_4 >> (8 * j));
                    if (T.wtr2) T.wtr2[pix + j] = (uint8_t)(w2_4 >> (8 * j));
                    if (T.shad) T.shad[pix + j] = (uint8_t)sh[j];
                }
            }
        }
    }

    // ---- counters: warp-aggregated, then one atomic per CTA and slot -------
    if (T.counters) {
        const uint32_t wv = __reduce_add_sync(0xffffffffu, acc_valid);
        const uint32_t wc = __reduce_add_sync(0xffffffffu, acc_cv);
        const uint32_t wn = __reduce_add_sync(0xffffffffu, acc_nno);
        if (lane == 0) {
            if (wv) atomicAdd(&s.cnt[0], wv);
            if (wc) atomicAdd(&s.cnt[1], wc);
            if (wn) atomicAdd(&s.cnt[2], wn);
        }
        if (histogram) {
#pragma unroll
            for (int bin = 0; bin < 9; ++bin) {
                const uint32_t hv = __reduce_add_sync(0xffffffffu, (uint32_t)(acc_hist >> (6 * bin)) & 63u);
                if (lane == 0 && hv) atomicAdd(&s.cnt[3 + bin], hv);
            }
        }
        __syncthreads();
        if (tid < N_CNT && s.cnt[tid]) atomicAdd(&T.counters[tid], (unsigned long long)s.cnt[tid]);
    }
}

// ---------------------------------------------------------------------------
// function-granular kernels (1-D grid-stride; one thread per pixel)
// ---------------------------------------------------------------------------
#define PB200_GRID_STRIDE(i, n) \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

struct BandPtrs { const int16_t *p[6]; };
struct BandOutPtrs { int16_t *p[6]; };

// D:2203-2209, D:2298-2299
__global__ void invalid_and_clip_kernel(BandPtrs raw, const uint8_t *__restrict__ fmask, BandOutPtrs out,
                                        uint8_t *__restrict__ invalid, long long n,
                                        const __grid_constant__ DevParams P) {
    PB200_GRID_STRIDE(i, n) {
        bool inv = fmask ? ((int)fmask[i] == P.fmask_fill) : false;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int v = raw.p[k][i];
            inv |= (v == P.band_fill[k]);
            if (out.p[k]) out.p[k][i] = (int16_t)max(v, 1);
        }
        if (invalid) invalid[i] = inv ? 1 : 0;
    }
}

// D:1840-1916 (inputs used as given: no clip, so d == 0 can occur)
__global__ void diagnostic_tests_kernel(BandPtrs band, uint16_t *__restrict__ diag, long long n,
                                        const __grid_constant__ DevParams P) {
    PB200_GRID_STRIDE(i, n) {
        diag[i] = (uint16_t)diagnostic_tests<true>(band.p[0][i], band.p[1][i], band.p[2][i], band.p[3][i],
                                                   band.p[4][i], band.p[5][i], P);
    }
}

// D:1840-1916 on float32 bands (--offset-and-scale-inputs, D:2300-2302): numpy evaluates every
// operation in float32, left to right, IEEE division, thresholds cast to float32.  Explicit *_rn
// intrinsics (and -fmad=false) keep the operation order and forbid contraction.
struct BandPtrsF { const float *p[6]; };
struct ThresholdsF { float wigt, awgt, p1_mndwi, p1_nir, p1_swir1, p1_ndvi, p2_mndwi, p2_blue, p2_nir, p2_swir1, p2_swir2; };
__global__ void diagnostic_tests_f32_kernel(BandPtrsF band, uint16_t *__restrict__ diag, long long n, ThresholdsF T) {
    PB200_GRID_STRIDE(i, n) {
        const float B = band.p[0][i], G = band.p[1][i], R = band.p[2][i];
        const float N = band.p[3][i], S1 = band.p[4][i], S2 = band.p[5][i];
        const float mndwi = __fdiv_rn(__fsub_rn(G, S1), __fadd_rn(G, S1));                       // D:1872
        const float mbsrv = __fadd_rn(G, R);                                                     // D:1875
        const float mbsrn = __fadd_rn(N, S1);                                                    // D:1878
        const float awesh = __fsub_rn(__fsub_rn(__fadd_rn(B, __fmul_rn(2.5f, G)), __fmul_rn(1.5f, mbsrn)),
                                      __fmul_rn(0.25f, S2));                                     // D:1881
        const float ndvi = __fdiv_rn(__fsub_rn(N, R), __fadd_rn(N, R));                          // D:1884
        uint32_t d = (mndwi > T.wigt) ? 1u : 0u;                                                 // D:1893
        d |= (mbsrv > mbsrn) ? 2u : 0u;                                                          // D:1896
        d |= (awesh > T.awgt) ? 4u : 0u;                                                         // D:1899
        d |= (mndwi > T.p1_mndwi && S1 < T.p1_swir1 && N < T.p1_nir && ndvi < T.p1_ndvi) ? 8u : 0u;          // D:1902
        d |= (mndwi > T.p2_mndwi && B < T.p2_blue && S1 < T.p2_swir1 && S2 < T.p2_swir2 && N < T.p2_nir) ? 16u : 0u;
        diag[i] = (uint16_t)d;
    }
}

// D:1687-1707
__global__ void interpreted_layer_kernel(const uint16_t *__restrict__ diag, uint8_t *__restrict__ wtr1,
                                         long long n) {
    PB200_GRID_STRIDE(i, n) { wtr1[i] = (uint8_t)interpreted_class(diag[i]); }
}

// D:4286-4317
__global__ void binary_representation_kernel(const uint16_t *__restrict__ diag, uint16_t *__restrict__ out,
                                             long long n) {
    PB200_GRID_STRIDE(i, n) { out[i] = (uint16_t)binary_representation(diag[i]); }
}

// D:1919-1993
__global__ void preliminary_cloud_kernel(const uint8_t *__restrict__ fmask, uint8_t *__restrict__ cloud,
                                         int mode, long long n) {
    PB200_GRID_STRIDE(i, n) { cloud[i] = (uint8_t)preliminary_cloud(fmask[i], mode); }
}

// D:1210-1302, in place
struct AerosolBits { uint32_t w[64]; };   // 256 bytes, byte v = class bits of Fmask value v
__global__ void aerosol_remap_kernel(uint8_t *__restrict__ wtr1, const int16_t *__restrict__ nir,
                                     uint8_t *__restrict__ cloud, const uint8_t *__restrict__ fmask,
                                     const __grid_constant__ AerosolBits bits, long long n) {
    __shared__ uint32_t sb[64];
    if (threadIdx.x < 64) sb[threadIdx.x] = bits.w[threadIdx.x];
    __syncthreads();
    PB200_GRID_STRIDE(i, n) {
        const uint32_t w = wtr1[i];
        const uint32_t f = fmask[i];
        const uint32_t cb = (sb[f >> 2] >> (8u * (f & 3u))) & 255u;
        if (w <= 4u && ((cb >> w) & 1u) && (int)nir[i] <= 1000) {
            wtr1[i] = 1;
            const uint32_t c = cloud[i];
            if (c != 255u) cloud[i] = (uint8_t)(c | 8u);
        }
    }
}

// D:1305-1378
__global__ void landcover_shadow_kernel(const uint8_t *__restrict__ wtr1, const int16_t *__restrict__ nir,
                                        const uint8_t *__restrict__ land, const uint8_t *__restrict__ shad,
                                        int lc_nir, uint8_t *__restrict__ wtr2, long long n) {
    PB200_GRID_STRIDE(i, n) {
        // nir is only looked at together with a land-cover raster (D:1354-1362); it may be null without one
        wtr2[i] = (uint8_t)landcover_shadow(wtr1[i], land != nullptr ? (int)nir[i] : 0, land != nullptr, land ? land[i] : 255u,
                                            shad != nullptr, shad ? shad[i] : 1u, lc_nir);
    }
}

// D:1996-2086 (mask / ignore), in place on cloud
__global__ void snow_to_cloud_kernel(const uint8_t *__restrict__ wtr2, uint8_t *__restrict__ cloud,
                                     const uint8_t *__restrict__ fmask, long long n) {
    PB200_GRID_STRIDE(i, n) {
        uint32_t c = cloud[i];
        if (fmask[i] & 16u) c = (c + 2u) & 255u;         // uint8 += 2 wraps (D:2081)
        if (wtr2[i] == 255u) c = 255u;                    // D:2084
        cloud[i] = (uint8_t)c;
    }
}

__global__ void cloud_masking_kernel(const uint8_t *__restrict__ wtr2, const uint8_t *__restrict__ cloud,
                                     uint8_t *__restrict__ wtr, long long n) {
    PB200_GRID_STRIDE(i, n) { wtr[i] = (uint8_t)cloud_masking(wtr2[i], cloud[i]); }
}
__global__ void binary_water_kernel(const uint8_t *__restrict__ wtr, uint8_t *__restrict__ bwtr, long long n) {
    PB200_GRID_STRIDE(i, n) { bwtr[i] = (uint8_t)binary_water(wtr[i]); }
}
__global__ void confidence_kernel(const uint8_t *__restrict__ wtr2, const uint8_t *__restrict__ cloud,
                                  uint8_t *__restrict__ conf, long long n) {
    PB200_GRID_STRIDE(i, n) { conf[i] = (uint8_t)confidence(wtr2[i], cloud[i]); }
}
__global__ void collapse_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n) {
    PB200_GRID_STRIDE(i, n) { out[i] = (uint8_t)collapse_class(in[i]); }
}

// D:3057-3129 (browse relabel) and any other byte -> byte relabel: table by value in the constant bank
struct ByteTable { uint8_t v[256]; };
__global__ void byte_table_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n,
                                  const __grid_constant__ ByteTable T) {
    __shared__ uint8_t lut[256];
    if (threadIdx.x < 256) lut[threadIdx.x] = T.v[threadIdx.x];
    __syncthreads();
    // 16 bytes per thread and iteration: enough bytes in flight to cover the HBM latency with a modest grid
    const long long n16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) ? 0 : n / 16;
    PB200_GRID_STRIDE(i, n16) {
        const int4 x = ldg_stream_v4(in + 16 * i);
        uint32_t w[4] = {(uint32_t)x.x, (uint32_t)x.y, (uint32_t)x.z, (uint32_t)x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = lut[w[k] & 255u] | (lut[(w[k] >> 8) & 255u] << 8) | (lut[(w[k] >> 16) & 255u] << 16) | (lut[w[k] >> 24] << 24);
        stg_stream_v4(out + 16 * i, w[0], w[1], w[2], w[3]);
    }
    for (long long i = 16 * n16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = lut[in[i]];
}

// D:2301-2302, D:3024-3036: float32 scale * (float32(x) - offset), NaN on invalid pixels
__global__ void scale_offset_kernel(const int16_t *__restrict__ in, const uint8_t *__restrict__ invalid,
                                    float *__restrict__ out, long long n, float scale, float offset) {
    // 8 pixels per thread and iteration when the planes allow 16-byte accesses
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                     (invalid == nullptr || (reinterpret_cast<uintptr_t>(invalid) & 7) == 0);
    const long long n8 = vec ? n / 8 : 0;
    PB200_GRID_STRIDE(i, n8) {
        const int4 x = ldg_stream_v4(in + 8 * i);
        const uint32_t w[4] = {(uint32_t)x.x, (uint32_t)x.y, (uint32_t)x.z, (uint32_t)x.w};
        unsigned long long inv = 0ull;
        if (invalid) { const int2 v = ldg_stream_v2(invalid + 8 * i); inv = (uint32_t)v.x | ((unsigned long long)(uint32_t)v.y << 32); }
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int v = (int)(short)((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            r[k] = __fmul_rn(scale, __fsub_rn((float)v, offset));
            if ((inv >> (8 * k)) & 255ull) r[k] = __int_as_float(0x7fc00000);
        }
        stg_stream_v4(out + 8 * i, __float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3]));
        stg_stream_v4(out + 8 * i + 4, __float_as_uint(r[4]), __float_as_uint(r[5]), __float_as_uint(r[6]), __float_as_uint(r[7]));
    }
    for (long long i = 8 * n8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __fmul_rn(scale, __fsub_rn((float)in[i], offset));
        out[i] = (invalid && invalid[i]) ? __int_as_float(0x7fc00000) : v;
    }
}

// D:1663 (np.histogram input): exact per-value counts of a uint8 raster.
// LANE-PRIVATE byte counters: a warp owns a [64 words][32 lanes] block of shared memory (8 KB); the counter of value v of
// lane l is byte (v & 3) of word [v >> 2][l] - bank l whatever v is, so the fire-and-forget shared-memory adds
// (RED.shared) of a warp never conflict, on uniform noise as on flat areas (the first build's per-warp 256-bin
// histograms serialised on equal values and ran at 0.07-0.10 of the HBM peak, profiles/).  A byte counter holds 255:
// after at most 240 increments per lane the warp folds its block into registers (ATOMS.EXCH reads and clears a word;
// lane l owns word rows l and l + 32, walks the lanes' copies rotated so that every access hits 32 banks), once per
// kernel at HLS tile size.  ONE CTA of 16 warps per SM (128 KB), 3 x 16-byte loads in flight per lane.
constexpr int HIST_WARPS = 16;
constexpr int HIST_WORDS_PER_WARP = 64 * 32;
constexpr size_t HIST_SMEM_BYTES = (size_t)(HIST_WARPS * HIST_WORDS_PER_WARP + 256) * sizeof(unsigned int);
__global__ void __launch_bounds__(32 * HIST_WARPS, 1) histogram_u8_kernel(const uint8_t *__restrict__ in, long long n,
                                                                         unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned int hist_sm[];
    unsigned int *cta_hist = hist_sm + HIST_WARPS * HIST_WORDS_PER_WARP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < HIST_WARPS * HIST_WORDS_PER_WARP + 256; i += 32 * HIST_WARPS) hist_sm[i] = 0u;
    __syncthreads();
    const uint32_t warp_base = smem_u32(hist_sm) + 4u * (uint32_t)(warp * HIST_WORDS_PER_WARP);
    const uint32_t mine = warp_base + 4u * (uint32_t)lane;
    uint32_t acc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};       // bins 4 lane + k (k < 4) and 128 + 4 lane + k
    auto count = [&](uint32_t v) {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(mine + 128u * (v >> 2)), "r"(1u << (8u * (v & 3u))) : "memory");
    };
    auto count4 = [&](uint32_t x) { count(x & 255u); count((x >> 8) & 255u); count((x >> 16) & 255u); count(x >> 24); };
    auto fold = [&]() {
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t row = warp_base + 128u * (uint32_t)(lane + 32 * half);
            uint32_t lo = 0u, hi = 0u;                         // 16x2 sums: 32 copies x 255 < 2^16
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                uint32_t w;
                asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(w) : "r"(row + 4u * (uint32_t)((lane + i) & 31)), "r"(0u) : "memory");
                lo += w & 0x00ff00ffu;
                hi += (w >> 8) & 0x00ff00ffu;
            }
            acc[4 * half + 0] += lo & 0xffffu; acc[4 * half + 2] += lo >> 16;
            acc[4 * half + 1] += hi & 0xffffu; acc[4 * half + 3] += hi >> 16;
        }
        __syncwarp();
    };
    const long long n16 = (reinterpret_cast<uintptr_t>(in) & 15) ? 0 : n / 16;
    const long long stride = (long long)gridDim.x * blockDim.x;
    int pending = 0;                                           // increments of a lane since the last fold (warp-uniform bound)
    for (long long base = blockIdx.x * (long long)blockDim.x; base < n16; base += 3 * stride) {   // warp-uniform trip count
        const long long i0 = base + tid, i1 = i0 + stride, i2 = i1 + stride;
        int4 x0, x1, x2;
        const bool on0 = i0 < n16, on1 = i1 < n16, on2 = i2 < n16;
        if (on0) x0 = ldg_stream_v4(in + 16 * i0);
        if (on1) x1 = ldg_stream_v4(in + 16 * i1);
        if (on2) x2 = ldg_stream_v4(in + 16 * i2);
        if (on0) { count4((uint32_t)x0.x); count4((uint32_t)x0.y); count4((uint32_t)x0.z); count4((uint32_t)x0.w); }
        if (on1) { count4((uint32_t)x1.x); count4((uint32_t)x1.y); count4((uint32_t)x1.z); count4((uint32_t)x1.w); }
        if (on2) { count4((uint32_t)x2.x); count4((uint32_t)x2.y); count4((uint32_t)x2.z); count4((uint32_t)x2.w); }
        pending += 48;
        if (pending > 255 - 48) { fold(); pending = 0; }
    }
    for (long long base = 16 * n16 + blockIdx.x * (long long)blockDim.x; base < n; base += stride) {
        const long long i = base + tid;
        if (i < n) count(in[i]);
        if (++pending == 255) { fold(); pending = 0; }
    }
    fold();
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (acc[k]) atomicAdd(&cta_hist[(k >> 2) * 128 + 4 * lane + (k & 3)], acc[k]);
    __syncthreads();
    if (tid < 256 && cta_hist[tid]) atomicAdd(&counts[tid], (unsigned long long)cta_hist[tid]);
}
// D:1684: image > threshold for a uint8 image and a float64 threshold == image >= limit with an integer limit
__global__ void greater_than_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n, int limit) {
    const long long n16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) ? 0 : n / 16;
    PB200_GRID_STRIDE(i, n16) {
        const int4 x = ldg_stream_v4(in + 16 * i);
        uint32_t w[4] = {(uint32_t)x.x, (uint32_t)x.y, (uint32_t)x.z, (uint32_t)x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t r = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) r |= ((int)((w[k] >> (8 * b)) & 255u) >= limit ? 1u : 0u) << (8 * b);
            w[k] = r;
        }
        stg_stream_v4(out + 16 * i, w[0], w[1], w[2], w[3]);
    }
    for (long long i = 16 * n16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (int)in[i] >= limit ? 1 : 0;
}

// D:4215-4283 over a whole DEM incl. np.gradient's one-sided border differences
__global__ void shadow_kernel(const float *__restrict__ dem, int rows, int cols, uint8_t *__restrict__ out,
                              SunTerms S, const __grid_constant__ DevParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const size_t i = (size_t)y * cols + x;
    float g_col, g_row;
    if (x == 0) g_col = __fsub_rn(dem[i + 1], dem[i]);                    // (f[1]-f[0]) / 1.0
    else if (x == cols - 1) g_col = __fsub_rn(dem[i], dem[i - 1]);
    else g_col = __fmul_rn(__fsub_rn(dem[i + 1], dem[i - 1]), 0.5f);
    if (y == 0) g_row = __fsub_rn(dem[i + cols], dem[i]);
    else if (y == rows - 1) g_row = __fsub_rn(dem[i], dem[i - cols]);
    else g_row = __fmul_rn(__fsub_rn(dem[i + cols], dem[i - cols]), 0.5f);
    out[i] = (uint8_t)shadow_from_gradient(g_col, g_row, P.dxf, P.dyf, S, P.cos_thr, P.tan_thr, (P.flags & PF_NUMPY1) != 0u);
}

// The same for a float64 DEM - what np.gradient makes of an integer-typed DEM too (D:4255: integer input is converted
// to float64 first): every operation of D:4255-4281 in float64.
__global__ void shadow_f64_kernel(const double *__restrict__ dem, int rows, int cols, uint8_t *__restrict__ out,
                                  SunTerms S, double dx, double dy_neg, double cos_thr, double tan_thr) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const size_t i = (size_t)y * cols + x;
    double g_col, g_row;
    if (x == 0) g_col = __dsub_rn(dem[i + 1], dem[i]);
    else if (x == cols - 1) g_col = __dsub_rn(dem[i], dem[i - 1]);
    else g_col = __dmul_rn(__dsub_rn(dem[i + 1], dem[i - 1]), 0.5);
    if (y == 0) g_row = __dsub_rn(dem[i + cols], dem[i]);
    else if (y == rows - 1) g_row = __dsub_rn(dem[i], dem[i - cols]);
    else g_row = __dmul_rn(__dsub_rn(dem[i + cols], dem[i - cols]), 0.5);
    const double nx = __ddiv_rn(-g_col, dx), ny = __ddiv_rn(-g_row, dy_neg);                          // D:4260-4261
    const double nf = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), 1.0));    // D:4264
    const double s = __dadd_rn(__dmul_rn(nx, S.sin_az), __dmul_rn(ny, S.cos_az));                     // D:4275-4277
    uint32_t lit = 1u;
    if (s <= tan_thr) {                                                                               // back slope
        const double xc = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, S.sx), __dmul_rn(ny, S.sy)), S.sz), nf);
        lit = (xc >= cos_thr && xc <= 1.0) ? 1u : 0u;                                                 // D:4267-4280
    }
    out[i] = (uint8_t)lit;
}

// exhaustive (n, d) in int16^2 check of the integer ratio test against IEEE
// float64 division (numpy's int16 / int16 -> float64 true_divide)
__global__ void ratio_sweep_kernel(int a, int b, double t, int is_less, unsigned long long *mismatches) {
    const int d = (int)blockIdx.x - 32768;
    unsigned int bad = 0;
    for (int n = -32768 + (int)threadIdx.x; n <= 32767; n += blockDim.x) {
        const Ratio r = make_ratio(n, d);
        const double q = __ddiv_rn((double)n, (double)d);
        const bool ref = is_less ? (q < t) : (q > t);
        const bool got = is_less ? ratio_lt_any(r, a, b) : ratio_gt_any(r, a, b);
        bad += (ref != got);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

}  // namespace pb200
