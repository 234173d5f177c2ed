// pb200_stream.cuh - K1s: the fused classification with TMA-FED inputs (lean variant: the four graded layers).
//
// Same items, same per-pixel code (pb200_fused_row.inc) and same DEM staging as dswx_fused_fast_kernel, but the
// reflectance bands and the byte rasters no longer travel global memory -> registers through per-lane loads (9 address
// computations + 9 LDG per lane and row, latency hidden by a register prefetch one row ahead - all a 24-warp CTA at
// 80 registers can afford; with the loads replaced by register arithmetic that kernel runs 28 % faster, profiles/).
// Here a PRODUCER WARP streams them with the tensor memory accelerator into a two-slot shared-memory ring:
//
//  * every plane is described to TMA as a 2-D tensor of "super-rows" of FOUR raster rows (4 W elements; the row pitch
//    of an HLS raster, 3660 elements, is not a multiple of 16 bytes, four rows are: 4 W * 2 B and 4 W * 1 B with
//    W % 4 == 0).  The raster rows r = c (mod 4) of an item are then the CONSECUTIVE super-rows of one 2-D box that
//    starts at column c W + x0 - a box of 24 rows is "row c of every warp" (a warp owns 4 consecutive rows of the
//    128 x 92 item);
//  * a chunk = that box for each of the 9 input planes: 6 x (136 int16 x 24) + 3 x (144 B x 24) = 48.4 KB, 9 TMA
//    instructions for 3072 pixels.  Boxes start on 16-byte boundaries (measured requirement on B200): the column is
//    rounded down, a lane adds the remainder (0 / 4 int16 elements, 0 / 4 / 8 / 12 bytes) to its shared-memory address;
//  * ring protocol per chunk q (slot q & 1): the producer waits for the 24 consumer arrivals of chunk q - 2 on
//    empty_in[slot], arms full_in[slot] with the byte count and issues the boxes; a consumer warp waits for
//    full_in[slot], copies its row into registers (6 LDS.64 + 3 LDS.32 per lane, conflict-free), arrives on
//    empty_in[slot] and classifies the row.  The ring is released as soon as the row is in registers, so the producer
//    runs up to two chunks (half an item) ahead of the slowest warp - far more than the DRAM latency;
//  * the producer also requests the DEM tile of every item (one box per item into the double buffer of FastSmem); an
//    item of a tile without a DEM gets a plain arrival, so the phases of full[] / empty[] count items.
//
// Preconditions (plan_build checks them per tile; anything else runs dswx_fused_fast_kernel): height % 4 == 0,
// width % 4 == 0 and >= 36, every input plane 16-byte aligned, all four graded layers and the counters requested.
#pragma once
#include "pb200_fused.cuh"

namespace pb200 {

// 23 consumer warps + the producer = 24 warps = 6 per SM sub-partition: 80 registers each (a 25th warp would put 7 on
// one sub-partition and cap every thread at 72 registers: measured 165 instead of 215 Gpixel/s, profiles/)
#ifndef PB200_ST_WARPS
#define PB200_ST_WARPS 23
#endif
#ifndef PB200_ST_CONTIGUOUS
#define PB200_ST_CONTIGUOUS 0      // a CTA takes a contiguous range of the item list (1) or every gridDim-th item (0)
#endif
#ifndef PB200_ST_HOLD_LANE_BAND
#define PB200_ST_HOLD_LANE_BAND 0  // keep the lane's band address in a register (opaque) instead of re-deriving it per row
#endif
#ifndef PB200_ST_SLEEP_NS
#define PB200_ST_SLEEP_NS 200      // producer: nanoseconds between polls of an empty barrier
#endif
#ifndef PB200_ST_PRODUCER_HWWAIT
#define PB200_ST_PRODUCER_HWWAIT 0  // producer waits with the (hinted) hardware try_wait instead of nanosleep polling
#endif
#if PB200_ST_PRODUCER_HWWAIT
#define ST_PRODUCER_WAIT(bar, parity) mbar_wait(bar, parity)
#else
#define ST_PRODUCER_WAIT(bar, parity) mbar_wait_sleep(bar, parity, PB200_ST_SLEEP_NS)
#endif
constexpr int ST_WARPS = PB200_ST_WARPS;                      // consumer warps; one more warp produces
constexpr int ST_THREADS = 32 * (ST_WARPS + 1);
constexpr int ST_ROWS_PER_WARP = 4;                           // a chunk is one row class modulo 4 of the item
constexpr int ST_H = ST_ROWS_PER_WARP * ST_WARPS;             // item height of this kernel (92 rows)
static_assert(ST_H + 2 <= FT_SMH, "the DEM tile of an item fits the double buffer of FastSmem");
constexpr uint32_t ST_DEM_BOX_BYTES = (ST_H + 2) * FT_SMW * sizeof(float);
constexpr int ST_BAND_W = FT_W + 8;                           // int16 elements per staged band row (272 B)
constexpr int ST_BYTE_W = FT_W + 16;                          // bytes per staged byte-raster row
constexpr int ST_MAPS = 10;                                   // tensor maps per tile: DEM, 6 bands, Fmask, LAND, ocean
enum { SM_DEM = 0, SM_BAND0 = 1, SM_FMASK = 7, SM_LAND = 8, SM_OCEAN = 9 };

// one plane of a slot = ST_WARPS rows; planes start on 128-byte boundaries (TMA destination alignment)
constexpr uint32_t ST_BAND_TX = sizeof(int16_t) * ST_WARPS * ST_BAND_W, ST_BYTE_TX = ST_WARPS * ST_BYTE_W;   // bytes per box
constexpr uint32_t ST_BAND_BYTES = (ST_BAND_TX + 127u) & ~127u, ST_BYTE_BYTES = (ST_BYTE_TX + 127u) & ~127u;   // plane pitch
struct __align__(128) InSlot {
    unsigned char band[6][ST_BAND_BYTES];                     // [ST_WARPS][ST_BAND_W] int16 each: 6 x 6528 B at 24 warps
    unsigned char byte[3][ST_BYTE_BYTES];                     // [ST_WARPS][ST_BYTE_W] bytes: Fmask, LAND, ocean
};

struct __align__(128) StreamSmem {
    FastSmem f;                                               // DEM double buffer, tables, tile descriptor, DEM barriers
    InSlot in[2];
    unsigned long long full_in[2], empty_in[2];
};

#define SS_OFF(member) ((uint32_t)offsetof(StreamSmem, member))

// Which items a CTA classifies: every gridDim-th item of the (tile-major, row-major) list, so that at any moment the
// 148 CTAs read 148 NEIGHBOURING items - the same DRAM pages of every plane, the halo rows and the over-fetched box
// edges of one CTA are the neighbour's payload.  Contiguous ranges per CTA (a CTA then stays inside one or two tiles
// and never re-synchronises for a tile change) measured 207 instead of 221 Gpixel/s (profiles/README.md).
#if PB200_ST_CONTIGUOUS
#define ST_ITEM_FIRST ((int)(((long long)blockIdx.x * n_items) / gridDim.x))
#define ST_ITEM_END ((int)(((long long)(blockIdx.x + 1) * n_items) / gridDim.x))
#define ST_ITEM_STEP 1
#else
#define ST_ITEM_FIRST ((int)blockIdx.x)
#define ST_ITEM_END n_items
#define ST_ITEM_STEP ((int)gridDim.x)
#endif

// Registers are allocated per SM sub-partition (16 384 each): 25 warps put 7 on one of them -> at most 72 registers per
// thread; 21 warps (20 consumers) put 6 -> 80.
template <bool FAST8>
__global__ void __launch_bounds__(ST_THREADS, 1)
dswx_fused_stream_kernel(const TileDev *__restrict__ tiles, const CUtensorMap *__restrict__ tmaps,
                         const FusedTables *__restrict__ tables, const ItemDesc *__restrict__ items, int n_items,
                         const __grid_constant__ DevParams P, const __grid_constant__ FastParams F) {
    constexpr bool OPTIONAL_LAYERS = false, ALL_GRADED = true;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamSmem &S = *reinterpret_cast<StreamSmem *>(smem_raw);
    FastSmem &s = S.f;
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(&s);    // FastSmem sits at offset 0 of StreamSmem
    asm volatile("mov.b32 %0, %0;" : "+r"(sb));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- tables and barriers: once per CTA ------------------------------------
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables);
        uint4 *dst = reinterpret_cast<uint4 *>(s.big_lut);
        for (int i = tid; i < (int)(sizeof(FusedTables) / 16); i += ST_THREADS) dst[i] = __ldg(src + i);
        if (tid == 0) {
            mbar_init(&s.full[0], 1); mbar_init(&s.full[1], 1);
            mbar_init(&s.empty[0], ST_WARPS); mbar_init(&s.empty[1], ST_WARPS);
            mbar_init(&S.full_in[0], 1); mbar_init(&S.full_in[1], 1);
            mbar_init(&S.empty_in[0], ST_WARPS); mbar_init(&S.empty_in[1], ST_WARPS);
        }
    }
    __syncthreads();

    // =========================== producer warp ==================================
    if (warp == ST_WARPS) {
        if (lane != 0) return;
        uint32_t k = 0, acquired_tile = 0xffffffffu;
#pragma unroll 1
        for (int it = ST_ITEM_FIRST; it < ST_ITEM_END; it += ST_ITEM_STEP, ++k) {
            const ItemDesc d = items[it];
            const TileDev &g = tiles[d.tile];
            const CUtensorMap *tm = tmaps + (size_t)d.tile * ST_MAPS;
            if (d.tile != acquired_tile) {
                // tensor maps live in global memory: one tensormap-proxy acquire per map when the CTA first meets the tile
                // (a system-scope fence costs microseconds - per box it made the producer the bottleneck, 75 Gpixel/s)
#pragma unroll 1
                for (int j = 0; j < ST_MAPS; ++j) tma_acquire_map(&tm[j]);
                acquired_tile = d.tile;
            }
            const int W = __ldg(&g.width);
            const bool has_dem = __ldg(reinterpret_cast<const unsigned long long *>(&g.dem)) != 0ull;
            const bool has_land = __ldg(reinterpret_cast<const unsigned long long *>(&g.land)) != 0ull;
            const bool has_ocean = __ldg(reinterpret_cast<const unsigned long long *>(&g.ocean)) != 0ull;
            const int x0 = d.tx * FT_W, row4 = d.ty * (ST_H / 4);            // item rows start at super-row ty * ST_WARPS
            {
                // DEM tile of the item (or a plain arrival: the phases of full[] / empty[] count items)
                const uint32_t b = k & 1u;
                if (k >= 2u) ST_PRODUCER_WAIT(&s.empty[b], ((k >> 1) - 1u) & 1u);
                if (has_dem) {
                    const int dox = __ldg(&g.dem_off_x), doy = __ldg(&g.dem_off_y);
                    const int padx = DEM_PADX + (dox & 3);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&s.full[b], ST_DEM_BOX_BYTES);
                    tma_load_2d_acquired(&s.dem[b].v[0][0], &tm[SM_DEM], dox + x0 - padx, doy + d.ty * ST_H - 1, &s.full[b]);
                } else {
                    mbar_arrive(&s.full[b]);
                }
            }
#pragma unroll 1
            for (int c = 0; c < ST_ROWS_PER_WARP; ++c) {
                const uint32_t q = 4u * k + (uint32_t)c, slot = q & 1u;
                if (q >= 2u) ST_PRODUCER_WAIT(&S.empty_in[slot], ((q >> 1) - 1u) & 1u);
                // generic-proxy reads of the slot (ordered by the empty barrier) before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&S.full_in[slot], 6u * ST_BAND_TX + ST_BYTE_TX * (1u + (has_land ? 1u : 0u) + (has_ocean ? 1u : 0u)));
                const int xe = c * W + x0;                                    // column inside the 4-row super-row
                InSlot &in = S.in[slot];
#pragma unroll
                for (int b = 0; b < 6; ++b) tma_load_2d_acquired(&in.band[b][0], &tm[SM_BAND0 + b], xe & ~7, row4, &S.full_in[slot]);
                tma_load_2d_acquired(&in.byte[0][0], &tm[SM_FMASK], xe & ~15, row4, &S.full_in[slot]);
                if (has_land) tma_load_2d_acquired(&in.byte[1][0], &tm[SM_LAND], xe & ~15, row4, &S.full_in[slot]);
                if (has_ocean) tma_load_2d_acquired(&in.byte[2][0], &tm[SM_OCEAN], xe & ~15, row4, &S.full_in[slot]);
            }
        }
        return;
    }

    // =========================== consumer warps =================================
    const int rgrp = warp;
    uint32_t cur_tile = 0xffffffffu;
    uint32_t acc_vc = 0, acc_nno = 0;
    unsigned long long acc_hist = 0ull;
    constexpr bool histogram = false;                         // lean variant: the three counters
    auto flush_counters = [&]() {
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
        }
        acc_vc = 0; acc_nno = 0; acc_hist = 0ull;
    };
    auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(32 * ST_WARPS) : "memory"); };

    uint32_t w[6][2], fm4 = 0u, ld4 = 0xffffffffu, oc4 = 0x01010101u;
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) w[kk][0] = w[kk][1] = 0u;

    // shared addresses of this lane's 4 pixels in row `warp` of slot 0: bands (plane 0) and byte rasters (Fmask plane)
    uint32_t lane_band = sb + SS_OFF(in) + (uint32_t)rgrp * (2u * ST_BAND_W) + 8u * (uint32_t)lane;
#if PB200_ST_HOLD_LANE_BAND
    asm volatile("" : "+r"(lane_band));
#endif
    const uint32_t lane_byte = sb + SS_OFF(in) + 6u * ST_BAND_BYTES + (uint32_t)rgrp * ST_BYTE_W + 4u * (uint32_t)lane;

    uint32_t k = 0;
#pragma unroll 1
    for (int it = ST_ITEM_FIRST; it < ST_ITEM_END; it += ST_ITEM_STEP, ++k) {
        const ItemDesc item = items[it];
        if (item.tile != cur_tile) {
            // tile change: the only synchronisation among all consumer warps (the producer runs ahead on its own)
            consumer_sync();                              // every warp is done with the previous tile's descriptor
            flush_counters();
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
                    // see dswx_fused_fast_kernel: tiles whose sun vector breaks the preconditions of the sign-bit shortcut
                    // run the exact sequence on every pixel
                    const double hz = g.sx * g.sin_az + g.sy * g.cos_az, n2 = g.sx * g.sx + g.sy * g.sy + g.sz * g.sz;
                    if (!(hz >= 0.0 && fabs(n2 - 1.0) < 1e-9 && fabs(g.sin_az * g.sin_az + g.cos_az * g.cos_az - 1.0) < 1e-9))
                        s.sun32[SK_XX] = __int_as_float(0x7fffffff);
                }
            }
            cur_tile = item.tile;
            ld4 = 0xffffffffu; oc4 = 0x01010101u;         // defaults of a tile without LAND / ocean raster
            consumer_sync();
        }
        const int W = s.tile.width, H = s.tile.height;
        const int x0 = item.tx * FT_W, y0 = item.ty * ST_H;
        // which optional rasters the tile has, as ONE opaque register: kept as `pointer != nullptr` the compiler holds the
        // two 64-bit pointers across the row loop (predicates do not survive the out-of-line calls) and re-tests them per row
        uint32_t tflags = (s.tile.dem != nullptr ? 1u : 0u) | (s.tile.land != nullptr ? 2u : 0u) | (s.tile.ocean != nullptr ? 4u : 0u);
        asm volatile("" : "+r"(tflags));
        const bool has_dem = (tflags & 1u) != 0u, has_land = (tflags & 2u) != 0u, has_ocean = (tflags & 4u) != 0u;
        constexpr bool has_counters = true, want_shad = false, all_graded = true;
        const uint32_t buf = k & 1u;
        bool dem_ready = false;                           // every warp observes the item's DEM phase before it arrives

        const int x = x0 + 4 * lane;
        // rows of this item the lane classifies: 0 for lanes right of the raster (one live value instead of x and a row count)
        const int nrows = x < W ? min(ST_ROWS_PER_WARP, H - (y0 + rgrp * ST_ROWS_PER_WARP)) : 0;   // may be <= 0
        uint32_t pix = (uint32_t)(y0 + rgrp * ST_ROWS_PER_WARP) * (uint32_t)W + (uint32_t)x;
#pragma unroll 1
        for (int rr = 0; rr < ST_ROWS_PER_WARP; ++rr, pix += (uint32_t)W) {
            // ---- this warp's row of chunk q: shared memory -> registers, then the slot is free again -------------
            // chunk q = 4 k + rr of this CTA: slot q & 1 = rr & 1, and the phase of its barrier (q >> 1) & 1 = (rr >> 1) & 1 -
            // both independent of the item (4 chunks per item, 2 slots)
            const uint32_t slot = (uint32_t)rr & 1u;
            mbar_wait_addr(sb + SS_OFF(full_in) + 8u * slot, ((uint32_t)rr >> 1) & 1u);
            const bool active = rr < nrows;
            if (active) {
                const uint32_t xe = (uint32_t)rr * (uint32_t)W + (uint32_t)x0;
                const uint32_t base = slot * (uint32_t)sizeof(InSlot);
                const uint32_t ab = base + lane_band + 2u * (xe & 7u);
                const uint32_t ay = base + lane_byte + (xe & 15u);
#pragma unroll
                for (int kk = 0; kk < 6; ++kk) {
                    uint32_t v0, v1;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(ab + (uint32_t)kk * ST_BAND_BYTES));
                    w[kk][0] = v0; w[kk][1] = v1;
                }
                fm4 = lds_u32(ay);
                if (has_land) ld4 = lds_u32(ay + ST_BYTE_BYTES);
                if (has_ocean) oc4 = lds_u32(ay + 2u * ST_BYTE_BYTES);
                // the row is IN REGISTERS (not merely requested) before the slot is handed back to the producer
                asm volatile("" ::"r"(w[5][1]), "r"(fm4), "r"(ld4), "r"(oc4) : "memory");
            }
            __syncwarp();
            mbar_arrive_elected(sb + SS_OFF(empty_in) + 8u * slot);      // one lane of the (converged) warp
            if (!active) continue;
            const int ly = rgrp * ST_ROWS_PER_WARP + rr;
#define FT_DEM_WAIT() mbar_wait_addr(sb + FS_OFF(full) + 8u * buf, (k >> 1) & 1u)
#define FT_ROW_MIDPOINT() do { } while (0)
// re-read from the tile descriptor where the shadow block needs it: no register (or spill slot) held across the rows
#define FT_PADX() (DEM_PADX + (int)(lds_u32(sb + FS_TILE(dem_off_x)) & 3u))
#include "pb200_fused_row.inc"
#undef FT_PADX
#undef FT_ROW_MIDPOINT
#undef FT_DEM_WAIT
        }
        // every warp sees the item's DEM phase complete before it arrives on empty[]: the producer's request for item
        // k + 2 waits for all arrivals of item k, so no warp can run two items ahead and arrive twice in one phase
        __syncwarp();
        if (!dem_ready) mbar_wait_addr(sb + FS_OFF(full) + 8u * buf, (k >> 1) & 1u);
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(sb + FS_OFF(empty) + 8u * buf);
    }
    flush_counters();
}

}  // namespace pb200
