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
#ifndef PB200_ST_ROTATE
#define PB200_ST_ROTATE 0          // row of chunk c that warp w classifies: (w + PB200_ST_ROTATE * c) mod ST_WARPS - a warp then samples the
#endif                             // whole height of the item instead of 4 neighbouring rows (water is blocky: balances the shadow work)
#if PB200_ST_ROTATE
#define PB200_ST_PIX_STEP
#else
#define PB200_ST_PIX_STEP , pix += (uint32_t)W
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
#if !PB200_ST_ROTATE
        uint32_t pix = (uint32_t)(y0 + rgrp * ST_ROWS_PER_WARP) * (uint32_t)W + (uint32_t)x;
#endif
#pragma unroll 1
        for (int rr = 0; rr < ST_ROWS_PER_WARP; ++rr PB200_ST_PIX_STEP) {
            // ---- this warp's row of chunk q: shared memory -> registers, then the slot is free again -------------
            // chunk q = 4 k + rr of this CTA: slot q & 1 = rr & 1, and the phase of its barrier (q >> 1) & 1 = (rr >> 1) & 1 -
            // both independent of the item (4 chunks per item, 2 slots)
            const uint32_t slot = (uint32_t)rr & 1u;
            mbar_wait_addr(sb + SS_OFF(full_in) + 8u * slot, ((uint32_t)rr >> 1) & 1u);
#if PB200_ST_ROTATE
            uint32_t rrow = (uint32_t)rgrp + (uint32_t)(PB200_ST_ROTATE * rr);
            if (rrow >= (uint32_t)ST_WARPS) rrow -= (uint32_t)ST_WARPS;
            const int ly = 4 * (int)rrow + rr;
            const bool active = x < W && y0 + ly < H;
            const uint32_t pix = (uint32_t)(y0 + ly) * (uint32_t)W + (uint32_t)x;
#else
            const bool active = rr < nrows;
#endif
            if (active) {
                const uint32_t xe = (uint32_t)rr * (uint32_t)W + (uint32_t)x0;
                const uint32_t base = slot * (uint32_t)sizeof(InSlot);
#if PB200_ST_ROTATE
                const uint32_t ab = base + sb + SS_OFF(in) + rrow * (2u * ST_BAND_W) + 8u * (uint32_t)lane + 2u * (xe & 7u);
                const uint32_t ay = base + sb + SS_OFF(in) + 6u * ST_BAND_BYTES + rrow * ST_BYTE_W + 4u * (uint32_t)lane + (xe & 15u);
#else
                const uint32_t ab = base + lane_band + 2u * (xe & 7u);
                const uint32_t ay = base + lane_byte + (xe & 15u);
#endif
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
#if !PB200_ST_ROTATE
            const int ly = rgrp * ST_ROWS_PER_WARP + rr;
#endif
#define FT_DEM_WAIT() mbar_wait_addr(sb + FS_OFF(full) + 8u * buf, (k >> 1) & 1u)
#define FT_ROW_MIDPOINT() do { } while (0)
// re-read from the tile descriptor where the shadow block needs it: no register (or spill slot) held across the rows
#define FT_PADX() (DEM_PADX + (int)(lds_u32(sb + FS_TILE(dem_off_x)) & 3u))
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
        // every warp sees the item's DEM phase complete before it arrives on empty[]: the producer's request for item
        // k + 2 waits for all arrivals of item k, so no warp can run two items ahead and arrive twice in one phase
        __syncwarp();
        if (!dem_ready) mbar_wait_addr(sb + FS_OFF(full) + 8u * buf, (k >> 1) & 1u);
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(sb + FS_OFF(empty) + 8u * buf);
    }
    flush_counters();
}

// ---------------------------------------------------------------------------
// K1d: the same TMA-fed kernel with DYNAMIC ROW ASSIGNMENT and a deep ring.
//
// In dswx_fused_stream_kernel warp w always classifies row w of every chunk.  Rows differ in cost (the terrain-shadow
// block runs only for rows that hold a water-class pixel), the two-slot ring lets the warps drift apart by little more
// than one row, and so warps with cheap rows wait for the ones with expensive rows: 16 of 123 executed thread-instructions
// per pixel were polls of full_in, and evaluating the shadow shortcut for EVERY row cost nothing (profiles/).  Here a row
// goes to whichever warp is free, and the ring is deep enough that nobody waits for data:
//
//  * items are 128 x 64 pixels (16 super-rows); a chunk = the 16 rows of one row class of the item (9 boxes of 16 rows,
//    33 KB); the DEM tile of an item is 66 x 136 floats - together 4 ring slots + 2 DEM tiles + tables = 216 KB.  A slot
//    is refilled as soon as its 16 rows have been copied to registers: between 3 and 4 chunks (2 - 2.8 row times of the
//    23 warps) are always requested ahead of the row being handed out;
//  * one ticket counter per CTA in shared memory; ticket t = row (t mod 16) of chunk q = t div 16 of the CTA's chunk
//    sequence (item k = q div 4, row class c = q mod 4 = ring slot).  A warp takes a ticket, waits for that chunk, copies
//    its row to registers, releases the ring slot, classifies, takes the next ticket;
//  * every slot has TWO full barriers used by alternate generations (= items): a warp that waits for generation k waits
//    on a barrier whose previous user was generation k - 2, complete long ago - the parity test cannot alias however far
//    the warps drift apart (8 chunks would need 128 outstanding tickets; there are 23 warps);
//  * nothing ties a warp to a tile: a row reads what it needs of its item and tile - item position, raster size, output
//    pointers, float32 sun constants: records the HOST builds per tile (TileSlot) - from global memory (the item, one
//    16-byte first-look vector and two 16-byte pointer pairs per row, L1 hits), so no descriptor is written into shared
//    memory by anybody; at most the items k and k + 1 are in flight, because
//    the producer requests the DEM tile of item k + 2 only after all 64 rows of item k have been released (empty[k & 1],
//    one arrival PER ROW, after the row's last look at the DEM tile) and after it has itself seen item k's DEM
//    transaction complete;
//  * the coverage counters stay in per-thread registers and are flushed (red.global) when a warp's rows move on to
//    another tile, and at the end.
// No synchronisation among the consumer warps at all (no named barrier at tile changes).
constexpr int SD_CH = 16;                                     // rows per chunk = TMA box height = tickets per chunk
constexpr int SD_H = 4 * SD_CH;                               // item height (64 rows)
constexpr int SD_NSLOT = 4;                                   // ring slots (= chunks per item: slot = row class, generation = item)
static_assert((SD_CH & (SD_CH - 1)) == 0 && SD_NSLOT == 4, "ticket -> (chunk, row, slot, generation) by shifts and masks");
constexpr int SD_SMH = SD_H + 2;                              // DEM tile rows
constexpr uint32_t SD_DEM_BOX_BYTES = SD_SMH * FT_SMW * sizeof(float);
constexpr uint32_t SD_BAND_TX = sizeof(int16_t) * SD_CH * ST_BAND_W, SD_BYTE_TX = SD_CH * ST_BYTE_W;          // bytes per box
constexpr uint32_t SD_BAND_BYTES = (SD_BAND_TX + 127u) & ~127u, SD_BYTE_BYTES = (SD_BYTE_TX + 127u) & ~127u;  // plane pitch
struct __align__(128) DynSlot {
    unsigned char band[6][SD_BAND_BYTES];                     // [SD_CH][ST_BAND_W] int16 each
    unsigned char byte[3][SD_BYTE_BYTES];                     // [SD_CH][ST_BYTE_W] bytes: Fmask, LAND, ocean
};
struct __align__(128) DynDem { float v[SD_SMH][FT_SMW]; };
// One record per tile, built on the HOST (plan_build), read by the rows through the L1 cache: the tile descriptor, the
// float32 sun constants of the shadow shortcut, and the 16 bytes a row needs first.
struct __align__(16) TileSlot {
    TileDev tile; float sun32[12];
    uint4 row_info;                                           // width, height, left pad of the DEM tile, TSF_* | tile index << 3
    unsigned long long out_ptrs[4];                           // DIAG, WTR, BWTR, CONF planes: two 16-byte loads
};
static_assert(sizeof(TileSlot) % 16 == 0, "16-byte vector loads of row_info / sun32");
struct __align__(128) StreamDynSmem {
    DynDem dem[2];
    // the table block in the order of FusedTables / FastSmem (the row body addresses it relative to big_lut)
    uint32_t big_lut[2048];
    uint32_t diag_lut[128];
    uint8_t  fk_lut[4096];
    uint8_t  land_lut[256];
    uint8_t  kill_lut[128];
    unsigned long long full[2], empty[2];                     // DEM tile of item k: transaction barrier / 64 row releases
    unsigned long long full_in[SD_NSLOT][2], empty_in[SD_NSLOT];
    unsigned int ticket;
    DynSlot in[SD_NSLOT];
};
static_assert(sizeof(StreamDynSmem) <= 227 * 1024, "fits the shared memory of an SM");
#define SD_OFF(member) ((uint32_t)offsetof(StreamDynSmem, member))
enum : uint32_t { TSF_DEM = 1u, TSF_LAND = 2u, TSF_OCEAN = 4u };   // TileSlot::tile.pad_ and row_info.w: which rasters the tile has

// ALL_GRADED: every tile of the batch writes all four graded layers (no pointer tests in the row loop); without it any
// subset of them (BASELINE configs[0]: DIAG + WTR of tiles without DEM / LAND / ocean).
template <bool FAST8, bool ALL_GRADED = true>
__global__ void __launch_bounds__(ST_THREADS, 1)
dswx_fused_stream_dyn_kernel(const TileSlot *__restrict__ slots, const CUtensorMap *__restrict__ tmaps,
                             const FusedTables *__restrict__ tables, const ItemDesc *__restrict__ items, int n_items,
                             const __grid_constant__ DevParams P, const __grid_constant__ FastParams F) {
    constexpr bool OPTIONAL_LAYERS = false;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamDynSmem &D = *reinterpret_cast<StreamDynSmem *>(smem_raw);
    uint32_t db = (uint32_t)__cvta_generic_to_shared(&D);    // shared address of the block, kept in a register
    asm volatile("mov.b32 %0, %0;" : "+r"(db));
    // what the row body calls `sb`: the address FastSmem WOULD start at for its table offsets to land on D's tables
    const uint32_t sb = db + SD_OFF(big_lut) - FS_OFF(big_lut);
    FastSmem &s = *reinterpret_cast<FastSmem *>(smem_raw);    // never dereferenced (optional-layer code of the row body is compiled out)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    (void)s;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tables);
        uint4 *dst = reinterpret_cast<uint4 *>(D.big_lut);
        static_assert(offsetof(StreamDynSmem, kill_lut) - offsetof(StreamDynSmem, big_lut) + 128 == sizeof(FusedTables), "table block mirrors FusedTables");
        static_assert(offsetof(StreamDynSmem, big_lut) % 16 == 0, "16-byte table copies");
        for (int i = tid; i < (int)(sizeof(FusedTables) / 16); i += ST_THREADS) dst[i] = __ldg(src + i);
        if (tid == 0) {
            mbar_init(&D.full[0], 1); mbar_init(&D.full[1], 1);
            mbar_init(&D.empty[0], SD_H); mbar_init(&D.empty[1], SD_H);                 // one arrival per ROW of an item
            for (int i = 0; i < SD_NSLOT; ++i) {
                mbar_init(&D.full_in[i][0], 1); mbar_init(&D.full_in[i][1], 1);
                mbar_init(&D.empty_in[i], SD_CH);                                       // one arrival per row of a chunk
            }
            D.ticket = 0u;
        }
    }
    __syncthreads();
    const int first = (int)blockIdx.x, step = (int)gridDim.x;
    const uint32_t n_loc = first < n_items ? (uint32_t)((n_items - first + step - 1) / step) : 0u;

    // =========================== producer warp ==================================
    if (warp == ST_WARPS) {
        if (lane != 0) return;
        uint32_t acquired_tile = 0xffffffffu;
#pragma unroll 1
        for (uint32_t k = 0; k < n_loc; ++k) {
            const ItemDesc d = items[first + (int)k * step];
            const TileSlot *g = slots + d.tile;                                     // the tile's record in global memory
            const CUtensorMap *tm = tmaps + (size_t)d.tile * ST_MAPS;
            if (d.tile != acquired_tile) {
#pragma unroll 1
                for (int j = 0; j < ST_MAPS; ++j) tma_acquire_map(&tm[j]);
                acquired_tile = d.tile;
            }
            const uint32_t b = k & 1u;
            if (k >= 2u) {
                // item k - 2 (same DEM buffer): every row released, and its transaction complete (a row without a water
                // pixel never waits for the DEM tile itself)
                ST_PRODUCER_WAIT(&D.empty[b], ((k >> 1) - 1u) & 1u);
                mbar_wait(&D.full[b], ((k >> 1) - 1u) & 1u);
            }
            // what the producer needs of the tile: the first-look vector of its record
            const uint4 ri = __ldg(&g->row_info);
            const uint32_t tsf = ri.w;
            const bool has_dem = (tsf & TSF_DEM) != 0u, has_land = (tsf & TSF_LAND) != 0u, has_ocean = (tsf & TSF_OCEAN) != 0u;
            const int W = (int)ri.x;
            const int x0 = d.tx * FT_W, row4 = d.ty * SD_CH;                        // item rows start at super-row ty * SD_CH
            // the DEM tile of the item (or a plain arrival: the phases of full[] count items)
            if (has_dem) {
                const int dox = __ldg(&g->tile.dem_off_x), doy = __ldg(&g->tile.dem_off_y);
                const int padx = DEM_PADX + (dox & 3);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&D.full[b], SD_DEM_BOX_BYTES);
                tma_load_2d_acquired(&D.dem[b].v[0][0], &tm[SM_DEM], dox + x0 - padx, doy + d.ty * SD_H - 1, &D.full[b]);
            } else {
                mbar_arrive(&D.full[b]);
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const uint32_t gen = k, sl = (uint32_t)c;                             // 4 chunks per item, 4 slots
                if (gen >= 1u) ST_PRODUCER_WAIT(&D.empty_in[sl], (gen - 1u) & 1u);
                // generic-proxy reads of the slot (ordered by the empty barrier) before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                unsigned long long *bar = &D.full_in[sl][gen & 1u];
                mbar_expect_tx(bar, 6u * SD_BAND_TX + SD_BYTE_TX * (1u + (has_land ? 1u : 0u) + (has_ocean ? 1u : 0u)));
                const int xe = c * W + x0;                                          // column inside the 4-row super-row
                DynSlot &in = D.in[sl];
#pragma unroll
                for (int bb = 0; bb < 6; ++bb) tma_load_2d_acquired(&in.band[bb][0], &tm[SM_BAND0 + bb], xe & ~7, row4, bar);
                tma_load_2d_acquired(&in.byte[0][0], &tm[SM_FMASK], xe & ~15, row4, bar);
                if (has_land) tma_load_2d_acquired(&in.byte[1][0], &tm[SM_LAND], xe & ~15, row4, bar);
                if (has_ocean) tma_load_2d_acquired(&in.byte[2][0], &tm[SM_OCEAN], xe & ~15, row4, bar);
            }
        }
        // no bulk copy may be in flight into this CTA's shared memory when it retires: the DEM tiles of the last two items
        for (uint32_t k = n_loc >= 2u ? n_loc - 2u : 0u; k < n_loc; ++k) mbar_wait(&D.full[k & 1u], (k >> 1) & 1u);
        return;
    }

    // =========================== consumer warps =================================
    const uint32_t total_rows = n_loc * (uint32_t)SD_H;
    uint32_t acc_vc = 0, acc_nno = 0;
    unsigned long long acc_hist = 0ull;
    constexpr bool histogram = false, has_counters = true, want_shad = false, all_graded = ALL_GRADED;
    uint32_t w[6][2], fm4 = 0u, ld4 = 0xffffffffu, oc4 = 0x01010101u;
#pragma unroll
    for (int kk = 0; kk < 6; ++kk) w[kk][0] = w[kk][1] = 0u;
    auto take_ticket = [&]() {
        // (plain PTX: atomicAdd() on a shared counter compiles to a warp-aggregation sequence of ~15 instructions)
        uint32_t t = 0u;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(db + SD_OFF(ticket)) : "memory");
        return __shfl_sync(0xffffffffu, t, 0);
    };

    // counters: per-thread registers, flushed (fire-and-forget reductions) when the rows of this warp move on to another
    // tile, and at the end
    uint32_t cur_tile = 0xffffffffu;
    auto flush_counters = [&]() {
        const uint32_t wv = __reduce_add_sync(0xffffffffu, acc_vc & 0xffffu);
        const uint32_t wc = __reduce_add_sync(0xffffffffu, acc_vc >> 16);
        const uint32_t wn = __reduce_add_sync(0xffffffffu, acc_nno);
        if (lane == 0 && cur_tile != 0xffffffffu) {
            unsigned long long *cnt = slots[cur_tile].tile.counters;
            if (cnt != nullptr) {
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(cnt), "l"((unsigned long long)wv) : "memory");
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(cnt + 1), "l"((unsigned long long)wc) : "memory");
                asm volatile("red.global.add.u64 [%0], %1;" ::"l"(cnt + 2), "l"((unsigned long long)wn) : "memory");
            }
        }
        acc_vc = 0u; acc_nno = 0u;
    };

    uint32_t t = take_ticket();
#pragma unroll 1
    while (t < total_rows) {
        // ticket -> row r of chunk q; item k = q / 4 (= generation of the ring slot), row class c = q % 4 (= ring slot)
        const uint32_t q = t / (uint32_t)SD_CH, r = t & (uint32_t)(SD_CH - 1);
        const uint32_t k = q >> 2, c = q & 3u, buf = k & 1u, slot = c;
        // the item and its tile: read-only records in global memory (L1 hits after the first row of an item), requested
        // before the wait for the chunk
        const ItemDesc item = items[first + (int)k * step];
        const TileSlot *g = slots + item.tile;
        const uint4 info = __ldg(&g->row_info);
        mbar_wait_addr(db + SD_OFF(full_in) + 16u * slot + 8u * buf, (k >> 1) & 1u);
        const int W = (int)info.x, H = (int)info.y;
        const bool has_dem = (info.w & TSF_DEM) != 0u, has_land = (info.w & TSF_LAND) != 0u, has_ocean = (info.w & TSF_OCEAN) != 0u;
        if (item.tile != cur_tile) {
            flush_counters();                                 // the registers hold counts of the previous tile
            cur_tile = item.tile;
        }
        const int x0 = (int)item.tx * FT_W;
        const int ly = 4 * (int)r + (int)c, y = (int)item.ty * SD_H + ly, x = x0 + 4 * lane;
        const bool active = y < H && x < W;
        bool dem_ready = false;
        if (active) {
            const uint32_t xe = c * (uint32_t)W + (uint32_t)x0;
            const uint32_t base = db + SD_OFF(in) + slot * (uint32_t)sizeof(DynSlot);
            const uint32_t ab = base + r * (2u * ST_BAND_W) + 2u * (xe & 7u) + 8u * (uint32_t)lane;
            const uint32_t ay = base + 6u * SD_BAND_BYTES + r * ST_BYTE_W + (xe & 15u) + 4u * (uint32_t)lane;
#pragma unroll
            for (int kk = 0; kk < 6; ++kk) {
                uint32_t v0, v1;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(ab + (uint32_t)kk * SD_BAND_BYTES));
                w[kk][0] = v0; w[kk][1] = v1;
            }
            fm4 = lds_u32(ay);
            ld4 = has_land ? lds_u32(ay + SD_BYTE_BYTES) : 0xffffffffu;
            oc4 = has_ocean ? lds_u32(ay + 2u * SD_BYTE_BYTES) : 0x01010101u;
            // the row is IN REGISTERS (not merely requested) before the slot is handed back to the producer
            asm volatile("" ::"r"(w[5][1]), "r"(fm4), "r"(ld4), "r"(oc4) : "memory");
        }
        __syncwarp();
        mbar_arrive_elected(db + SD_OFF(empty_in) + 8u * slot);
        if (active) {
            const uint32_t pix = (uint32_t)y * (uint32_t)W + (uint32_t)x;
#define FT_DEM_WAIT() mbar_wait_addr(db + SD_OFF(full) + 8u * buf, (k >> 1) & 1u)
#define FT_ROW_MIDPOINT() do { } while (0)
#define FT_PADX() ((int)info.z)
#define FT_OUT_PTRS(pd, pw, pb, pc)                                                                   \
    do {                                                                                              \
        const ulonglong2 p01 = __ldg(reinterpret_cast<const ulonglong2 *>(g->out_ptrs)),              \
                         p23 = __ldg(reinterpret_cast<const ulonglong2 *>(g->out_ptrs) + 1);          \
        pd = reinterpret_cast<uint16_t *>(p01.x); pw = reinterpret_cast<uint8_t *>(p01.y);            \
        pb = reinterpret_cast<uint8_t *>(p23.x); pc = reinterpret_cast<uint8_t *>(p23.y);             \
    } while (0)
#define FT_SUN4(i) __ldg(reinterpret_cast<const float4 *>(g->sun32) + (i))
#define FT_EXACT4(am) shadow_exact4_global(am, &g->tile, P)
#define FT_DEM_BASE (db + SD_OFF(dem) + buf * (uint32_t)sizeof(DynDem))
#include "pb200_fused_row.inc"
#undef FT_DEM_BASE
#undef FT_EXACT4
#undef FT_SUN4
#undef FT_OUT_PTRS
#undef FT_PADX
#undef FT_ROW_MIDPOINT
#undef FT_DEM_WAIT
        }
        const uint32_t t_next = take_ticket();
        __syncwarp();
        mbar_arrive_elected(db + SD_OFF(empty) + 8u * buf);
        t = t_next;
    }
    flush_counters();
    (void)acc_hist;
}

}  // namespace pb200
