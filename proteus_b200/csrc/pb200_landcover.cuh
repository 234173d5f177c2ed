// pb200_landcover.cuh - SURVEY 8f "next #1": the numpy tail of create_landcover_mask
// (D:1003-1115) as one pass: ESA WorldCover warped to 10 m (3 x 3 samples per product pixel) +
// CGLS land cover on the product grid -> LAND classes.
//
// Per product pixel: counts of water {80, 90, 95}, urban 50 and tree 10 samples in its 3 x 3
// block (decimate_by_summation, D:874-904), tree count kept only on CGLS forest classes
// (D:1033-1043), then the threshold hierarchy (D:1045-1115), later rules overriding earlier ones.
// HBM bound: 9 + 1 bytes in, 1 byte out per product pixel.  A thread produces 4 product pixels of
// a row: three 12-byte row segments of the 10 m raster arrive as 3 x 3 aligned 4-byte loads.
#pragma once
#include "pb200_device.cuh"

namespace pb200 {

struct LandParams {
    int32_t thr_tree, thr_low, thr_high, thr_water;   // landcover_threshold_dict (D:270-271)
    uint32_t cls_tree, cls_low, cls_high, cls_water;  // 201, year_offset, 100 + year_offset, 200
    uint32_t forest_bits[8];                          // bit v set: CGLS class v is a forest class
};

// sample value -> 1 (water) | 1 << 8 (urban) | 1 << 16 (tree): summing nine of them counts all three
__device__ __forceinline__ uint32_t worldcover_code(uint32_t v) {
    return ((v == 80u || v == 90u || v == 95u) ? 1u : 0u) | (v == 50u ? 0x100u : 0u) | (v == 10u ? 0x10000u : 0u);
}

__device__ __forceinline__ uint32_t land_class(uint32_t counts, uint32_t copernicus, const LandParams &L) {
    const int water = counts & 255u, urban = (counts >> 8) & 255u;
    int tree = (counts >> 16) & 255u;
    if (!((L.forest_bits[copernicus >> 5] >> (copernicus & 31u)) & 1u)) tree = 0;     // D:1043
    uint32_t c = 255u;                                                                 // D:1050
    if (tree >= L.thr_tree) c = L.cls_tree;                                            // D:1062
    if (urban >= L.thr_low) c = L.cls_low;                                             // D:1099
    if (urban >= L.thr_high) c = L.cls_high;                                           // D:1106
    if (water >= L.thr_water) c = L.cls_water;                                         // D:1112
    return c;
}

// rows x cols = product grid; worldcover is (3 rows) x (3 cols)
template <bool VEC>
__global__ void landcover_aggregate_kernel(const uint8_t *__restrict__ worldcover, const uint8_t *__restrict__ copernicus,
                                           uint8_t *__restrict__ land, int rows, int cols,
                                           const __grid_constant__ LandParams L) {
    __shared__ uint32_t code[256];
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y) code[i] = worldcover_code(i);
    __syncthreads();
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= cols) return;
    const size_t wpitch = (size_t)cols * 3;
    // a CTA walks down the raster (grid.y is sized for ~8 CTAs per SM): the class-code table is built once per CTA, not
    // once per 1024 pixels (that was 30 % of the executed instructions, profiles/)
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < rows; y += gridDim.y * blockDim.y) {
    uint32_t counts[4] = {0u, 0u, 0u, 0u};
    if (VEC) {
        // cols % 4 == 0 and 4-byte aligned planes: 12 bytes per 10 m row = 3 aligned words
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(worldcover + ((size_t)y * 3 + r) * wpitch + (size_t)x4 * 3);
            const uint32_t w0 = ldg_stream_u32(p), w1 = ldg_stream_u32(p + 1), w2 = ldg_stream_u32(p + 2);
            const uint32_t b[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                    w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                    w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll
            for (int j = 0; j < 4; ++j) counts[j] += code[b[3 * j]] + code[b[3 * j + 1]] + code[b[3 * j + 2]];
        }
        const uint32_t c4 = ldg_stream_u32(copernicus + (size_t)y * cols + x4);
        uint32_t out = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) out |= land_class(counts[j], (c4 >> (8 * j)) & 255u, L) << (8 * j);
        stg_stream_u32(land + (size_t)y * cols + x4, out);
    } else {
        for (int j = 0; j < 4 && x4 + j < cols; ++j) {
            uint32_t cnt = 0u;
            for (int r = 0; r < 3; ++r) {
                const uint8_t *p = worldcover + ((size_t)y * 3 + r) * wpitch + (size_t)(x4 + j) * 3;
                cnt += code[p[0]] + code[p[1]] + code[p[2]];
            }
            land[(size_t)y * cols + x4 + j] = (uint8_t)land_class(cnt, copernicus[(size_t)y * cols + x4 + j], L);
        }
    }
    }
}

}  // namespace pb200
