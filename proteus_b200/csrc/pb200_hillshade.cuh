// pb200_hillshade.cuh - SURVEY 8f "next #3": the hillshade of the 'otsu' shadow algorithm (D:4177-4212, called at
// D:5152-5155) fused with the 256-bin count that _compute_otsu_threshold (D:1638-1684) starts from.
//
// PARITY UNPINNED.  The reference obtains the hillshade from GDAL (gdal.DEMProcessing(..., "hillshade", azimuth,
// altitude), GDAL 3.6.2, un-vendored: its arithmetic is not under /root/reference and GDAL is absent from the build and
// GPU images).  What is restated here - and in the test oracle's compute_hillshade_gdal (oracle/), which this kernel matches
// bit for bit - is the published gdaldem algorithm with gdaldem's defaults (Horn gradient, z = 1, scale = 1, no
// -compute_edges, Byte output):
//     x = ((w0 + w3 + w3 + w6) - (w2 + w5 + w5 + w8)) / ewres        window sums in float32 (the band's type),
//     y = ((w6 + w7 + w7 + w8) - (w0 + w1 + w1 + w2)) / nsres        w0..w2 = northern row; nsres < 0
//     c = (254 sin(alt) - (y 254 cos(az) cos(alt) z/8 - x 254 sin(az) cos(alt) z/8)) / sqrt(1 + (z/8)^2 (x^2 + y^2))
//     shade = c <= 0 ? 1 : 1 + c      -> float32 -> Byte (round half up);  first / last row and column: 0 (no data)
// GDAL builds with SSE2 replace the division by an approximate reciprocal square root with one Newton step; a shade
// within 1e-6 of a rounding boundary may then differ by one grey level from this restatement.
//
// One pass: the DEM tile of a CTA (128 x 32 pixels + 1-pixel halo) is staged by TMA, a lane computes 4 consecutive
// pixels of 4 rows, writes the bytes and counts them in its warp's private shared-memory histogram.
#pragma once
#include "pb200_kernels.cuh"

namespace pb200 {

struct HillParams {
    double inv_ewres, inv_nsres;
    double sin_alt_254, cos_az_cos_alt_z_254, sin_az_cos_alt_z_254, square_z;
};

__device__ __forceinline__ uint32_t hillshade_pixel(const float (&w)[9], const HillParams &H) {
    // float32 sums, left to right, as `afWin[0] + afWin[3] + afWin[3] + afWin[6]` evaluates for a float band
    const float xs = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(w[0], w[3]), w[3]), w[6]),
                               __fadd_rn(__fadd_rn(__fadd_rn(w[2], w[5]), w[5]), w[8]));
    const float ys = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(w[6], w[7]), w[7]), w[8]),
                               __fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), w[1]), w[2]));
    const double x = __dmul_rn((double)xs, H.inv_ewres), y = __dmul_rn((double)ys, H.inv_nsres);
    const double xx_plus_yy = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double num = __dsub_rn(H.sin_alt_254, __dsub_rn(__dmul_rn(y, H.cos_az_cos_alt_z_254), __dmul_rn(x, H.sin_az_cos_alt_z_254)));
    const double c254 = __ddiv_rn(num, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(H.square_z, xx_plus_yy))));
    const double cang = c254 <= 0.0 ? 1.0 : __dadd_rn(1.0, c254);
    const float f = __fadd_rn((float)cang, 0.5f);                    // GDALCopyWords float -> Byte: + 0.5, clamp, truncate
    if (!(f > 0.0f)) return 0u;                                      // also NaN
    return f >= 255.0f ? 255u : (uint32_t)f;
}

template <bool USE_TMA>
__global__ void __launch_bounds__(NTHREADS) hillshade_hist_kernel(const float *__restrict__ dem, const __grid_constant__ CUtensorMap tmap,
                                                                  int rows, int cols, uint8_t *__restrict__ out,
                                                                  unsigned long long *__restrict__ counts,
                                                                  const __grid_constant__ HillParams H) {
    __shared__ __align__(128) float tile[SMH][SMW];
    __shared__ unsigned int hist[NWARPS][256];
    __shared__ unsigned long long mbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (USE_TMA) {
        if (tid == 0) {
            mbar_init(&mbar, 1);
            mbar_expect_tx(&mbar, (uint32_t)sizeof(tile));
            // box start x0 - 4: a multiple of 4 floats (16-byte aligned); elements outside the raster arrive as zeros
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                    "r"(smem_u32(&tile[0][0])), "l"(&tmap), "r"(x0 - DEM_PADX), "r"(y0 - 1), "r"(smem_u32(&mbar))
                : "memory");
        }
    } else {
        for (int i = tid; i < SMH * SMW; i += NTHREADS) {
            const int r = i / SMW, c = i - r * SMW;
            const int gy = y0 - 1 + r, gx = x0 - DEM_PADX + c;
            tile[r][c] = (gy >= 0 && gy < rows && gx >= 0 && gx < cols) ? __ldg(dem + (size_t)gy * cols + gx) : 0.0f;
        }
    }
    if (counts)
        for (int i = tid; i < NWARPS * 256; i += NTHREADS) (&hist[0][0])[i] = 0u;
    __syncthreads();
    if (USE_TMA) mbar_wait(&mbar, 0);

    const int x = x0 + 4 * lane;
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        const int ly = warp * ROWS_PER_WARP + rr, y = y0 + ly;
        if (y >= rows || x >= cols) continue;
        const int sc = 4 * lane + DEM_PADX;                           // shared-memory column of this lane's first pixel
        uint32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gx = x + j;
            if (gx >= cols) { b[j] = 0u; continue; }
            if (y == 0 || y == rows - 1 || gx == 0 || gx == cols - 1) { b[j] = 0u; continue; }   // edges: no data (0)
            const float w[9] = {tile[ly][sc + j - 1],     tile[ly][sc + j],     tile[ly][sc + j + 1],
                                tile[ly + 1][sc + j - 1], tile[ly + 1][sc + j], tile[ly + 1][sc + j + 1],
                                tile[ly + 2][sc + j - 1], tile[ly + 2][sc + j], tile[ly + 2][sc + j + 1]};
            b[j] = hillshade_pixel(w, H);
        }
        uint8_t *o = out + (size_t)y * cols + x;
        const int nv = min(4, cols - x);
        if (nv == 4 && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
            *reinterpret_cast<uint32_t *>(o) = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        } else {
            for (int j = 0; j < nv; ++j) o[j] = (uint8_t)b[j];
        }
        if (counts)
            for (int j = 0; j < nv; ++j) atomicAdd(&hist[warp][b[j]], 1u);
    }
    if (counts) {
        __syncthreads();
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) t += hist[w][tid];
        if (t) atomicAdd(&counts[tid], (unsigned long long)t);
    }
}

}  // namespace pb200
