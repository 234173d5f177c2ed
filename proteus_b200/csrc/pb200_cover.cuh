// pb200_cover.cuh - mask_adjacent_to_cloud_mode = 'cover' (SURVEY 8f "next #2").
//
// _add_snow_to_cloud_layer, branch D:2055-2078: the snow mask is dilated 10 times
// over pixels flagged "adjacent to cloud/shadow" whose CLOUD value is still 0, then
// the not-masked area is dilated back 7 times over the part of that area WTR-2 calls
// water.  scipy.ndimage.binary_dilation(input, iterations=k, mask=m) with the default
// structure (4-connected cross incl. the centre, border_value 0): at every iteration
//     new[p] = m[p] ? (old[p] | old[up] | old[down] | old[left] | old[right]) : old[p]
// Up to MD_HALO iterations run inside ONE launch: a CTA stages its tile plus a halo as wide as the iteration count
// in shared memory (an iteration moves information by one pixel), iterates there and writes the tile centre.
#pragma once
#include "pb200_device.cuh"

namespace pb200 {

// snow = fmask & 16 ; area = (fmask & 4) && cloud == 0          (D:2052, D:2057-2058)
__global__ void cover_init_kernel(const uint8_t *__restrict__ fmask, const uint8_t *__restrict__ cloud,
                                  uint8_t *__restrict__ snow, uint8_t *__restrict__ area, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t f = fmask[i];
        snow[i] = (f & 16u) ? 1 : 0;
        area[i] = ((f & 4u) && cloud[i] == 0) ? 1 : 0;
    }
}

// one masked dilation step (scipy semantics above)
__global__ void masked_dilation_step_kernel(const uint8_t *__restrict__ in, const uint8_t *__restrict__ mask,
                                            uint8_t *__restrict__ out, int rows, int cols) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const size_t i = (size_t)y * cols + x;
    uint32_t v = in[i];
    if (mask[i] && !v) {
        if (x > 0) v |= in[i - 1];
        if (x + 1 < cols) v |= in[i + 1];
        if (y > 0) v |= in[i - cols];
        if (y + 1 < rows) v |= in[i + cols];
    }
    out[i] = (uint8_t)(v ? 1 : 0);
}

// `iters` (<= MD_HALO) masked dilation steps of a tile in shared memory, ONE BIT PER PIXEL: a 32-bit word holds 32
// consecutive pixels of a row, a step is  v | (m & (v << 1 | v >> 1 | up | down))  with the carries of the
// neighbouring words - about 0.4 instructions per pixel and step.  The staged window is 256 x 128 pixels (8 x 128
// words); pixels outside the raster are 0 with mask 0 (scipy border_value = 0).  After k steps only the pixels at
// least k away from the window's edge are exact: the halo is MD_HALO wide on every side, the 224 x 96 centre is exact
// after up to MD_HALO steps.
constexpr int MD_HALO = 16;
constexpr int MD_WW = 8, MD_SW = 32 * MD_WW, MD_SH = 128;                   // staged window: 8 words x 128 rows
constexpr int MD_TW = MD_SW - 2 * MD_HALO, MD_TH = MD_SH - 2 * MD_HALO;     // 224 x 96 output pixels per CTA
__global__ void __launch_bounds__(256) masked_dilation_tiled_kernel(const uint8_t *__restrict__ in,
                                                                     const uint8_t *__restrict__ mask,
                                                                     uint8_t *__restrict__ out, int rows, int cols,
                                                                     int iters) {
    __shared__ uint32_t img[2][MD_SH + 2][MD_WW + 2];     // one guard row / word column of zeros all around
    __shared__ uint32_t msk[MD_SH][MD_WW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * MD_TW - MD_HALO, y0 = blockIdx.y * MD_TH - MD_HALO;
    for (int i = threadIdx.x; i < 2 * (MD_SH + 2) * (MD_WW + 2); i += 256) (&img[0][0][0])[i] = 0u;
    __syncthreads();
    // stage: a warp packs 32 consecutive pixels of a row into one word with a ballot
    for (int wi = warp; wi < MD_SH * MD_WW; wi += 8) {
        const int ly = wi / MD_WW, lw = wi - ly * MD_WW;
        const int gy = y0 + ly, gx = x0 + 32 * lw + lane;
        const bool inside = gy >= 0 && gy < rows && gx >= 0 && gx < cols;
        const size_t g = (size_t)(inside ? gy : 0) * cols + (inside ? gx : 0);
        const uint32_t vb = __ballot_sync(0xffffffffu, inside && in[g] != 0);
        const uint32_t mb = __ballot_sync(0xffffffffu, inside && mask[g] != 0);
        if (lane == 0) { img[0][ly + 1][lw + 1] = vb; msk[ly][lw] = mb; }
    }
    __syncthreads();
    int cur = 0;
    for (int it = 0; it < iters; ++it) {
        for (int wi = threadIdx.x; wi < MD_SH * MD_WW; wi += 256) {
            const int ly = wi / MD_WW, lw = wi - ly * MD_WW;
            const uint32_t v = img[cur][ly + 1][lw + 1];
            const uint32_t nb = (v << 1) | (img[cur][ly + 1][lw] >> 31) | (v >> 1) | (img[cur][ly + 1][lw + 2] << 31) |
                                img[cur][ly][lw + 1] | img[cur][ly + 2][lw + 1];
            img[cur ^ 1][ly + 1][lw + 1] = v | (msk[ly][lw] & nb);
        }
        __syncthreads();
        cur ^= 1;
    }
    // unpack the centre: a warp writes 32 consecutive pixels of a row
    for (int wi = warp; wi < MD_TH * (MD_TW / 32); wi += 8) {
        const int ty = wi / (MD_TW / 32), tw = wi - ty * (MD_TW / 32);
        const int px = MD_HALO + 32 * tw + lane;                             // pixel column inside the window
        const uint32_t v = img[cur][ty + MD_HALO + 1][(px >> 5) + 1];
        const int gy = blockIdx.y * MD_TH + ty, gx = blockIdx.x * MD_TW + 32 * tw + lane;
        if (gy < rows && gx < cols) out[(size_t)gy * cols + gx] = (uint8_t)((v >> (px & 31)) & 1u);
    }
}

// area &= wtr2 in 1..4 ; not_masked = !snow && cloud == 0       (D:2070-2074)
__global__ void cover_mid_kernel(const uint8_t *__restrict__ snow, const uint8_t *__restrict__ cloud,
                                 const uint8_t *__restrict__ wtr2, uint8_t *__restrict__ area,
                                 uint8_t *__restrict__ not_masked, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t w = wtr2[i];
        area[i] = (area[i] && w >= 1u && w <= 4u) ? 1 : 0;
        not_masked[i] = (!snow[i] && cloud[i] == 0) ? 1 : 0;
    }
}

// snow[not_masked] = False ; cloud[snow] += 2 ; cloud[wtr2 == 255] = 255   (D:2078-2084)
__global__ void cover_final_kernel(const uint8_t *__restrict__ snow, const uint8_t *__restrict__ not_masked,
                                   const uint8_t *__restrict__ wtr2, uint8_t *__restrict__ cloud, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t c = cloud[i];
        if (snow[i] && !not_masked[i]) c = (c + 2u) & 255u;
        if (wtr2[i] == 255u) c = 255u;
        cloud[i] = (uint8_t)c;
    }
}

// point-wise tail of the cover flow in one pass (D:2089-2133, 1710-1730, 1733-1837, 2578-2598): WTR, BWTR, CONF from
// WTR-2 and the final CLOUD; WTR / WTR-1 / WTR-1-remapped / WTR-2 collapsed in place when asked (null = not wanted)
__global__ void cover_tail_kernel(uint8_t *__restrict__ wtr2, const uint8_t *__restrict__ cloud,
                                  uint8_t *__restrict__ wtr, uint8_t *__restrict__ bwtr, uint8_t *__restrict__ conf,
                                  uint8_t *__restrict__ wtr1, uint8_t *__restrict__ wtr1r, int collapse, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t w2 = wtr2[i], c = cloud[i];
        const uint32_t w = cloud_masking(w2, c);
        if (bwtr) bwtr[i] = (uint8_t)binary_water(w);
        if (conf) conf[i] = (uint8_t)confidence(w2, c);
        if (wtr) wtr[i] = (uint8_t)(collapse ? collapse_class(w) : w);
        if (collapse) {
            wtr2[i] = (uint8_t)collapse_class(w2);
            if (wtr1) wtr1[i] = (uint8_t)collapse_class(wtr1[i]);
            if (wtr1r) wtr1r[i] = (uint8_t)collapse_class(wtr1r[i]);
        }
    }
}

}  // namespace pb200
