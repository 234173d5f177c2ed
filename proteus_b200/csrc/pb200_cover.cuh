// pb200_cover.cuh - mask_adjacent_to_cloud_mode = 'cover' (SURVEY 8f "next #2").
//
// _add_snow_to_cloud_layer, branch D:2055-2078: the snow mask is dilated 10 times
// over pixels flagged "adjacent to cloud/shadow" whose CLOUD value is still 0, then
// the not-masked area is dilated back 7 times over the part of that area WTR-2 calls
// water.  scipy.ndimage.binary_dilation(input, iterations=k, mask=m) with the default
// structure (4-connected cross incl. the centre, border_value 0): at every iteration
//     new[p] = m[p] ? (old[p] | old[up] | old[down] | old[left] | old[right]) : old[p]
// One launch per iteration on ping-pong byte rasters (17 launches of ~3 B/px each).
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

}  // namespace pb200
