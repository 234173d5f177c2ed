/*
 * proteus_b200.h - C ABI of libproteus_b200.so
 *
 * B200-native (sm_100a) implementation of the per-pixel DSWx-HLS
 * classification path of nasa/PROTEUS v1.0.2.  The reference has no FFI layer
 * of its own (it is pure Python/numpy): the seam this library replaces is the
 * set of module-level functions of src/proteus/dswx_hls.py ("D:" below) that
 * generate_dswx_layers (D:4610) calls between loading the rasters and saving
 * the layers.  Each entry point cites the reference statement(s) it replaces.
 *
 * Conventions
 *   - plain C types only; every raster pointer is a DEVICE pointer unless the
 *     function name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, counters); the library
 *     allocates nothing per call except inside a plan / the host pipeline;
 *   - every function returns 0 on success, <0 for an invalid argument
 *     (PB200_E_*), >0 for a cudaError_t; pb200_last_error() gives the text
 *     (thread-local);
 *   - calls are asynchronous on the given CUDA stream (a cudaStream_t passed
 *     as void*; NULL = legacy default stream) unless stated otherwise;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with a cudaError_t.
 */
#ifndef PROTEUS_B200_H
#define PROTEUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 1

/* error codes (<0) */
#define PB200_E_INVALID_ARG   (-1)
#define PB200_E_BAD_MODE      (-2)   /* D:1977-1981 raise Exception(...) */
#define PB200_E_UNSUPPORTED   (-3)
#define PB200_E_NO_DRIVER_API (-4)
#define PB200_E_ALIGNMENT     (-5)
#define PB200_E_NCCL          (-6)   /* NCCL missing or an ncclResult_t != ncclSuccess (text in pb200_last_error) */

/* "this raster has no fill value" (D:2196-2199 always yields one, but the
 * function-granular entry points are also used on already-clean data) */
#define PB200_NO_FILL INT32_MIN

/* mask_adjacent_to_cloud_mode (D:1929-1935) */
#define PB200_ADJ_MASK   0
#define PB200_ADJ_IGNORE 1
#define PB200_ADJ_COVER  2   /* D:2055-2078: pb200_snow_to_cloud_cover + pb200_masked_dilation; the fused
                                entry points run it as fused pass (defer_snow) -> dilations -> final pass */

/* counters slot layout (uint64 each) */
#define PB200_CNT_VALID           0  /* D:5110  n_valid                     */
#define PB200_CNT_CLOUD_AND_VALID 1  /* D:5111  n_cloud_and_valid           */
#define PB200_CNT_NOT_OCEAN       2  /* D:5105  sum(ocean_mask)             */
#define PB200_CNT_CLASS0          3  /* extension: histogram of the UNCOLLAPSED WTR layer,
                                        9 bins in the order 0,1,2,3,4,252,253,254,255 */
#define PB200_N_COUNTERS          12

/* Mirror of class HlsThresholds (D:274-318); values as in the runconfig
 * (defaults: src/proteus/defaults/dswx_hls.yaml:176-212). */
typedef struct pb200_thresholds {
    double wigt, awgt;
    double pswt_1_mndwi, pswt_1_nir, pswt_1_swir1, pswt_1_ndvi;
    double pswt_2_mndwi, pswt_2_blue, pswt_2_nir, pswt_2_swir1, pswt_2_swir2;
    double lcmask_nir;
} pb200_thresholds;

/* Everything generate_dswx_layers passes down the hot path besides rasters. */
typedef struct pb200_params {
    pb200_thresholds th;
    int32_t band_fill[6];        /* D:2176/2196-2199, blue..swir2; PB200_NO_FILL = none */
    int32_t fmask_fill;          /* HLS v2 Fmask: 255 */
    int32_t adjacent_mode;       /* PB200_ADJ_* */
    int32_t apply_aerosol_class_remapping;       /* D:5260 */
    /* bit k of aerosol_class_bits[v] is set when Fmask value v is in the
     * list that remaps WTR-1 class k to 1 (k = 0, 2, 3, 4; D:1283-1296). */
    uint8_t aerosol_class_bits[256];
    double  min_slope_angle;           /* degrees, D:4279 */
    double  max_sun_local_inc_angle;   /* degrees, D:4280 */
    /* Decision thresholds of the two angle tests in the cosine / tangent
     * domain:  degrees(arccos(x)) <= max_inc  <=>  x >= cos_inc_threshold,
     *          degrees(arctan(s)) <= min_slope <=> s <= tan_slope_threshold.
     * Set both to NaN to let the library derive them with libm; the Python
     * host derives them by bisection on numpy's own arccos/arctan so that the
     * boundary is the reference's boundary to the last bit. */
    double  cos_inc_threshold;
    double  tan_slope_threshold;
    double  pixel_spacing_x, pixel_spacing_y;    /* D:4217: always 30, 30 */
    int32_t collapse_wtr_classes;  /* write WTR / WTR-1 / WTR-2 collapsed (D:2688-2689) */
    int32_t class_histogram;       /* also fill counters[3..11] */
    /* 1: leave the Fmask snow bit out of CLOUD (and of WTR / CONF): first phase of the 'cover' mode,
     * where the snow mask is dilated (pb200_snow_to_cloud_cover) before it is added (D:2055-2081). */
    int32_t defer_snow;
    /* 1: the terrain-shadow test (D:4264-4281) with numpy 1.x value-based casting - the reference pins numpy 1.23.5
     * (setup.py:78), where float32_array * float64_scalar stays float32: sun terms rounded to float32, dot product,
     * division, arccos / arctan / degrees in float32; the two thresholds above then hold float32 values.
     * 0: numpy >= 2 promotion (float64 from the dot product on).  Only float32 DEMs are affected. */
    int32_t numpy1_promotion;
} pb200_params;

/* One raster tile (an MGRS tile, one acquisition of a time series, or one row
 * strip of a mosaic).  Inputs are planar, C-contiguous, row pitch = width.
 * NULL outputs are skipped.  The DEM is addressed as
 *   dem[(dem_off_y + row) * dem_pitch + (dem_off_x + col)]
 * and must have at least one valid element on every side of the tile
 * (the reference warps it with a 50-px margin, D:58, D:5145-5150). */
typedef struct pb200_tile {
    int32_t height, width;
    const int16_t *band[6];      /* RAW blue, green, red, nir, swir1, swir2 (before D:2299) */
    const uint8_t *fmask;
    const float   *dem;          /* NULL: no terrain-shadow masking */
    int32_t dem_pitch;           /* elements per DEM row */
    int32_t dem_rows;            /* rows in the DEM array */
    int32_t dem_off_y, dem_off_x;
    const uint8_t *land;         /* NULL: no land-cover masking (D:1346) */
    const uint8_t *ocean;        /* NULL: no shoreline given (D:5096) */
    double sun_azimuth, sun_elevation;   /* degrees (D:5044-5059) */
    /* Optional: the five float64 scalars D:4245-4252, 4276-4277 derive from
     * the two angles, in this order: sun_x, sun_y, sun_z, sin(az), cos(az).
     * sun_terms[0] = NaN: the library derives them with libm.  The Python host
     * fills them with numpy so that they are the reference's own values. */
    double sun_terms[5];
    uint16_t *diag;              /* DIAG  (D:5231) */
    uint8_t  *wtr1;              /* WTR-1 as saved, i.e. BEFORE aerosol remapping (D:5251) */
    uint8_t  *wtr1_remapped;     /* WTR-1 after D:5261 (band of the combined file, D:5390) */
    uint8_t  *wtr2;              /* WTR-2 (D:5268) */
    uint8_t  *cloud;             /* CLOUD (D:5282) */
    uint8_t  *shad;              /* SHAD, 1 = not shadow (D:5166) */
    uint8_t  *wtr;               /* WTR   (D:5286) */
    uint8_t  *bwtr;              /* BWTR  (D:5358) */
    uint8_t  *conf;              /* CONF  (D:5368) */
    uint64_t *counters;          /* PB200_N_COUNTERS slots, ADDED to (caller zeroes) */
} pb200_tile;

typedef struct pb200_ctx  pb200_ctx;
typedef struct pb200_plan pb200_plan;

/* ---- library / context ------------------------------------------------ */
int         pb200_version(void);
const char *pb200_last_error(void);
int  pb200_ctx_create(int device, pb200_ctx **out);
int  pb200_ctx_destroy(pb200_ctx *ctx);
/* defaults of src/proteus/defaults/dswx_hls.yaml:64-109,176-212 */
int  pb200_params_default(pb200_params *p);

/* ---- fused path: D:5088-5369 in one pass -------------------------------- */
/* One launch over n_tiles tile descriptors (host array of device pointers). */
int  pb200_classify(pb200_ctx *ctx, const pb200_tile *tiles, int n_tiles,
                    const pb200_params *params, void *stream);
/* Plan = descriptors + TMA tensor maps resident on the device; run many times
 * (time series, benchmarks, CUDA-graph capture). */
int  pb200_plan_create(pb200_ctx *ctx, const pb200_tile *tiles, int n_tiles,
                       const pb200_params *params, pb200_plan **out);
int  pb200_plan_run(pb200_plan *plan, void *stream);
/* Which kernels the plan launches (tests, benchmarks): bit 0 fused kernel with direct loads, bit 1 fused kernel with
 * TMA-fed inputs (tile height % 4 == 0, width >= 36, 16-byte aligned planes, the four graded layers + counters), bit 2
 * the FAST8 flavour of either (parameters shaped like the defaults), bit 3 the generic kernel for some tile (width % 4
 * != 0, unaligned planes, DEM not TMA-addressable). */
#define PB200_KERNEL_FAST    1
#define PB200_KERNEL_STREAM  2
#define PB200_KERNEL_FAST8   4
#define PB200_KERNEL_GENERIC 8
#define PB200_KERNEL_STREAM_DYN 16   /* with PB200_KERNEL_STREAM: dswx_fused_stream_dyn_kernel (rows handed to free warps, deep ring) */
int  pb200_plan_kernels(const pb200_plan *plan, int *mask);
int  pb200_plan_destroy(pb200_plan *plan);
/* Same contract as pb200_classify for ONE tile whose pointers are HOST
 * pointers (what generate_dswx_layers holds after gdal.ReadAsArray, D:2192).
 * Row strips are copied in, classified and copied out on three streams so
 * H2D, the kernel and D2H overlap; returns when the outputs are in host
 * memory.  Buffers from pb200_host_alloc (pinned) copy at full PCIe rate. */
int  pb200_classify_host(pb200_ctx *ctx, const pb200_tile *host_tile,
                         const pb200_params *params, int strip_rows);
/* Time series (BASELINE configs[3]: many acquisitions of one MGRS tile share DEM, LAND and ocean mask): with
 * PB200_HOST_REUSE_ANCILLARY the DEM / LAND / ocean rasters uploaded by the previous successful pb200_classify_host*
 * call on this context are used again and not copied - only the bands and Fmask cross PCIe (174 MB instead of 256 MB
 * per HLS tile).  The tile must have the same size, DEM geometry and the same rasters present as that call
 * (PB200_E_INVALID_ARG otherwise); that their CONTENT is unchanged is the caller's statement.  flags = 0 is
 * pb200_classify_host. */
#define PB200_HOST_REUSE_ANCILLARY 1
/* Two tiles in flight.  The context holds two independent host pipelines ("slots": device mirror, streams, arena);
 * PB200_HOST_SLOT1 selects the second.  With PB200_HOST_ASYNC the call returns as soon as the tile's copies and
 * kernels are enqueued; pb200_host_wait(ctx, slot) returns when its outputs are in host memory (the next call on the
 * same slot waits implicitly).  A caller that alternates the slots - enqueue tile k + 1, then wait for tile k - keeps
 * the host-to-device engine busy during the kernel / device-to-host tail of the previous tile.  Every buffer of a
 * tile must stay untouched until its wait has returned.  PB200_HOST_REUSE_ANCILLARY refers to the rasters resident in
 * the SAME slot. */
#define PB200_HOST_ASYNC 2
#define PB200_HOST_SLOT1 4
int  pb200_classify_host_ex(pb200_ctx *ctx, const pb200_tile *host_tile,
                            const pb200_params *params, int strip_rows, int flags);
int  pb200_host_wait(pb200_ctx *ctx, int slot);
int  pb200_host_alloc(size_t bytes, void **out);
int  pb200_host_free(void *p);

/* ---- function-granular entry points (parity with one reference function) */
/* D:2203-2209 + D:2298-2299: invalid |= raw == fill ; band = max(band, 1) */
int  pb200_invalid_and_clip(pb200_ctx *ctx, const int16_t *const raw[6],
                            const uint8_t *fmask, const pb200_params *params,
                            int64_t n, int16_t *const clipped[6],
                            uint8_t *invalid, void *stream);
/* D:1840-1916 _compute_diagnostic_tests (int16 bands, used as given) */
int  pb200_diagnostic_tests(pb200_ctx *ctx, const int16_t *const band[6],
                            const pb200_thresholds *th, int64_t n,
                            uint16_t *diag_decimal, void *stream);
/* D:1840-1916 on FLOAT32 bands (the --offset-and-scale-inputs mode, D:2300-2302): every operation in
 * float32 in numpy's order, no FMA contraction, thresholds cast to float32 */
int  pb200_diagnostic_tests_f32(pb200_ctx *ctx, const float *const band[6],
                                const pb200_thresholds *th, int64_t n,
                                uint16_t *diag_decimal, void *stream);
/* D:1687-1707 generate_interpreted_layer */
int  pb200_interpreted_layer(pb200_ctx *ctx, const uint16_t *diag_decimal,
                             int64_t n, uint8_t *wtr1, void *stream);
/* D:4286-4317 _get_binary_representation (nbits = 6) */
int  pb200_binary_representation(pb200_ctx *ctx, const uint16_t *diag_decimal,
                                 int64_t n, uint16_t *diag, void *stream);
/* D:1919-1993 _compute_preliminary_cloud_layer */
int  pb200_preliminary_cloud(pb200_ctx *ctx, const uint8_t *fmask, int mode,
                             int64_t n, uint8_t *cloud, void *stream);
/* D:1249-1302 _apply_aerosol_class_remapping: wtr1 and cloud IN PLACE */
int  pb200_aerosol_remap(pb200_ctx *ctx, uint8_t *wtr1, const int16_t *nir,
                         uint8_t *cloud, const uint8_t *fmask,
                         const uint8_t aerosol_class_bits[256], int64_t n,
                         void *stream);
/* D:1305-1378 _apply_landcover_and_shadow_masks; land / shad may be NULL */
int  pb200_landcover_shadow_masks(pb200_ctx *ctx, const uint8_t *wtr1,
                                  const int16_t *nir, const uint8_t *land,
                                  const uint8_t *shad, double lcmask_nir,
                                  int64_t n, uint8_t *wtr2, void *stream);
/* D:1996-2086 _add_snow_to_cloud_layer, modes mask / ignore; cloud IN PLACE */
int  pb200_snow_to_cloud(pb200_ctx *ctx, const uint8_t *wtr2, uint8_t *cloud,
                         const uint8_t *fmask, int mode, int64_t n,
                         void *stream);
/* D:2055-2084 _add_snow_to_cloud_layer, mode 'cover': masked dilation of the snow mask (10 x) and of
 * the not-masked area (7 x), scipy.ndimage.binary_dilation semantics; cloud IN PLACE.
 * scratch: 4 * rows * cols bytes of device memory. */
int  pb200_snow_to_cloud_cover(pb200_ctx *ctx, const uint8_t *wtr2, uint8_t *cloud,
                               const uint8_t *fmask, int rows, int cols,
                               uint8_t *scratch, void *stream);
/* scipy.ndimage.binary_dilation(in, iterations, mask) on 0/1 byte rasters, default structure
 * (4-connected), border 0.  scratch: rows * cols bytes.  out may not alias in. */
int  pb200_masked_dilation(pb200_ctx *ctx, const uint8_t *in, const uint8_t *mask, int rows,
                           int cols, int iterations, uint8_t *out, uint8_t *scratch,
                           void *stream);
/* Point-wise tail of the 'cover' flow in one pass: WTR = _apply_cloud_masking(WTR-2, CLOUD) (D:2089), BWTR (D:1710),
 * CONF (D:1733) and, when collapse != 0, _collapse_wtr_classes (D:2578) of WTR, WTR-2 and the optional WTR-1 planes
 * IN PLACE.  wtr / bwtr / conf / wtr1 / wtr1_remapped may be NULL. */
int  pb200_cover_tail(pb200_ctx *ctx, uint8_t *wtr2, const uint8_t *cloud, int64_t n, uint8_t *wtr,
                      uint8_t *bwtr, uint8_t *conf, uint8_t *wtr1, uint8_t *wtr1_remapped, int collapse,
                      void *stream);
/* D:2089-2133 _apply_cloud_masking */
int  pb200_cloud_masking(pb200_ctx *ctx, const uint8_t *wtr2,
                         const uint8_t *cloud, int64_t n, uint8_t *wtr,
                         void *stream);
/* D:1710-1730 _get_binary_water_layer */
int  pb200_binary_water(pb200_ctx *ctx, const uint8_t *wtr, int64_t n,
                        uint8_t *bwtr, void *stream);
/* D:1733-1837 _get_confidence_layer */
int  pb200_confidence(pb200_ctx *ctx, const uint8_t *wtr2,
                      const uint8_t *cloud, int64_t n, uint8_t *conf,
                      void *stream);
/* D:2578-2598 _collapse_wtr_classes */
int  pb200_collapse(pb200_ctx *ctx, const uint8_t *layer, int64_t n,
                    uint8_t *collapsed, void *stream);
/* D:4215-4283 _compute_opera_shadow_layer over a whole float32 DEM
 * (rows x cols, pitch = cols), one-sided differences on the array border
 * exactly like np.gradient; out[rows*cols] uint8 1 = not shadow. */
int  pb200_shadow(pb200_ctx *ctx, const float *dem, int rows, int cols,
                  double sun_azimuth, double sun_elevation,
                  const double *sun_terms /* 5 doubles as in pb200_tile, or NULL */,
                  const pb200_params *params, uint8_t *out, void *stream);

/* The same for a float64 DEM; an integer-typed DEM is converted to float64 by the caller, which is what np.gradient
 * does with it (D:4255): every operation of D:4255-4281 in float64. */
int  pb200_shadow_f64(pb200_ctx *ctx, const double *dem, int rows, int cols,
                      double sun_azimuth, double sun_elevation, const double *sun_terms,
                      const pb200_params *params, uint8_t *out, void *stream);

/* SURVEY 8f next #1 - the numpy tail of create_landcover_mask (D:1003-1115): 3x3 block counts of the
 * 10 m ESA WorldCover raster [3*rows, 3*cols] (water {80,90,95}, urban 50, tree 10), tree count kept
 * on CGLS forest classes only (forest_class_table[v] != 0), threshold hierarchy thresholds[4] =
 * {evergreen, low-intensity, high-intensity, water} (D:270-271) -> LAND [rows, cols] */
int  pb200_landcover_aggregate(pb200_ctx *ctx, const uint8_t *worldcover_up3,
                               const uint8_t *copernicus, int rows, int cols,
                               const uint8_t forest_class_table[256], int year_offset,
                               const int32_t thresholds[4], uint8_t *land, void *stream);

/* SURVEY 8f next #4 - the point-wise tails that let the whole product leave the GPU.
 *
 * pb200_browse_table (host only): the 256-entry relabel table of _compute_browse_array (D:3057-3129):
 * PSW-aggressive -> not water (exclude_psw_aggressive), collapse (D:2578), then not-water / cloud / snow /
 * ocean-masked -> 255 as the four flags say.  pb200_byte_table applies any such table to a uint8 raster
 * (table in HOST memory, passed by value to the kernel). */
int  pb200_browse_table(int collapse_wtr_classes, int exclude_psw_aggressive,
                        int set_not_water_to_nodata, int set_cloud_to_nodata,
                        int set_snow_to_nodata, int set_ocean_masked_to_nodata,
                        uint8_t table[256]);
int  pb200_byte_table(pb200_ctx *ctx, const uint8_t *in, int64_t n, const uint8_t table[256],
                      uint8_t *out, void *stream);
/* D:2301-2302 and D:3024-3036: out = float32(scale) * (float32(band) - float32(offset)), float32
 * arithmetic as numpy evaluates it with Python-float scalars; invalid (bool raster or NULL): NaN where
 * set (D:3033-3036). */
int  pb200_scale_offset(pb200_ctx *ctx, const int16_t *band, int64_t n, double scale, double offset,
                        const uint8_t *invalid, float *out, void *stream);

/* SURVEY 8f next #3 - the numpy half of the 'otsu' shadow algorithm, _compute_otsu_threshold (D:1638-1684), for a
 * uint8 raster (the hillshade GDAL's DEMProcessing writes, D:4206; the hillshade itself is a GDAL file operation and
 * stays on the host).  pb200_histogram_u8: exact count of every byte value (counts[256], uint64, DEVICE memory,
 * accumulated into - zero it first; ranks of a mosaic all-reduce these).  pb200_otsu_threshold (host only): numpy's
 * np.histogram(image, bins=256) binning of those counts and the float64 Otsu arithmetic in numpy's operation order ->
 * the threshold the reference compares with.  pb200_greater_than_u8: out = image > threshold (bool as uint8). */
/* The hillshade of the 'otsu' algorithm (D:4177-4212: gdal.DEMProcessing(..., "hillshade", azimuth, altitude)) over a
 * float32 DEM (rows x cols, pitch = cols; ewres / nsres = pixel spacing of the geotransform, nsres < 0 for a north-up
 * raster), Byte output with 0 on the first / last row and column (gdaldem without -compute_edges), fused with the
 * 256-bin count of its own output (counts may be NULL; accumulated into, DEVICE memory).  PARITY UNPINNED: GDAL's
 * arithmetic is not part of the reference tree; this is the published gdaldem Horn formula (z = 1, scale = 1), pinned
 * to oracle/dswx_oracle.py:compute_hillshade_gdal. */
int  pb200_hillshade(pb200_ctx *ctx, const float *dem, int rows, int cols, double sun_azimuth, double sun_elevation,
                     double ewres, double nsres, uint8_t *out, unsigned long long *counts, void *stream);
int  pb200_histogram_u8(pb200_ctx *ctx, const uint8_t *image, int64_t n, unsigned long long *counts, void *stream);
int  pb200_otsu_threshold(const unsigned long long counts[256], int is_normalized, double *threshold);
int  pb200_greater_than_u8(pb200_ctx *ctx, const uint8_t *image, int64_t n, double threshold, uint8_t *out,
                           void *stream);

/* ---- multi-GPU: one oversized raster in row strips (BASELINE configs[4]) -------------------------
 * Every function of the path is point-wise except the one-row stencil of np.gradient inside
 * _compute_opera_shadow_layer (D:4255): the rank that owns pixel rows [r0, r1) needs DEM rows r0 - 1 and r1.  The
 * reference never splits a raster; parity is defined against the reference run on the whole raster.
 *
 * One process per GPU.  Rank 0 calls pb200_comm_unique_id and hands the 128 bytes to the other ranks over any host
 * channel (MPI, torch.distributed, a file); every rank then calls pb200_comm_init on its context (collective).  NCCL
 * is bound at run time (dlopen of libnccl.so.2, preferring the copy already in the process; PB200_NCCL_PATH
 * overrides): PB200_E_NCCL if it cannot be found.
 *
 * pb200_halo_exchange_dem: dem_ext is the strip's DEM as (n_rows + 2) x pitch float32 on the device, rows 1..n_rows
 * = the strip's own DEM rows (column margins included), row 0 / row n_rows + 1 = the halo rows.  Sends row 1 to rank
 * - 1 and row n_rows to rank + 1, receives row 0 from rank - 1 and row n_rows + 1 from rank + 1 (ncclSend / ncclRecv
 * in one group, asynchronous on `stream`; NVLink on one box).  Rank 0 keeps its own row 0 and the last rank its own
 * last row (both come from the DEM margin the host supplies).  Then classify with a pb200_tile whose dem = dem_ext,
 * dem_rows = n_rows + 2, dem_off_y = 1.
 * pb200_comm_allreduce_u64: in-place sum over the ranks (the three coverage counters, D:5104-5111; the 256 Otsu
 * histogram counts). */
#define PB200_COMM_ID_BYTES 128
int  pb200_comm_unique_id(uint8_t id[PB200_COMM_ID_BYTES]);
int  pb200_comm_init(pb200_ctx *ctx, const uint8_t id[PB200_COMM_ID_BYTES], int rank, int nranks);
int  pb200_halo_exchange_dem(pb200_ctx *ctx, float *dem_ext, int n_rows, int pitch, void *stream);
int  pb200_comm_allreduce_u64(pb200_ctx *ctx, uint64_t *values, int n, void *stream);
int  pb200_comm_destroy(pb200_ctx *ctx);

/* ---- helpers exported for tests ---------------------------------------- */
/* The exact integer form of "float64(n)/float64(d) > t" (is_less = 0) or
 * "< t" (is_less = 1) for int16 n, d:  with p/q = n/d, q > 0,
 *   >  t  <=>  p * b >= a * q          <  t  <=>  p * b <= a * q
 * b == 0 encodes always (a = -1 / +1) or never (a = +1 / -1). */
int  pb200_ratio_bound(double t, int is_less, int32_t *a, int32_t *b);
/* Exhaustive check of that form against IEEE float64 division over all
 * 2^32 (n, d) int16 pairs on the GPU; *mismatches = number of disagreeing
 * pairs (d == 0 included: inf / nan semantics). Synchronous. */
int  pb200_ratio_sweep(pb200_ctx *ctx, double t, int is_less,
                       uint64_t *mismatches);
/* The FAST8 kernel variant's integer forms of the rational tests (one IDP.2A each on per-pixel packs, sign bookkeeping for
 * int16 sums that wrapped) and of 4*awesh against numpy's arithmetic (D:1872-1914), over EVERY clipped (green, swir1)
 * and (nir, red) pair in [1, 32767]^2.  counts[0..4] = mismatches of mndwi > wigt, > pswt_1_mndwi, > pswt_2_mndwi,
 * ndvi < pswt_1_ndvi, awesh > awgt (must all be 0); counts[5] = visited pairs whose int16 sum wrapped.
 * PB200_E_UNSUPPORTED when the parameters do not select FAST8.  Synchronous. */
int  pb200_fast8_sweep(pb200_ctx *ctx, const pb200_params *params, uint64_t counts[6]);
/* The float32 shortcuts of the terrain-shadow test (D:4264-4281) against the exact float64 sequence on n_samples DEM
 * neighbourhoods generated on the GPU: mode 0 random gradients, 1 on the back-slope boundary (+- a relative 2^-25 ..
 * 2^-12), 2 on the incidence boundary, 3 special values (zeros, denormals, 1e30, +-inf, NaN).  counts[8] = samples,
 * exact-shadow samples, then (decided, decided-but-wrong) for the compare shortcut, the sign-bit shortcut (FAST8) and
 * its packed flavour.  "wrong" must be 0: a decided pixel never differs from the reference sequence.  Synchronous. */
int  pb200_shadow_sweep(pb200_ctx *ctx, const pb200_params *params, double sun_azimuth, double sun_elevation,
                        const double *sun_terms /* 5 doubles or NULL */, int mode, uint64_t seed, uint64_t n_samples,
                        uint64_t counts[8]);
/* Derived angle thresholds the library would use for these params (libm). */
int  pb200_angle_thresholds(const pb200_params *params, double *cos_inc,
                            double *tan_slope);

#ifdef __cplusplus
}
#endif
#endif /* PROTEUS_B200_H */
