"""Drop-in replacements for the per-pixel functions of ``proteus.dswx_hls``.

Same names, argument meaning, dtypes, in-place / return conventions and error
behaviour as the reference functions they replace (nasa/PROTEUS v1.0.2,
src/proteus/dswx_hls.py, cited per function as ``D:``); the work is done by
the function-granular CUDA entry points of libproteus_b200.so.  numpy arrays
in, numpy arrays out.  ``proteus_b200.install()`` rebinds these onto the
reference module so that ``generate_dswx_layers`` (D:4610) runs on them
unchanged; the one-pass fused path is ``proteus_b200.classify_tile``.

Scope notes (explicit errors, no silent fallback):
  * reflectance bands are int16 (the default, D:4640) or, for ``_compute_diagnostic_tests`` only, all
    float32 (``--offset-and-scale-inputs``); the fused path is int16 only;
  * the DEM of the FUSED path must be float32 (what the cubic warp of D:5145 produces); the function-level
    ``_compute_opera_shadow_layer`` also takes float64 and integer DEMs (np.gradient's float64 promotion).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import get_context
from .params import (HlsThresholds, aerosol_class_bits, check_adjacent_mode,
                     make_params, sun_terms)

# constants with the reference's names and values (D:26-58, D:94-185)
FLAG_COLLAPSE_WTR_CLASSES = True
FLAG_CLIP_NEGATIVE_REFLECTANCE = True
SCALE_FACTOR = 0.0001
AEROSOL_REMAPPING_MAX_NIR = 0.1 / SCALE_FACTOR
UINT8_FILL_VALUE = 255
DEM_MARGIN_IN_PIXELS = 50
DIAGNOSTIC_LAYER_NO_DATA_DECIMAL = 0b100000
DIAGNOSTIC_LAYER_NO_DATA_BINARY_REPR = 65535
WATER_NOT_WATER_CLEAR = 0
FIRST_UNCOLLAPSED_WATER_CLASS = 1
LAST_UNCOLLAPSED_WATER_CLASS = 4
WTR_SNOW_MASKED = 252
WTR_CLOUD_MASKED = 253
WTR_OCEAN_MASKED = 254
SHAD_NOT_MASKED = 1
SHAD_MASKED = 0
BWTR_WATER = 1

REPLACED_FUNCTIONS = (
    '_compute_diagnostic_tests', 'generate_interpreted_layer',
    '_get_binary_representation', '_compute_preliminary_cloud_layer',
    '_apply_aerosol_class_remapping', '_apply_landcover_and_shadow_masks',
    '_add_snow_to_cloud_layer', '_apply_cloud_masking',
    '_get_binary_water_layer', '_get_confidence_layer',
    '_collapse_wtr_classes', '_compute_opera_shadow_layer',
    '_compute_browse_array', '_compute_otsu_threshold',
)


# ---------------------------------------------------------------------------
# plumbing: numpy <-> device
# ---------------------------------------------------------------------------
def _torch():
    import torch
    return torch


def _to_device(a, dtype, name):
    torch = _torch()
    a = np.asarray(a)
    if a.dtype != np.dtype(dtype):
        raise TypeError(f'{name}: expected {np.dtype(dtype).name}, got {a.dtype}')
    a = np.ascontiguousarray(a)
    if dtype == np.uint16:                       # torch: move the bits as int16
        return torch.from_numpy(a.view(np.int16)).cuda()
    if dtype == np.bool_:
        return torch.from_numpy(a.view(np.uint8)).cuda()
    return torch.from_numpy(a).cuda()


def _empty_like_device(shape, dtype):
    torch = _torch()
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.int16,
           np.dtype(np.int16): torch.int16, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    return torch.empty(tuple(shape), dtype=tdt, device='cuda')


def _to_host(t, dtype):
    a = t.cpu().numpy()
    return a.view(dtype) if a.dtype != np.dtype(dtype) else a


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _band(a, name):
    a = np.asarray(a)
    if np.issubdtype(a.dtype, np.floating):
        raise NotImplementedError(
            f'{name}: float reflectances (--offset-and-scale-inputs, D:2300-2302) '
            'are not supported by the int16 GPU path yet')
    return _to_device(a, np.int16, name)


# ---------------------------------------------------------------------------
# D:1840-1916
# ---------------------------------------------------------------------------
def _compute_diagnostic_tests(blue, green, red, nir, swir1, swir2,
                              hls_thresholds):
    """Diagnostic tests over the six bands -> uint16 code 0..31 (D:1840)."""
    ctx = get_context()
    shape = np.shape(blue)
    params = make_params(hls_thresholds)
    all_bands = (blue, green, red, nir, swir1, swir2)
    if all(np.asarray(b).dtype == np.float32 for b in all_bands):
        # scaled reflectances (--offset-and-scale-inputs, D:2300-2302): float32 arithmetic in numpy's order
        dev = [_to_device(b, np.float32, 'band') for b in all_bands]
        for d in dev:
            if tuple(d.shape) != tuple(shape):
                raise ValueError('operands could not be broadcast together: all bands must have the same shape')
        out = _empty_like_device(shape, np.uint16)
        ptrs = (C.c_void_p * 6)(*[d.data_ptr() for d in dev])
        _lib.check(ctx._lib.pb200_diagnostic_tests_f32(
            ctx.handle, ptrs, C.byref(params.th), int(np.prod(shape)), out.data_ptr(), _stream()))
        return _to_host(out, np.uint16)
    dev = [_band(b, n) for b, n in zip(all_bands,
                                        ('blue', 'green', 'red', 'nir', 'swir1', 'swir2'))]
    for d in dev:
        if tuple(d.shape) != tuple(shape):
            raise ValueError('operands could not be broadcast together: all bands '
                             'must have the same shape')
    out = _empty_like_device(shape, np.uint16)
    ptrs = (C.c_void_p * 6)(*[d.data_ptr() for d in dev])
    _lib.check(ctx._lib.pb200_diagnostic_tests(
        ctx.handle, ptrs, C.byref(params.th), int(np.prod(shape)), out.data_ptr(), _stream()))
    return _to_host(out, np.uint16)


# ---------------------------------------------------------------------------
# D:1687-1707
# ---------------------------------------------------------------------------
def generate_interpreted_layer(diagnostic_layer):
    """DIAG code -> WTR-1 class through the 33-entry table (D:97-143); every
    other value -> 255.  Accepts any integer dtype like the reference."""
    ctx = get_context()
    d = np.asarray(diagnostic_layer)
    if not np.issubdtype(d.dtype, np.integer):
        raise TypeError('diagnostic_layer must be an integer array')
    if d.dtype != np.uint16:
        outside = (d < 0) | (d > 65535)
        d = np.where(outside, 65535, d).astype(np.uint16)     # 65535 is not in the table
    dev = _to_device(d, np.uint16, 'diagnostic_layer')
    out = _empty_like_device(d.shape, np.uint8)
    _lib.check(ctx._lib.pb200_interpreted_layer(
        ctx.handle, dev.data_ptr(), d.size, out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


# ---------------------------------------------------------------------------
# D:4286-4317
# ---------------------------------------------------------------------------
def _get_binary_representation(diagnostic_layer_decimal, nbits=6):
    if nbits != 6:
        raise NotImplementedError('only nbits=6 (the value the reference uses, D:5231)')
    ctx = get_context()
    d = np.asarray(diagnostic_layer_decimal)
    dev = _to_device(d, np.uint16, 'diagnostic_layer_decimal')
    out = _empty_like_device(d.shape, np.uint16)
    _lib.check(ctx._lib.pb200_binary_representation(
        ctx.handle, dev.data_ptr(), d.size, out.data_ptr(), _stream()))
    return _to_host(out, np.uint16)


# ---------------------------------------------------------------------------
# D:1919-1993
# ---------------------------------------------------------------------------
def _compute_preliminary_cloud_layer(fmask, mask_adjacent_to_cloud_mode):
    mode = check_adjacent_mode(mask_adjacent_to_cloud_mode)      # raises like D:1977-1981
    ctx = get_context()
    f = np.asarray(fmask)
    dev = _to_device(f, np.uint8, 'fmask')
    out = _empty_like_device(f.shape, np.uint8)
    _lib.check(ctx._lib.pb200_preliminary_cloud(
        ctx.handle, dev.data_ptr(), mode, f.size, out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


# ---------------------------------------------------------------------------
# D:1249-1302  (mutates wtr_1_layer and preliminary_cloud_layer, returns None)
# ---------------------------------------------------------------------------
def _apply_aerosol_class_remapping(
        wtr_1_layer, nir, preliminary_cloud_layer, fmask,
        aerosol_not_water_to_high_conf_water_fmask_values,
        aerosol_water_moderate_conf_to_high_conf_water_fmask_values,
        aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values,
        aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values):
    ctx = get_context()
    bits = aerosol_class_bits(
        aerosol_not_water_to_high_conf_water_fmask_values,
        aerosol_water_moderate_conf_to_high_conf_water_fmask_values,
        aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values,
        aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values)
    w = _to_device(wtr_1_layer, np.uint8, 'wtr_1_layer')
    n = _band(nir, 'nir')
    c = _to_device(preliminary_cloud_layer, np.uint8, 'preliminary_cloud_layer')
    f = _to_device(fmask, np.uint8, 'fmask')
    cbits = (C.c_uint8 * 256)(*[int(v) for v in bits])
    _lib.check(ctx._lib.pb200_aerosol_remap(
        ctx.handle, w.data_ptr(), n.data_ptr(), c.data_ptr(), f.data_ptr(), cbits,
        int(w.numel()), _stream()))
    wtr_1_layer[...] = _to_host(w, np.uint8)
    preliminary_cloud_layer[...] = _to_host(c, np.uint8)
    return None


# ---------------------------------------------------------------------------
# D:1305-1378
# ---------------------------------------------------------------------------
def _apply_landcover_and_shadow_masks(interpreted_layer, nir, landcover_mask,
                                      shadow_layer, hls_thresholds):
    ctx = get_context()
    w = _to_device(interpreted_layer, np.uint8, 'interpreted_layer')
    n = _band(nir, 'nir') if nir is not None else None
    land = (_to_device(landcover_mask, np.uint8, 'landcover_mask')
            if landcover_mask is not None else None)
    shad = None
    if shadow_layer is not None:
        s = np.asarray(shadow_layer)
        shad = _to_device(s, s.dtype if s.dtype == np.bool_ else np.uint8, 'shadow_layer')
    out = _empty_like_device(w.shape, np.uint8)
    lc = hls_thresholds.lcmask_nir if land is not None else 0.0
    _lib.check(ctx._lib.pb200_landcover_shadow_masks(
        ctx.handle, w.data_ptr(), n.data_ptr() if n is not None else None,
        land.data_ptr() if land is not None else None,
        shad.data_ptr() if shad is not None else None,
        float(lc), int(w.numel()), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


# ---------------------------------------------------------------------------
# D:1996-2086  (mutates and returns cloud_layer)
# ---------------------------------------------------------------------------
def _add_snow_to_cloud_layer(wtr_2_layer, cloud_layer, fmask,
                             mask_adjacent_to_cloud_mode):
    # the reference does not validate the mode here: anything but 'cover' behaves like 'mask'/'ignore'
    ctx = get_context()
    w = _to_device(wtr_2_layer, np.uint8, 'wtr_2_layer')
    c = _to_device(cloud_layer, np.uint8, 'cloud_layer')
    f = _to_device(fmask, np.uint8, 'fmask')
    if mask_adjacent_to_cloud_mode == 'cover':
        if w.dim() != 2:
            raise ValueError("mode 'cover' dilates a 2-D raster")
        scratch = _torch().empty(4 * max(int(w.numel()), 1), dtype=_torch().uint8, device='cuda')
        _lib.check(ctx._lib.pb200_snow_to_cloud_cover(
            ctx.handle, w.data_ptr(), c.data_ptr(), f.data_ptr(), int(w.shape[0]), int(w.shape[1]),
            scratch.data_ptr(), _stream()))
    else:
        _lib.check(ctx._lib.pb200_snow_to_cloud(
            ctx.handle, w.data_ptr(), c.data_ptr(), f.data_ptr(), 0, int(w.numel()), _stream()))
    cloud_layer[...] = _to_host(c, np.uint8)
    return cloud_layer


# ---------------------------------------------------------------------------
# D:2089-2133, D:1710-1730, D:1733-1837, D:2578-2598
# ---------------------------------------------------------------------------
def _apply_cloud_masking(wtr_2_layer, cloud_layer):
    ctx = get_context()
    w = _to_device(wtr_2_layer, np.uint8, 'wtr_2_layer')
    c = _to_device(cloud_layer, np.uint8, 'cloud_layer')
    out = _empty_like_device(w.shape, np.uint8)
    _lib.check(ctx._lib.pb200_cloud_masking(
        ctx.handle, w.data_ptr(), c.data_ptr(), int(w.numel()), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


def _get_binary_water_layer(wtr_layer):
    ctx = get_context()
    w = _to_device(wtr_layer, np.uint8, 'wtr_layer')
    out = _empty_like_device(w.shape, np.uint8)
    _lib.check(ctx._lib.pb200_binary_water(
        ctx.handle, w.data_ptr(), int(w.numel()), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


def _get_confidence_layer(wtr_2_layer, cloud_layer):
    ctx = get_context()
    w = _to_device(wtr_2_layer, np.uint8, 'wtr_2_layer')
    c = _to_device(cloud_layer, np.uint8, 'cloud_layer')
    out = _empty_like_device(w.shape, np.uint8)
    _lib.check(ctx._lib.pb200_confidence(
        ctx.handle, w.data_ptr(), c.data_ptr(), int(w.numel()), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


def _collapse_wtr_classes(interpreted_layer):
    ctx = get_context()
    w = _to_device(interpreted_layer, np.uint8, 'interpreted_layer')
    out = _empty_like_device(w.shape, np.uint8)
    _lib.check(ctx._lib.pb200_collapse(
        ctx.handle, w.data_ptr(), int(w.numel()), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


# ---------------------------------------------------------------------------
# D:4215-4283, D:4320-4337
# ---------------------------------------------------------------------------
def _compute_opera_shadow_layer(dem, sun_azimuth_angle, sun_elevation_angle,
                                min_slope_angle, max_sun_local_inc_angle,
                                pixel_spacing_x=30, pixel_spacing_y=30):
    """Bool mask (True = not shadow) over the whole DEM incl. its margin."""
    ctx = get_context()
    d = np.asarray(dem)
    if d.ndim != 2:
        raise ValueError('dem must be 2-D')
    params = make_params(min_slope_angle=min_slope_angle,
                         max_sun_local_inc_angle=max_sun_local_inc_angle,
                         pixel_spacing=(pixel_spacing_x, pixel_spacing_y))
    # the five float64 sun scalars come from numpy, like the reference's (D:4245-4252)
    terms = (C.c_double * 5)(*sun_terms(sun_azimuth_angle, sun_elevation_angle))
    if d.dtype == np.float32:                     # the cubic-warped DEM (D:5145-5150): float32 up to the normalisation
        entry, dev = ctx._lib.pb200_shadow, _to_device(d, np.float32, 'dem')
    elif d.dtype == np.float64 or np.issubdtype(d.dtype, np.integer):
        # np.gradient converts an integer DEM to float64 first; from there on everything is float64
        entry = ctx._lib.pb200_shadow_f64
        dev = _torch().from_numpy(np.ascontiguousarray(d, dtype=np.float64)).cuda()
    else:
        raise NotImplementedError(f'DEM dtype {d.dtype}: float32, float64 or an integer type')
    out = _empty_like_device(d.shape, np.uint8)
    _lib.check(entry(
        ctx.handle, dev.data_ptr(), d.shape[0], d.shape[1], float(sun_azimuth_angle),
        float(sun_elevation_angle), terms, C.byref(params), out.data_ptr(), _stream()))
    return _to_host(out, np.uint8).astype(bool)


# ---------------------------------------------------------------------------
# D:874-904 + D:1003-1115: the numpy tail of create_landcover_mask
# ---------------------------------------------------------------------------
landcover_threshold_dict = {"standard": [6, 3, 7, 3], "water heavy": [6, 3, 7, 1]}     # D:270-271


def landcover_aggregate(worldcover_array_up_3, copernicus_landcover_array,
                        forest_mask_landcover_classes, year, mask_type='standard'):
    """LAND layer from the two warped rasters ``create_landcover_mask`` holds at D:1003: the ESA
    WorldCover raster at 10 m (3x the product grid, uint8) and the CGLS land cover on the product
    grid (uint8).  ``year`` is the WorldCover map year the reference reads from the file metadata
    (D:1064-1094).  The reference has no separate function for these statements; this one replaces
    D:1003-1115."""
    ctx = get_context()
    wc = np.asarray(worldcover_array_up_3)
    cp = np.asarray(copernicus_landcover_array)
    if wc.ndim != 2 or cp.ndim != 2 or wc.shape != (3 * cp.shape[0], 3 * cp.shape[1]):
        raise ValueError('worldcover_array_up_3 must be exactly 3x the CGLS / product grid')
    th = landcover_threshold_dict[mask_type.lower()]
    table = np.zeros(256, np.uint8)
    for c in (forest_mask_landcover_classes or ()):
        if 0 <= int(c) <= 255:
            table[int(c)] = 1
    dwc = _to_device(wc, np.uint8, 'worldcover_array_up_3')
    dcp = _to_device(cp, np.uint8, 'copernicus_landcover_array')
    out = _empty_like_device(cp.shape, np.uint8)
    _lib.check(ctx._lib.pb200_landcover_aggregate(
        ctx.handle, dwc.data_ptr(), dcp.data_ptr(), int(cp.shape[0]), int(cp.shape[1]),
        (C.c_uint8 * 256)(*[int(v) for v in table]), int(year) - 2000, (C.c_int32 * 4)(*th),
        out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


# ---------------------------------------------------------------------------
# D:3057-3129 browse relabel; D:2301-2302 / D:3024-3036 offset and scale
# ---------------------------------------------------------------------------
def _compute_browse_array(masked_interpreted_water_layer, flag_collapse_wtr_classes=True,
                          exclude_psw_aggressive=False, set_not_water_to_nodata=False,
                          set_cloud_to_nodata=False, set_snow_to_nodata=False,
                          set_ocean_masked_to_nodata=True):
    """Same signature and result as the reference function: a fresh uint8 array."""
    ctx = get_context()
    table = (C.c_uint8 * 256)()
    _lib.check(ctx._lib.pb200_browse_table(
        int(bool(flag_collapse_wtr_classes)), int(bool(exclude_psw_aggressive)),
        int(bool(set_not_water_to_nodata)), int(bool(set_cloud_to_nodata)),
        int(bool(set_snow_to_nodata)), int(bool(set_ocean_masked_to_nodata)), table))
    w = _to_device(masked_interpreted_water_layer, np.uint8, 'masked_interpreted_water_layer')
    out = _empty_like_device(w.shape, np.uint8)
    _lib.check(ctx._lib.pb200_byte_table(
        ctx.handle, w.data_ptr(), int(w.numel()), table, out.data_ptr(), _stream()))
    return _to_host(out, np.uint8)


def scale_and_offset_band(image, scale_factor, offset, invalid_ind=None):
    """``scale_factor * (np.asarray(image, dtype=np.float32) - offset)`` (D:2301-2302, D:3024-3031) for an
    int16 band, with ``invalid_ind`` (the ``np.where`` tuple of D:5040, or the bool raster it came from) set to NaN
    (D:3033-3036).  The
    reference has these statements inline in ``_load_hls_band_from_file`` and ``_save_output_rgb_file``."""
    ctx = get_context()
    x = _to_device(image, np.int16, 'image')
    inv = None
    if invalid_ind is not None:
        if isinstance(invalid_ind, tuple):            # the np.where(...) tuple of D:5040
            inv_np = np.zeros(tuple(x.shape), dtype=np.bool_)
            inv_np[invalid_ind] = True
        else:
            inv_np = np.asarray(invalid_ind)
        if inv_np.dtype != np.bool_ or inv_np.shape != tuple(x.shape):
            raise NotImplementedError('invalid_ind: a bool raster of the band\'s shape or an np.where tuple')
        inv = _to_device(inv_np.view(np.uint8), np.uint8, 'invalid_ind')
    out = _empty_like_device(x.shape, np.float32)
    _lib.check(ctx._lib.pb200_scale_offset(
        ctx.handle, x.data_ptr(), int(x.numel()), float(scale_factor), float(offset),
        inv.data_ptr() if inv is not None else None, out.data_ptr(), _stream()))
    return _to_host(out, np.float32)


# ---------------------------------------------------------------------------
# D:1638-1684: the numpy half of the 'otsu' shadow algorithm
# ---------------------------------------------------------------------------
def _otsu_counts(image):
    """Exact per-value counts (numpy uint64[256]) of a uint8 raster, counted on the GPU."""
    torch = _torch()
    ctx = get_context()
    x = _to_device(image, np.uint8, 'image')
    counts = torch.zeros(256, dtype=torch.int64, device='cuda')
    _lib.check(ctx._lib.pb200_histogram_u8(ctx.handle, x.data_ptr(), int(x.numel()), counts.data_ptr(), _stream()))
    return x, counts


def otsu_threshold_from_counts(counts, is_normalized=True):
    """np.histogram(image, bins=256) binning + the float64 Otsu arithmetic of D:1663-1686 on 256 exact per-value
    counts (host; no GPU needed).  Ranks holding strips of one raster sum their counts first."""
    arr = (C.c_uint64 * 256)(*[int(v) for v in np.asarray(counts).ravel()])
    thr = C.c_double()
    _lib.check(_lib.load().pb200_otsu_threshold(arr, int(bool(is_normalized)), C.byref(thr)))
    return thr.value


def _compute_otsu_threshold(image, is_normalized=True):
    """Same result as the reference function for the uint8 hillshade GDAL writes (D:4206-4209): bool array
    ``image > threshold``.  The hillshade itself is a GDAL file operation and stays on the host."""
    img = np.asarray(image)
    if img.dtype != np.uint8:
        raise NotImplementedError(f'_compute_otsu_threshold: image dtype {img.dtype}; only the uint8 hillshade')
    if img.size == 0:
        raise ValueError('attempt to get argmax of an empty sequence')        # what numpy raises at D:1684
    ctx = get_context()
    x, counts = _otsu_counts(img)
    thr = otsu_threshold_from_counts(counts.cpu().numpy(), is_normalized)
    out = _empty_like_device(x.shape, np.uint8)
    _lib.check(ctx._lib.pb200_greater_than_u8(
        ctx.handle, x.data_ptr(), int(x.numel()), thr, out.data_ptr(), _stream()))
    return _to_host(out, np.uint8).astype(bool)


def compute_hillshade(dem, sun_azimuth_angle, sun_elevation_angle, pixel_spacing_x=30.0, pixel_spacing_y=-30.0,
                      return_counts=False):
    """What ``_compute_hillshade`` (D:4177-4212) reads back from ``gdal.DEMProcessing(..., "hillshade", azimuth,
    altitude)``, computed on the GPU from the DEM ARRAY (the reference's function takes the DEM *file*; see
    INTEGRATION.md): uint8, 0 on the border.  PARITY UNPINNED - GDAL's arithmetic is not part of the reference tree; this
    is the published gdaldem Horn formula (the test oracle's ``compute_hillshade_gdal``).  With ``return_counts``
    also the exact 256-bin histogram of the result (device tensor), counted in the same pass."""
    torch = _torch()
    ctx = get_context()
    d = np.asarray(dem)
    if d.dtype != np.float32 or d.ndim != 2:
        raise NotImplementedError(f'compute_hillshade: DEM dtype {d.dtype}; a 2-D float32 array')
    x = _to_device(d, np.float32, 'dem')
    out = _empty_like_device(d.shape, np.uint8)
    counts = torch.zeros(256, dtype=torch.int64, device='cuda') if return_counts else None
    _lib.check(ctx._lib.pb200_hillshade(
        ctx.handle, x.data_ptr(), int(d.shape[0]), int(d.shape[1]), float(sun_azimuth_angle), float(sun_elevation_angle),
        float(pixel_spacing_x), float(pixel_spacing_y), out.data_ptr(),
        counts.data_ptr() if counts is not None else None, _stream()))
    if return_counts:
        return out, counts
    return _to_host(out, np.uint8)


def compute_otsu_shadow_layer(dem, sun_azimuth_angle, sun_elevation_angle, pixel_spacing_x=30.0, pixel_spacing_y=-30.0):
    """The 'otsu' branch of D:5152-5157 on the device: hillshade + its histogram in one pass over the DEM, the Otsu
    threshold from the 256 counts on the host, one compare pass.  Bool array, True = not shadow."""
    ctx = get_context()
    hill, counts = compute_hillshade(dem, sun_azimuth_angle, sun_elevation_angle, pixel_spacing_x, pixel_spacing_y,
                                     return_counts=True)
    if hill.numel() == 0:
        raise ValueError('attempt to get argmax of an empty sequence')
    thr = otsu_threshold_from_counts(counts.cpu().numpy(), True)
    out = _empty_like_device(tuple(hill.shape), np.uint8)
    _lib.check(ctx._lib.pb200_greater_than_u8(ctx.handle, hill.data_ptr(), int(hill.numel()), thr, out.data_ptr(), _stream()))
    return _to_host(out, np.uint8).astype(bool)


def _crop_2d_array_all_sides(input_2d_array, margin):
    return input_2d_array[margin:-margin, margin:-margin]


# ---------------------------------------------------------------------------
# rebinding onto the reference module
# ---------------------------------------------------------------------------
_saved = {}


def install(module=None):
    """Rebind the functions of REPLACED_FUNCTIONS on ``proteus.dswx_hls`` (or
    the given module object).  ``generate_dswx_layers`` looks its helpers up
    as module globals at call time, so the unchanged orchestrator then runs on
    the GPU entry points.  Returns the module."""
    if module is None:
        import proteus.dswx_hls as module
    g = globals()
    for name in REPLACED_FUNCTIONS:
        if hasattr(module, name) and (module, name) not in _saved:
            _saved[(module, name)] = getattr(module, name)
        setattr(module, name, g[name])
    return module


def uninstall(module=None):
    if module is None:
        import proteus.dswx_hls as module
    for (mod, name), fn in list(_saved.items()):
        if mod is module:
            setattr(mod, name, fn)
            del _saved[(mod, name)]
    return module
