"""Host-side parameter handling: HlsThresholds mirror, processing options and
their conversion to the C ABI's ``pb200_params`` / sun terms.

Reference: ``HlsThresholds`` src/proteus/dswx_hls.py:274-318, defaults
src/proteus/defaults/dswx_hls.yaml:64-109 and :176-212.
"""
from __future__ import annotations

import functools

import numpy as np

from . import _lib

BAND_NAMES = ('blue', 'green', 'red', 'nir', 'swir1', 'swir2')
DEM_MARGIN_IN_PIXELS = 50            # dswx_hls.py:58

THRESHOLD_FIELDS = ('wigt', 'awgt', 'pswt_1_mndwi', 'pswt_1_nir',
                    'pswt_1_swir1', 'pswt_1_ndvi', 'pswt_2_mndwi',
                    'pswt_2_blue', 'pswt_2_nir', 'pswt_2_swir1',
                    'pswt_2_swir2', 'lcmask_nir')
_DEFAULT_THRESHOLDS = dict(
    wigt=0.124, awgt=0.0, pswt_1_mndwi=-0.44, pswt_1_nir=1500,
    pswt_1_swir1=900, pswt_1_ndvi=0.7, pswt_2_mndwi=-0.5, pswt_2_blue=1000,
    pswt_2_nir=2500, pswt_2_swir1=3000, pswt_2_swir2=1000, lcmask_nir=1200)

# band thresholds that numpy compares as Python ints against int16 arrays
_INT_COMPARED = ('pswt_1_nir', 'pswt_1_swir1', 'pswt_2_blue', 'pswt_2_nir',
                 'pswt_2_swir1', 'pswt_2_swir2', 'lcmask_nir')


class HlsThresholds:
    """Same attribute names as the reference class (dswx_hls.py:274-318);
    unlike the reference's, a fresh instance carries the packaged defaults."""

    def __init__(self, **kw):
        vals = dict(_DEFAULT_THRESHOLDS)
        vals.update(kw)
        for name in THRESHOLD_FIELDS:
            setattr(self, name, vals[name])

    def __repr__(self):
        inner = ', '.join(f'{n}={getattr(self, n)!r}' for n in THRESHOLD_FIELDS)
        return f'HlsThresholds({inner})'


DEFAULT_AEROSOL_FMASK_VALUES = dict(
    aerosol_not_water_to_high_conf_water_fmask_values=[224, 160, 96],
    aerosol_water_moderate_conf_to_high_conf_water_fmask_values=[224, 160, 96],
    aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values=[224, 192, 160, 128, 96],
    aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values=[224, 192, 160, 128, 96])
_AEROSOL_KEYS_TO_CLASS = (
    ('aerosol_not_water_to_high_conf_water_fmask_values', 0),
    ('aerosol_water_moderate_conf_to_high_conf_water_fmask_values', 2),
    ('aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values', 3),
    ('aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values', 4))

DEFAULT_PROCESSING = dict(
    apply_aerosol_class_remapping=True,
    min_slope_angle=-5, max_sun_local_inc_angle=40,
    mask_adjacent_to_cloud_mode='mask',
    **DEFAULT_AEROSOL_FMASK_VALUES)


def check_adjacent_mode(mode):
    """Same failure as dswx_hls.py:1977-1981."""
    if mode not in ('mask', 'ignore', 'cover'):
        raise Exception('ERROR mask adjacent to cloud/cloud-shadow mode:'
                        f' {mode}')
    return _lib.ADJ_MODES[mode]


def aerosol_class_bits(values_not_water, values_moderate,
                       values_psw_conservative, values_psw_aggressive):
    """256-byte table: bit k of entry v set when Fmask value v is in the list
    that remaps WTR-1 class k (dswx_hls.py:1283-1296)."""
    bits = np.zeros(256, dtype=np.uint8)
    for values, cls in ((values_not_water, 0), (values_moderate, 2),
                        (values_psw_conservative, 3),
                        (values_psw_aggressive, 4)):
        for v in (values or ()):
            v = int(v)
            if 0 <= v <= 255:                 # np.isin on a uint8 raster
                bits[v] |= np.uint8(1 << cls)
    return bits


# ---------------------------------------------------------------------------
# angle tests of _compute_opera_shadow_layer in the cosine / tangent domain
# ---------------------------------------------------------------------------
def _f64_key(x):
    k = np.array([x], dtype=np.float64).view(np.int64)[0]
    return int(k) if k >= 0 else -(2 ** 63) - int(k)


def _f64_unkey(k):
    b = k if k >= 0 else -(2 ** 63) - k
    return float(np.array([b], dtype=np.int64).view(np.float64)[0])


def _f32_key(x):
    k = np.array([x], dtype=np.float32).view(np.int32)[0]
    return int(k) if k >= 0 else -(2 ** 31) - int(k)


def _f32_unkey(k):
    b = k if k >= 0 else -(2 ** 31) - k
    return float(np.array([b], dtype=np.int32).view(np.float32)[0])


_NUMPY1_PROMOTION_OVERRIDE = None


def set_numpy1_promotion(value):
    """Force the promotion rules of the terrain-shadow test for every parameter set built from now on: True = numpy 1.x
    (float32), False = numpy >= 2 (float64), None = follow the installed numpy.  Returns the previous setting."""
    global _NUMPY1_PROMOTION_OVERRIDE
    previous, _NUMPY1_PROMOTION_OVERRIDE = _NUMPY1_PROMOTION_OVERRIDE, (None if value is None else bool(value))
    return previous


def numpy1_promotion_default():
    """True when the installed numpy promotes ``float32_array * float64_scalar`` to float32 (numpy 1.x value-based
    casting; the reference pins numpy 1.23.5, setup.py:78), False under NEP 50 (numpy >= 2): the drop-in follows the
    numpy it replaces.  ``set_numpy1_promotion`` overrides."""
    if _NUMPY1_PROMOTION_OVERRIDE is not None:
        return _NUMPY1_PROMOTION_OVERRIDE
    return bool((np.ones(1, np.float32) * np.float64(0.1)).dtype == np.float32)


@functools.lru_cache(maxsize=64)
def angle_thresholds(min_slope_angle, max_sun_local_inc_angle, numpy1_promotion=False):
    """(cos_thr, tan_thr) such that, with numpy's own arccos / arctan /
    degrees (the functions the reference calls at dswx_hls.py:4267-4277),

        degrees(arccos(x)) <= max_inc   <=>  cos_thr <= x <= 1
        degrees(arctan(s)) <= min_slope <=>  s <= tan_thr

    Found by bisection over the ordered float64 values (float32 values with
    ``numpy1_promotion``: under numpy 1.x the reference evaluates both on
    float32 arrays), so the GPU compares against the reference's decision
    boundary to the last bit without evaluating a transcendental per pixel."""
    if numpy1_promotion:
        ftype, key, unkey = np.float32, _f32_key, _f32_unkey
    else:
        ftype, key, unkey = np.float64, _f64_key, _f64_unkey
    with np.errstate(invalid='ignore'):
        # one-element ARRAYS and the Python threshold, exactly the operand kinds of D:4279-4280
        def inc_ok(x):
            return bool((np.degrees(np.arccos(np.array([x], dtype=ftype))) <= max_sun_local_inc_angle)[0])

        def slope_ok(s):
            return bool((np.degrees(np.arctan(np.array([s], dtype=ftype))) <= min_slope_angle)[0])

        if not inc_ok(1.0):
            cos_thr = 2.0
        elif inc_ok(-1.0):
            cos_thr = -1.0
        else:
            lo, hi = key(-1.0), key(1.0)
            while hi - lo > 1:
                mid = (lo + hi) // 2
                if inc_ok(unkey(mid)):
                    hi = mid
                else:
                    lo = mid
            cos_thr = unkey(hi)
        if slope_ok(np.inf):
            tan_thr = float('inf')
        elif not slope_ok(-np.inf):
            tan_thr = float('nan')
        else:
            lo, hi = key(-np.inf), key(np.inf)
            while hi - lo > 1:
                mid = (lo + hi) // 2
                if slope_ok(unkey(mid)):
                    lo = mid
                else:
                    hi = mid
            tan_thr = unkey(lo)
    return cos_thr, tan_thr


def sun_terms(sun_azimuth_angle, sun_elevation_angle):
    """The five float64 scalars of dswx_hls.py:4245-4252 and :4276-4277,
    computed with numpy exactly as the reference does."""
    az = np.radians(sun_azimuth_angle)
    zen = np.radians(90 - sun_elevation_angle)
    return (float(np.sin(az) * np.sin(zen)), float(np.cos(az) * np.sin(zen)),
            float(np.cos(zen)), float(np.sin(az)), float(np.cos(az)))


def make_params(hls_thresholds=None, *, mask_adjacent_to_cloud_mode='mask',
                apply_aerosol_class_remapping=True, aerosol_fmask_values=None,
                min_slope_angle=-5, max_sun_local_inc_angle=40,
                band_fill=-9999, fmask_fill=255, collapse_wtr_classes=True,
                class_histogram=False, pixel_spacing=(30, 30), defer_snow=False,
                numpy1_promotion=None):
    """Build a ``pb200_params`` structure.  ``numpy1_promotion``: evaluate the terrain-shadow test as numpy 1.x does
    (float32 from the dot product on) instead of numpy >= 2 (float64); None = what the installed numpy does."""
    if numpy1_promotion is None:
        numpy1_promotion = numpy1_promotion_default()
    th = hls_thresholds or HlsThresholds()
    p = _lib.Params()
    for name in THRESHOLD_FIELDS:
        v = getattr(th, name)
        if v is None:
            raise ValueError(f'hls_thresholds.{name} is not set')
        if name in _INT_COMPARED and isinstance(v, (int, np.integer)) \
                and not (-32768 <= int(v) <= 32767):
            # numpy >= 2 refuses to compare an int16 array with such an int
            raise OverflowError(
                f'Python integer {v} out of bounds for int16')
        setattr(p.th, name, float(v))
    fills = list(band_fill) if np.ndim(band_fill) else [band_fill] * 6
    for k, f in enumerate(fills):
        p.band_fill[k] = _fill_to_int(f, -32768, 32767)
    p.fmask_fill = _fill_to_int(fmask_fill, 0, 255)
    p.adjacent_mode = check_adjacent_mode(mask_adjacent_to_cloud_mode)
    p.apply_aerosol_class_remapping = int(bool(apply_aerosol_class_remapping))
    lists = dict(DEFAULT_AEROSOL_FMASK_VALUES)
    lists.update(aerosol_fmask_values or {})
    bits = aerosol_class_bits(*(lists[k] for k, _ in _AEROSOL_KEYS_TO_CLASS))
    for v in range(256):
        p.aerosol_class_bits[v] = int(bits[v])
    p.min_slope_angle = float(min_slope_angle)
    p.max_sun_local_inc_angle = float(max_sun_local_inc_angle)
    p.numpy1_promotion = int(bool(numpy1_promotion))
    p.cos_inc_threshold, p.tan_slope_threshold = angle_thresholds(
        min_slope_angle, max_sun_local_inc_angle, bool(numpy1_promotion))
    p.pixel_spacing_x, p.pixel_spacing_y = float(pixel_spacing[0]), float(pixel_spacing[1])
    p.collapse_wtr_classes = int(bool(collapse_wtr_classes))
    p.class_histogram = int(bool(class_histogram))
    p.defer_snow = int(bool(defer_snow))
    return p


def _fill_to_int(fill, lo, hi):
    """``image == fill_value`` (dswx_hls.py:2204) can only be true for an
    integral fill inside the raster dtype's range."""
    if fill is None:
        return _lib.NO_FILL
    f = float(fill)
    if f != f or f != int(f) or not (lo <= int(f) <= hi):
        return _lib.NO_FILL
    return int(f)
