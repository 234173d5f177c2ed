"""ctypes binding of libproteus_b200.so (include/proteus_b200.h).

The library is the product: if it cannot be loaded, or a call fails, this
module raises - there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PB200_LIB_PATH') or os.path.join(HERE, 'libproteus_b200.so')

ABI_VERSION = 1
NO_FILL = -2 ** 31
ADJ_MODES = {'mask': 0, 'ignore': 1, 'cover': 2}
N_COUNTERS = 12
E_INVALID_ARG, E_BAD_MODE, E_UNSUPPORTED, E_NO_DRIVER_API, E_ALIGNMENT, E_NCCL = -1, -2, -3, -4, -5, -6
COMM_ID_BYTES = 128
HOST_REUSE_ANCILLARY, HOST_ASYNC, HOST_SLOT1 = 1, 2, 4
KERNEL_FAST, KERNEL_STREAM, KERNEL_FAST8, KERNEL_GENERIC, KERNEL_STREAM_DYN = 1, 2, 4, 8, 16

EXPORTS = (
    'pb200_version', 'pb200_last_error', 'pb200_ctx_create', 'pb200_ctx_destroy',
    'pb200_params_default', 'pb200_classify', 'pb200_plan_create',
    'pb200_plan_run', 'pb200_plan_kernels', 'pb200_plan_destroy', 'pb200_classify_host', 'pb200_classify_host_ex', 'pb200_host_wait',
    'pb200_host_alloc', 'pb200_host_free', 'pb200_invalid_and_clip',
    'pb200_diagnostic_tests', 'pb200_diagnostic_tests_f32', 'pb200_interpreted_layer',
    'pb200_binary_representation', 'pb200_preliminary_cloud',
    'pb200_aerosol_remap', 'pb200_landcover_shadow_masks',
    'pb200_snow_to_cloud', 'pb200_snow_to_cloud_cover', 'pb200_masked_dilation', 'pb200_cover_tail', 'pb200_cloud_masking', 'pb200_binary_water',
    'pb200_confidence', 'pb200_collapse', 'pb200_shadow', 'pb200_shadow_f64', 'pb200_landcover_aggregate',
    'pb200_browse_table', 'pb200_byte_table', 'pb200_scale_offset',
    'pb200_hillshade', 'pb200_histogram_u8', 'pb200_otsu_threshold', 'pb200_greater_than_u8', 'pb200_ratio_bound',
    'pb200_ratio_sweep', 'pb200_fast8_sweep', 'pb200_shadow_sweep', 'pb200_angle_thresholds',
    'pb200_comm_unique_id', 'pb200_comm_init', 'pb200_halo_exchange_dem', 'pb200_comm_allreduce_u64',
    'pb200_comm_destroy',
)


class Thresholds(C.Structure):
    """pb200_thresholds - mirror of HlsThresholds (reference dswx_hls.py:274-318)."""
    _fields_ = [(n, C.c_double) for n in (
        'wigt', 'awgt', 'pswt_1_mndwi', 'pswt_1_nir', 'pswt_1_swir1',
        'pswt_1_ndvi', 'pswt_2_mndwi', 'pswt_2_blue', 'pswt_2_nir',
        'pswt_2_swir1', 'pswt_2_swir2', 'lcmask_nir')]


class Params(C.Structure):
    """pb200_params"""
    _fields_ = [
        ('th', Thresholds),
        ('band_fill', C.c_int32 * 6),
        ('fmask_fill', C.c_int32),
        ('adjacent_mode', C.c_int32),
        ('apply_aerosol_class_remapping', C.c_int32),
        ('aerosol_class_bits', C.c_uint8 * 256),
        ('min_slope_angle', C.c_double),
        ('max_sun_local_inc_angle', C.c_double),
        ('cos_inc_threshold', C.c_double),
        ('tan_slope_threshold', C.c_double),
        ('pixel_spacing_x', C.c_double),
        ('pixel_spacing_y', C.c_double),
        ('collapse_wtr_classes', C.c_int32),
        ('class_histogram', C.c_int32),
        ('defer_snow', C.c_int32),
        ('numpy1_promotion', C.c_int32),
    ]


class Tile(C.Structure):
    """pb200_tile"""
    _fields_ = [
        ('height', C.c_int32), ('width', C.c_int32),
        ('band', C.c_void_p * 6),
        ('fmask', C.c_void_p),
        ('dem', C.c_void_p),
        ('dem_pitch', C.c_int32), ('dem_rows', C.c_int32),
        ('dem_off_y', C.c_int32), ('dem_off_x', C.c_int32),
        ('land', C.c_void_p), ('ocean', C.c_void_p),
        ('sun_azimuth', C.c_double), ('sun_elevation', C.c_double),
        ('sun_terms', C.c_double * 5),
        ('diag', C.c_void_p), ('wtr1', C.c_void_p),
        ('wtr1_remapped', C.c_void_p), ('wtr2', C.c_void_p),
        ('cloud', C.c_void_p), ('shad', C.c_void_p), ('wtr', C.c_void_p),
        ('bwtr', C.c_void_p), ('conf', C.c_void_p),
        ('counters', C.c_void_p),
    ]


class Pb200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libproteus_b200 error {code}: {message}')
        self.code = code
        self.message = message


_lib = None


def load():
    """Load the shared library (once).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with `python -m proteus_b200.build` '
            '(needs nvcc; the package has no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    lib.pb200_version.restype = C.c_int
    lib.pb200_last_error.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)          # AttributeError = header/library drift
        if name not in ('pb200_last_error',):
            fn.restype = C.c_int
    if lib.pb200_version() != ABI_VERSION:
        raise ImportError('libproteus_b200.so ABI version mismatch')
    lib.pb200_ratio_bound.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.pb200_ratio_sweep.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_uint64)]
    lib.pb200_shadow_sweep.argtypes = [C.c_void_p, C.POINTER(Params), C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int,
                                       C.c_uint64, C.c_uint64, C.c_uint64 * 8]
    lib.pb200_fast8_sweep.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(C.c_uint64)]
    lib.pb200_angle_thresholds.argtypes = [C.POINTER(Params), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pb200_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.pb200_ctx_destroy.argtypes = [C.c_void_p]
    lib.pb200_params_default.argtypes = [C.POINTER(Params)]
    lib.pb200_classify.argtypes = [C.c_void_p, C.POINTER(Tile), C.c_int, C.POINTER(Params), C.c_void_p]
    lib.pb200_plan_create.argtypes = [C.c_void_p, C.POINTER(Tile), C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.pb200_plan_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.pb200_plan_kernels.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.pb200_plan_destroy.argtypes = [C.c_void_p]
    lib.pb200_classify_host.argtypes = [C.c_void_p, C.POINTER(Tile), C.POINTER(Params), C.c_int]
    lib.pb200_classify_host_ex.argtypes = [C.c_void_p, C.POINTER(Tile), C.POINTER(Params), C.c_int, C.c_int]
    lib.pb200_host_wait.argtypes = [C.c_void_p, C.c_int]
    lib.pb200_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.pb200_host_free.argtypes = [C.c_void_p]
    P6 = C.c_void_p * 6
    vp, i64 = C.c_void_p, C.c_int64
    lib.pb200_invalid_and_clip.argtypes = [vp, P6, vp, C.POINTER(Params), i64, P6, vp, vp]
    lib.pb200_diagnostic_tests.argtypes = [vp, P6, C.POINTER(Thresholds), i64, vp, vp]
    lib.pb200_diagnostic_tests_f32.argtypes = [vp, P6, C.POINTER(Thresholds), i64, vp, vp]
    lib.pb200_interpreted_layer.argtypes = [vp, vp, i64, vp, vp]
    lib.pb200_binary_representation.argtypes = [vp, vp, i64, vp, vp]
    lib.pb200_preliminary_cloud.argtypes = [vp, vp, C.c_int, i64, vp, vp]
    lib.pb200_aerosol_remap.argtypes = [vp, vp, vp, vp, vp, C.c_uint8 * 256, i64, vp]
    lib.pb200_landcover_shadow_masks.argtypes = [vp, vp, vp, vp, vp, C.c_double, i64, vp, vp]
    lib.pb200_snow_to_cloud.argtypes = [vp, vp, vp, vp, C.c_int, i64, vp]
    lib.pb200_snow_to_cloud_cover.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    lib.pb200_masked_dilation.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.pb200_cover_tail.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, C.c_int, vp]
    lib.pb200_cloud_masking.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.pb200_binary_water.argtypes = [vp, vp, i64, vp, vp]
    lib.pb200_confidence.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.pb200_collapse.argtypes = [vp, vp, i64, vp, vp]
    lib.pb200_landcover_aggregate.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_uint8 * 256, C.c_int, C.c_int32 * 4, vp, vp]
    lib.pb200_browse_table.argtypes = [C.c_int] * 6 + [C.c_uint8 * 256]
    lib.pb200_byte_table.argtypes = [vp, vp, i64, C.c_uint8 * 256, vp, vp]
    lib.pb200_scale_offset.argtypes = [vp, vp, i64, C.c_double, C.c_double, vp, vp, vp]
    lib.pb200_hillshade.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, vp]
    lib.pb200_histogram_u8.argtypes = [vp, vp, i64, vp, vp]
    lib.pb200_otsu_threshold.argtypes = [C.c_uint64 * 256, C.c_int, C.POINTER(C.c_double)]
    lib.pb200_greater_than_u8.argtypes = [vp, vp, i64, C.c_double, vp, vp]
    lib.pb200_shadow.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(Params), vp, vp]
    lib.pb200_comm_unique_id.argtypes = [C.c_uint8 * COMM_ID_BYTES]
    lib.pb200_comm_init.argtypes = [vp, C.c_uint8 * COMM_ID_BYTES, C.c_int, C.c_int]
    lib.pb200_halo_exchange_dem.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.pb200_comm_allreduce_u64.argtypes = [vp, vp, C.c_int, vp]
    lib.pb200_comm_destroy.argtypes = [vp]
    lib.pb200_shadow_f64.argtypes = lib.pb200_shadow.argtypes
    _lib = lib
    return lib


def check(rc):
    """Raise Pb200Error for a non-zero return code."""
    if rc != 0:
        msg = load().pb200_last_error()
        raise Pb200Error(rc, msg.decode('utf-8', 'replace') if msg else '')


def ratio_bound(t, is_less=False):
    a, b = C.c_int32(), C.c_int32()
    check(load().pb200_ratio_bound(float(t), int(bool(is_less)), C.byref(a), C.byref(b)))
    return a.value, b.value
