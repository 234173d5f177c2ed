"""proteus_b200 - B200-native (sm_100a) per-pixel DSWx-HLS classification.

A drop-in for the per-pixel science path of nasa/PROTEUS v1.0.2
(``proteus.dswx_hls``): hand-written CUDA kernels behind a C ABI
(``include/proteus_b200.h``), reached from Python through ctypes.  PyTorch is
used for device buffers and streams only.  There is no CPU fallback: importing
this package without the built library, or calling it without a GPU, raises.

    import proteus_b200
    layers = proteus_b200.classify_tile(bands, fmask, dem, land, ocean,
                                        sun_azimuth, sun_elevation)
    proteus_b200.install()      # rebind proteus.dswx_hls helpers onto the GPU
"""
from . import _lib as _lib_module
from .params import (DEFAULT_PROCESSING, DEM_MARGIN_IN_PIXELS, HlsThresholds,
                     make_params)

_lib_module.load()              # fail loudly if the CUDA library is missing

from .engine import (ALL_LAYERS, GRADED_LAYERS, Context, Plan, TilePipeline,  # noqa: E402
                     classify_device, classify_tile, counters_to_dict,
                     get_context, pinned_copy, pinned_empty, wait_tile)
from .dswx_hls import REPLACED_FUNCTIONS, install, uninstall  # noqa: E402
from . import dswx_hls  # noqa: E402

__version__ = '0.1.0'
__all__ = ['ALL_LAYERS', 'GRADED_LAYERS', 'Context', 'Plan', 'classify_device',
           'classify_tile', 'wait_tile', 'TilePipeline', 'counters_to_dict', 'get_context', 'pinned_copy',
           'pinned_empty', 'HlsThresholds', 'make_params', 'install',
           'uninstall', 'dswx_hls', 'REPLACED_FUNCTIONS', 'DEFAULT_PROCESSING',
           'DEM_MARGIN_IN_PIXELS']
