"""Fused classification of HLS tiles on the GPU.

Replaces, in one pass per tile, the statement sequence of
``generate_dswx_layers`` between loading the rasters and saving the layers
(src/proteus/dswx_hls.py:5088-5369).

Two call styles:

* ``classify_tile(...)``     - numpy arrays in, numpy arrays out (what the
  reference's orchestrator holds after ``gdal.ReadAsArray``); goes through
  ``pb200_classify_host`` which overlaps H2D, the kernel and D2H per row strip.
* ``Plan`` / ``classify_device(...)`` - torch CUDA tensors in and out; tile
  descriptors live on the device, one launch per batch of tiles.

PyTorch is used for device buffers and streams only.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from .params import DEM_MARGIN_IN_PIXELS, make_params, sun_terms

LAYERS = ('DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD', 'SHAD', 'WTR',
          'BWTR', 'CONF')
_FIELD = {'DIAG': 'diag', 'WTR1': 'wtr1', 'WTR1_REMAPPED': 'wtr1_remapped',
          'WTR2': 'wtr2', 'CLOUD': 'cloud', 'SHAD': 'shad', 'WTR': 'wtr',
          'BWTR': 'bwtr', 'CONF': 'conf'}
GRADED_LAYERS = ('WTR', 'BWTR', 'CONF', 'DIAG')
ALL_LAYERS = LAYERS
HISTOGRAM_CLASSES = (0, 1, 2, 3, 4, 252, 253, 254, 255)


# ---------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------
class Context:
    """One ``pb200_ctx`` (per host thread and GPU)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('proteus_b200 needs a CUDA device (sm_100a); '
                               'there is no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(torch.device('cuda', device).index
                          if not isinstance(device, int) else device)
        self._lib = _lib.load()
        handle = C.c_void_p()
        _lib.check(self._lib.pb200_ctx_create(self.device, C.byref(handle)))
        self.handle = handle

    def close(self):
        if getattr(self, 'handle', None):
            self._lib.pb200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}
_contexts_lock = threading.Lock()


def get_context(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (threading.get_ident(), int(device))
    with _contexts_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = _contexts[key] = Context(int(device))
    return ctx


# ---------------------------------------------------------------------------
# pinned host memory
# ---------------------------------------------------------------------------
class _PinnedBlock:
    def __init__(self, nbytes):
        self.lib = _lib.load()
        p = C.c_void_p()
        _lib.check(self.lib.pb200_host_alloc(int(nbytes), C.byref(p)))
        self.ptr = p

    def __del__(self):
        try:
            if self.ptr:
                self.lib.pb200_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked memory (pb200_host_alloc)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    block = _PinnedBlock(max(n, 1))
    buf = (C.c_ubyte * max(n, 1)).from_address(block.ptr.value)
    buf._pb200_owner = block
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def pinned_copy(a):
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


# ---------------------------------------------------------------------------
# descriptor helpers
# ---------------------------------------------------------------------------
def _check_raster(name, a, dtype, shape=None):
    if a.dtype != dtype:
        raise TypeError(f'{name}: expected {np.dtype(dtype).name}, got {a.dtype}')
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f'{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}')


def _fill_tile(tile, *, height, width, band_ptrs, fmask_ptr, dem_ptr, dem_shape,
               dem_off, land_ptr, ocean_ptr, sun, out_ptrs, counters_ptr):
    tile.height, tile.width = int(height), int(width)
    for k in range(6):
        tile.band[k] = band_ptrs[k]
    tile.fmask = fmask_ptr
    tile.dem = dem_ptr
    if dem_ptr:
        tile.dem_rows, tile.dem_pitch = int(dem_shape[0]), int(dem_shape[1])
        tile.dem_off_y, tile.dem_off_x = int(dem_off[0]), int(dem_off[1])
    tile.land = land_ptr
    tile.ocean = ocean_ptr
    tile.sun_azimuth, tile.sun_elevation = float(sun[0]), float(sun[1])
    terms = sun_terms(sun[0], sun[1])
    for i in range(5):
        tile.sun_terms[i] = terms[i]
    for name in LAYERS:
        setattr(tile, _FIELD[name], out_ptrs.get(name))
    tile.counters = counters_ptr


def _dem_offsets(t, h, w):
    dem = t.get('dem')
    if dem is None:
        return None
    if 'dem_off' in t and t['dem_off'] is not None:
        return tuple(t['dem_off'])
    m = t.get('dem_margin', DEM_MARGIN_IN_PIXELS)
    if tuple(dem.shape) != (h + 2 * m, w + 2 * m):
        raise ValueError(f'DEM shape {tuple(dem.shape)} does not match the '
                         f'{h}x{w} tile with a {m}-pixel margin')
    if m < 1:
        raise ValueError('the DEM needs a margin of at least 1 pixel '
                         '(the reference warps it with 50, dswx_hls.py:58)')
    return (m, m)


def counters_to_dict(counters, total_pixels, has_ocean):
    """Raw counter slots -> the three floor-percentages of
    dswx_hls.py:5115-5124 (+ the optional class histogram)."""
    c = [int(v) for v in counters]
    n_valid, n_cloud_and_valid, n_not_ocean = c[0], c[1], c[2]
    if not has_ocean:
        n_not_ocean = total_pixels                                   # D:5107
    spatial = int(100 * float(n_valid) / total_pixels)
    cloud = 0 if n_valid == 0 else int(100 * float(n_cloud_and_valid) / n_valid)
    spatial_no_ocean = 0 if n_not_ocean == 0 else int(100 * float(n_valid) / n_not_ocean)
    return dict(n_valid=n_valid, n_cloud_and_valid=n_cloud_and_valid,
                n_not_ocean=n_not_ocean, SPATIAL_COVERAGE=spatial,
                SPATIAL_COVERAGE_EXCLUDING_MASKED_OCEAN=spatial_no_ocean,
                CLOUD_COVERAGE=cloud,
                class_histogram=dict(zip(HISTOGRAM_CLASSES, c[3:12])))


# ---------------------------------------------------------------------------
# device-resident batches
# ---------------------------------------------------------------------------
class Plan:
    """A batch of device-resident tiles with its descriptors on the GPU.

    ``tiles``: list of dicts with torch CUDA tensors: 'bands' (6 x int16
    [H, W]), 'fmask' (uint8), optional 'dem' (float32 with margin), 'land',
    'ocean' (uint8), 'sun_azimuth', 'sun_elevation', optional 'dem_margin'
    (default 50) or 'dem_off' = (row, col) of pixel (0, 0) inside the DEM.
    Output tensors are allocated here (``plan.outputs[i][layer]``) unless
    given in ``outputs_into``."""

    def __init__(self, tiles, params=None, outputs=GRADED_LAYERS, *,
                 counters=True, ctx=None, outputs_into=None, counters_into=None):
        import torch
        self.ctx = ctx or get_context()
        self.params = params if params is not None else make_params()
        self.tiles = tiles
        self.layers = tuple(outputs)
        for name in self.layers:
            if name not in _FIELD:
                raise ValueError(f'unknown layer {name!r}')
        dev = torch.device('cuda', self.ctx.device)
        n = len(tiles)
        self.outputs = []
        if counters_into is not None:
            # caller-owned slots, shape (n, 12) int64; rows may alias (several
            # pieces of one raster adding into the same counters)
            if counters_into.dtype != torch.int64 or tuple(counters_into.shape) != (n, _lib.N_COUNTERS):
                raise ValueError('counters_into: expected an int64 tensor of shape (n_tiles, 12)')
            self.counters = counters_into
            counters = True
        else:
            self.counters = (torch.zeros((n, _lib.N_COUNTERS), dtype=torch.int64, device=dev)
                             if counters else None)
        self._tile_array = (_lib.Tile * n)()
        for i, t in enumerate(tiles):
            bands, fmask = t['bands'], t['fmask']
            h, w = int(fmask.shape[0]), int(fmask.shape[1])
            for k, b in enumerate(bands):
                self._check_tensor(f'bands[{k}]', b, torch.int16, (h, w))
            self._check_tensor('fmask', fmask, torch.uint8, (h, w))
            dem, land, ocean = t.get('dem'), t.get('land'), t.get('ocean')
            if dem is not None:
                self._check_tensor('dem', dem, torch.float32)
            if land is not None:
                self._check_tensor('land', land, torch.uint8, (h, w))
            if ocean is not None:
                self._check_tensor('ocean', ocean, torch.uint8, (h, w))
            outs = {}
            for name in self.layers:
                if name == 'SHAD' and dem is None:
                    continue
                given = (outputs_into[i].get(name) if outputs_into else None)
                if given is not None:
                    want = ((torch.int16, torch.uint16) if name == 'DIAG' else (torch.uint8,))
                    if given.dtype not in want:
                        raise TypeError(f'{name}: expected {want[0]}, got {given.dtype}')
                    self._check_tensor(name, given, given.dtype, (h, w))
                    outs[name] = given
                else:
                    # DIAG is uint16 in the product; int16 storage, viewed as uint16 on the host
                    outs[name] = torch.empty((h, w), device=dev,
                                             dtype=torch.int16 if name == 'DIAG' else torch.uint8)
            self.outputs.append(outs)
            _fill_tile(self._tile_array[i], height=h, width=w,
                       band_ptrs=[b.data_ptr() for b in bands], fmask_ptr=fmask.data_ptr(),
                       dem_ptr=dem.data_ptr() if dem is not None else None,
                       dem_shape=dem.shape if dem is not None else None,
                       dem_off=_dem_offsets(t, h, w),
                       land_ptr=land.data_ptr() if land is not None else None,
                       ocean_ptr=ocean.data_ptr() if ocean is not None else None,
                       sun=(t.get('sun_azimuth', 0.0), t.get('sun_elevation', 90.0)),
                       out_ptrs={k: v.data_ptr() for k, v in outs.items()},
                       counters_ptr=self.counters[i].data_ptr() if counters else None)
        handle = C.c_void_p()
        _lib.check(self.ctx._lib.pb200_plan_create(
            self.ctx.handle, self._tile_array, n, C.byref(self.params), C.byref(handle)))
        self.handle = handle
        self.n_pixels = sum(int(t['fmask'].numel()) for t in tiles)

    @staticmethod
    def _check_tensor(name, t, dtype, shape=None):
        if not t.is_cuda:
            raise TypeError(f'{name}: expected a CUDA tensor')
        if t.dtype != dtype:
            if name.startswith('bands') and t.dtype.is_floating_point:
                raise NotImplementedError(
                    f'{name}: float reflectances (--offset-and-scale-inputs, '
                    'dswx_hls.py:2300-2302) are not part of the fused int16 path yet')
            raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
        if not t.is_contiguous():
            raise ValueError(f'{name}: expected a C-contiguous tensor')
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f'{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}')

    def run(self, stream=None):
        """Launch the fused kernel over the whole batch (asynchronous)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.ctx.device)
        _lib.check(self.ctx._lib.pb200_plan_run(self.handle, C.c_void_p(stream.cuda_stream)))

    @property
    def kernels(self):
        """Bit mask of the kernels this plan launches (``_lib.KERNEL_*``)."""
        m = C.c_int()
        _lib.check(self.ctx._lib.pb200_plan_kernels(self.handle, C.byref(m)))
        return m.value

    @property
    def kernel_name(self):
        m = self.kernels
        names = []
        if m & _lib.KERNEL_STREAM:
            kern = 'dswx_fused_stream_dyn_kernel' if m & _lib.KERNEL_STREAM_DYN else 'dswx_fused_stream_kernel'
            names.append(f'pb200::{kern}<{"true" if m & _lib.KERNEL_FAST8 else "false"}>')
        if m & _lib.KERNEL_FAST:
            names.append('pb200::dswx_fused_fast_kernel' + ('<FAST8>' if m & _lib.KERNEL_FAST8 else ''))
        if m & _lib.KERNEL_GENERIC:
            names.append('pb200::dswx_fused_kernel')
        return ' + '.join(names)

    def zero_counters(self):
        if self.counters is not None:
            self.counters.zero_()

    def results(self, i=None):
        """Host copies (numpy) of the outputs of tile ``i`` (or all tiles)."""
        import torch
        torch.cuda.synchronize(self.ctx.device)
        idx = range(len(self.tiles)) if i is None else [i]
        res = []
        for j in idx:
            d = {}
            for name, t in self.outputs[j].items():
                a = t.cpu().numpy()
                d[name] = a.view(np.uint16) if name == 'DIAG' else a
            if self.counters is not None:
                h, w = self.tiles[j]['fmask'].shape
                d['counters'] = self.counters[j].cpu().numpy().astype(np.uint64)
                d['coverage'] = counters_to_dict(
                    d['counters'], int(h) * int(w), self.tiles[j].get('ocean') is not None)
            res.append(d)
        return res if i is None else res[0]

    def close(self):
        if getattr(self, 'handle', None):
            self.ctx._lib.pb200_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_cover_staging = {}                # (context, raster shape, DEM shape) -> {name: device tensor}
_cover_plans = {}                  # (context, raster pointers, shapes, parameter bytes) -> Plan of the first phase


def _cached_cover_plan(ctx, tile, params, layers):
    """Plan of the 'cover' flow's first phase, kept between calls on the same device rasters (a time series re-runs it
    after overwriting the bands in place): building a plan - descriptors, tensor maps, item list, output rasters - was
    0.6 of the 1.08 ms the whole flow took (profiles/).  A handful of plans is kept; the outputs of a cached plan are
    overwritten by the next call that hits it."""
    def ptr(x):
        return (int(x.data_ptr()), tuple(x.shape)) if x is not None else None
    key = (id(ctx), tuple(ptr(b) for b in tile['bands']), ptr(tile['fmask']), ptr(tile.get('dem')), ptr(tile.get('land')),
           ptr(tile.get('ocean')), float(tile.get('sun_azimuth', 0.0)), float(tile.get('sun_elevation', 90.0)),
           tile.get('dem_margin'), tuple(tile['dem_off']) if tile.get('dem_off') is not None else None,
           bytes(params), layers)
    plan = _cover_plans.pop(key, None)
    if plan is None:
        plan = Plan([tile], params, layers, ctx=ctx)
        while len(_cover_plans) >= 8:
            _cover_plans.pop(next(iter(_cover_plans))).close()
    else:
        plan.zero_counters()
    _cover_plans[key] = plan           # most recently used last
    return plan


def classify_device_cover(tile, hls_thresholds=None, outputs=ALL_LAYERS, *, collapse_wtr_classes=True,
                          class_histogram=False, ctx=None, **processing):
    """``mask_adjacent_to_cloud_mode='cover'`` for one device-resident tile.

    The masked dilations of dswx_hls.py:2055-2078 need the whole WTR-2 / CLOUD rasters, so this mode
    runs in three steps on the device: (1) the fused kernel with the snow bit deferred (CLOUD =
    preliminary layer + aerosol bit, WTR-2 uncollapsed), (2) ``pb200_snow_to_cloud_cover`` (10 + 7
    masked dilation steps), (3) the point-wise tail (cloud masking, BWTR, CONF, collapsing).
    Returns a dict of torch tensors (+ 'counters')."""
    import torch
    ctx = ctx or get_context()
    lib = ctx._lib
    processing = dict(processing)
    processing.pop('mask_adjacent_to_cloud_mode', None)
    params = make_params(hls_thresholds, mask_adjacent_to_cloud_mode='ignore', defer_snow=True,
                         collapse_wtr_classes=False, **processing)
    has_dem = tile.get('dem') is not None
    phase1 = ['DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD'] + (['SHAD'] if has_dem else [])
    plan = _cached_cover_plan(ctx, tile, params, tuple(phase1))
    plan.run()
    o = plan.outputs[0]
    fmask = tile['fmask']
    h, w = int(fmask.shape[0]), int(fmask.shape[1])
    n = h * w
    stream = C.c_void_p(torch.cuda.current_stream(ctx.device).cuda_stream)
    scratch = torch.empty(4 * max(n, 1), dtype=torch.uint8, device=fmask.device)
    _lib.check(lib.pb200_snow_to_cloud_cover(ctx.handle, o['WTR2'].data_ptr(), o['CLOUD'].data_ptr(),
                                             fmask.data_ptr(), h, w, scratch.data_ptr(), stream))
    res = dict(DIAG=o['DIAG'], CLOUD=o['CLOUD'])
    if has_dem:
        res['SHAD'] = o['SHAD']
    wtr = torch.empty_like(o['WTR2'])
    bwtr = torch.empty_like(o['WTR2'])
    conf = torch.empty_like(o['WTR2'])
    counters = plan.counters[0].clone()
    if class_histogram:                                   # of the uncollapsed WTR (D:5104 ff. count before collapsing)
        _lib.check(lib.pb200_cover_tail(ctx.handle, o['WTR2'].data_ptr(), o['CLOUD'].data_ptr(), n, wtr.data_ptr(),
                                        None, None, None, None, 0, stream))
        hist = torch.bincount(wtr.reshape(-1).to(torch.int64), minlength=256)
        for k, cls in enumerate(HISTOGRAM_CLASSES):
            counters[3 + k] = hist[cls]
    # one pass: WTR / BWTR / CONF, and the collapsed WTR, WTR-1, WTR-1 (remapped), WTR-2 in place
    _lib.check(lib.pb200_cover_tail(ctx.handle, o['WTR2'].data_ptr(), o['CLOUD'].data_ptr(), n, wtr.data_ptr(),
                                    bwtr.data_ptr(), conf.data_ptr(), o['WTR1'].data_ptr(),
                                    o['WTR1_REMAPPED'].data_ptr(), int(bool(collapse_wtr_classes)), stream))
    res.update(BWTR=bwtr, CONF=conf, WTR=wtr, WTR1=o['WTR1'], WTR1_REMAPPED=o['WTR1_REMAPPED'], WTR2=o['WTR2'])
    res = {k: v for k, v in res.items() if k in outputs}
    res['counters'] = counters
    return res


def classify_device(tiles, params=None, outputs=GRADED_LAYERS, *, stream=None, ctx=None):
    """One-shot: build a plan, run it once, return the plan (outputs on device)."""
    plan = Plan(tiles, params, outputs, ctx=ctx)
    plan.run(stream)
    return plan


# ---------------------------------------------------------------------------
# host-buffer path (the call a generate_dswx_layers drop-in makes)
# ---------------------------------------------------------------------------
def classify_tile(bands, fmask, dem_with_margin=None, landcover_mask=None,
                  ocean_mask=None, sun_azimuth_angle=0.0,
                  sun_elevation_angle=90.0, hls_thresholds=None, *,
                  outputs=ALL_LAYERS, params=None, dem_margin=DEM_MARGIN_IN_PIXELS,
                  dem_off=None, out=None, strip_rows=0, ctx=None, reuse_ancillary=False,
                  slot=0, wait=True, **processing):
    """Classify one tile held in host memory.

    Parameters mirror what ``generate_dswx_layers`` has in hand at
    dswx_hls.py:5088: ``bands`` = (blue, green, red, nir, swir1, swir2) int16
    arrays as read from the HLS files (before the clip of D:2299), ``fmask``
    uint8, the float32 DEM with its margin (D:5145-5150), the LAND array
    (D:5196) and the ocean mask (D:5097) or ``None``.  ``processing`` takes the
    runconfig keys ``mask_adjacent_to_cloud_mode``,
    ``apply_aerosol_class_remapping``, ``aerosol_fmask_values``,
    ``min_slope_angle``, ``max_sun_local_inc_angle``, ``band_fill``,
    ``fmask_fill``, ``collapse_wtr_classes``, ``class_histogram``.

    ``reuse_ancillary=True`` (time series: the next acquisition of the same MGRS tile): the DEM, LAND and ocean
    rasters the previous call on this context uploaded are used again, only bands and Fmask are copied; the arrays
    must still be passed (same shapes) and be unchanged.

    ``slot`` (0 / 1) selects one of the context's two host pipelines; ``wait=False`` returns as soon as the tile is
    enqueued - the arrays of the returned dict are complete only after ``wait_tile(result)`` (``TilePipeline`` does
    this bookkeeping for a stream of tiles: two tiles in flight hide the un-overlapped head and tail of each).

    Returns a dict: requested layers (numpy, pinned), 'counters' (uint64[12])
    and 'coverage' (the three percentages of D:5115-5124)."""
    ctx = ctx or get_context()
    if params is None and processing.get('mask_adjacent_to_cloud_mode') == 'cover':
        return _classify_tile_cover(bands, fmask, dem_with_margin, landcover_mask, ocean_mask,
                                    sun_azimuth_angle, sun_elevation_angle, hls_thresholds,
                                    outputs=outputs, dem_margin=dem_margin, dem_off=dem_off, ctx=ctx,
                                    **processing)
    if params is None:
        params = make_params(hls_thresholds, **processing)
    elif processing or hls_thresholds is not None:
        raise ValueError('give either params or thresholds/processing options')
    fmask = np.ascontiguousarray(fmask)
    h, w = fmask.shape
    _check_raster('fmask', fmask, np.uint8)
    bands = [np.ascontiguousarray(b) for b in bands]
    if len(bands) != 6:
        raise ValueError('bands: expected blue, green, red, nir, swir1, swir2')
    for k, b in enumerate(bands):
        if np.issubdtype(b.dtype, np.floating):
            raise NotImplementedError(
                'float reflectances (--offset-and-scale-inputs, dswx_hls.py:2300-2302) '
                'are not part of the fused int16 path yet')
        _check_raster(f'bands[{k}]', b, np.int16, (h, w))
    t = dict(dem=dem_with_margin, dem_margin=dem_margin, dem_off=dem_off)
    if dem_with_margin is not None:
        dem_with_margin = np.ascontiguousarray(dem_with_margin)
        if dem_with_margin.dtype != np.float32:
            raise NotImplementedError(
                f'DEM dtype {dem_with_margin.dtype}: only the float32 DEM produced by the cubic '
                'warp (dswx_hls.py:5145-5150) is supported')
        t['dem'] = dem_with_margin
    if landcover_mask is not None:
        landcover_mask = np.ascontiguousarray(landcover_mask)
        _check_raster('landcover_mask', landcover_mask, np.uint8, (h, w))
    if ocean_mask is not None:
        ocean_mask = np.ascontiguousarray(ocean_mask)
        _check_raster('ocean_mask', ocean_mask, np.uint8, (h, w))

    res = {}
    for name in outputs:
        if name not in _FIELD:
            raise ValueError(f'unknown layer {name!r}')
        if name == 'SHAD' and dem_with_margin is None:
            continue
        dt = np.uint16 if name == 'DIAG' else np.uint8
        if out is not None and name in out:
            _check_raster(name, out[name], dt, (h, w))
            res[name] = out[name]
        else:
            res[name] = pinned_empty((h, w), dt)
    counters = (out['counters'] if out is not None and 'counters' in out
                else pinned_empty((_lib.N_COUNTERS,), np.uint64))
    tile = _lib.Tile()
    _fill_tile(tile, height=h, width=w,
               band_ptrs=[b.ctypes.data for b in bands], fmask_ptr=fmask.ctypes.data,
               dem_ptr=dem_with_margin.ctypes.data if dem_with_margin is not None else None,
               dem_shape=dem_with_margin.shape if dem_with_margin is not None else None,
               dem_off=_dem_offsets(t, h, w),
               land_ptr=landcover_mask.ctypes.data if landcover_mask is not None else None,
               ocean_ptr=ocean_mask.ctypes.data if ocean_mask is not None else None,
               sun=(sun_azimuth_angle, sun_elevation_angle),
               out_ptrs={k: v.ctypes.data for k, v in res.items()},
               counters_ptr=counters.ctypes.data)
    if slot not in (0, 1):
        raise ValueError('slot: 0 or 1')
    flags = ((_lib.HOST_REUSE_ANCILLARY if reuse_ancillary else 0) | (_lib.HOST_SLOT1 if slot else 0) |
             (0 if wait else _lib.HOST_ASYNC))
    _lib.check(ctx._lib.pb200_classify_host_ex(ctx.handle, C.byref(tile), C.byref(params), int(strip_rows), flags))
    res['counters'] = counters
    if wait:
        res['coverage'] = counters_to_dict(counters, h * w, ocean_mask is not None)
    else:
        # keep every input alive (and unmodified) until the wait; 'coverage' is filled in by wait_tile
        res['_pending'] = dict(ctx=ctx, slot=slot, total=h * w, has_ocean=ocean_mask is not None,
                               keep=(bands, fmask, dem_with_margin, landcover_mask, ocean_mask, tile))
    return res


def wait_tile(res):
    """Complete a tile enqueued with ``classify_tile(..., wait=False)``; returns the same dict, now final."""
    pend = res.pop('_pending', None)
    if pend is not None:
        _lib.check(pend['ctx']._lib.pb200_host_wait(pend['ctx'].handle, pend['slot']))
        res['coverage'] = counters_to_dict(res['counters'], pend['total'], pend['has_ocean'])
    return res


class TilePipeline:
    """A stream of host tiles with TWO in flight (the context's two host pipelines): while tile k is classified and
    its layers travel back, the rasters of tile k + 1 already cross PCIe.

        pipe = TilePipeline()
        for args in tiles:                      # args / kwargs of classify_tile
            done = pipe.submit(*args, out=...)  # -> the result of the tile submitted two calls ago, or None
        last = pipe.flush()                     # -> the results still in flight, oldest first

    Every tile in flight needs its own output buffers (``out=``), inputs may be shared."""

    def __init__(self, ctx=None):
        self.ctx = ctx or get_context()
        self._inflight = [None, None]
        self._n = 0

    def submit(self, *args, **kwargs):
        slot = self._n & 1
        self._n += 1
        done = self._inflight[slot]
        if done is not None:
            wait_tile(done)
        self._inflight[slot] = classify_tile(*args, ctx=self.ctx, slot=slot, wait=False, **kwargs)
        return done

    def flush(self):
        order = [(self._n - 2) & 1, (self._n - 1) & 1] if self._n >= 2 else [0]
        out = []
        for slot in order:
            r = self._inflight[slot]
            if r is not None:
                out.append(wait_tile(r))
                self._inflight[slot] = None
        return out


def _classify_tile_cover(bands, fmask, dem_with_margin, landcover_mask, ocean_mask, sun_azimuth_angle,
                         sun_elevation_angle, hls_thresholds, *, outputs, dem_margin, dem_off, ctx,
                         **processing):
    """Host-array front end of ``classify_device_cover`` (whole rasters go to the device: the dilations
    of the 'cover' mode are not strip-local)."""
    import torch
    dev = torch.device('cuda', ctx.device)

    # persistent device staging rasters per (context, shape): the same pointers on every call, so that the cached plan
    # of the first phase is found again (_cached_cover_plan)
    stage = _cover_staging.setdefault((id(ctx), tuple(np.shape(fmask)),
                                       tuple(np.shape(dem_with_margin)) if dem_with_margin is not None else None), {})

    def up(a, dtype, name=None):
        if a is None:
            return None
        a = np.ascontiguousarray(a)
        if a.dtype != dtype:
            raise TypeError(f'expected {np.dtype(dtype).name}, got {a.dtype}')
        src = torch.from_numpy(a)
        dst = stage.get(name)
        if dst is None or tuple(dst.shape) != tuple(src.shape):
            dst = stage[name] = torch.empty(src.shape, dtype=src.dtype, device=dev)
        dst.copy_(src, non_blocking=False)
        return dst
    tile = dict(bands=[up(b, np.int16, f'band{k}') for k, b in enumerate(bands)], fmask=up(fmask, np.uint8, 'fmask'),
                dem=up(dem_with_margin, np.float32, 'dem'), land=up(landcover_mask, np.uint8, 'land'),
                ocean=up(ocean_mask, np.uint8, 'ocean'), sun_azimuth=sun_azimuth_angle,
                sun_elevation=sun_elevation_angle, dem_margin=dem_margin, dem_off=dem_off)
    collapse = processing.pop('collapse_wtr_classes', True)
    hist = processing.pop('class_histogram', False)
    res = classify_device_cover(tile, hls_thresholds, outputs, collapse_wtr_classes=collapse,
                                class_histogram=hist, ctx=ctx, **processing)
    torch.cuda.synchronize(dev)
    out = {}
    for k, v in res.items():
        a = v.cpu().numpy()
        out[k] = a.view(np.uint16) if k == 'DIAG' else (a.astype(np.uint64) if k == 'counters' else a)
    h, w = fmask.shape
    out['coverage'] = counters_to_dict(out['counters'], h * w, ocean_mask is not None)
    return out
