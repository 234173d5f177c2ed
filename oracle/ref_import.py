"""TEST INFRASTRUCTURE ONLY - live import of the unmodified reference module.

Imports ``proteus.dswx_hls`` from ``/root/reference/src`` with stand-in modules
for the GDAL / yamale / ruamel imports the reference does at module scope
(``src/proteus/dswx_hls.py:8-13``; none of them is touched by the per-pixel
functions of SURVEY.md section 8a).  Nothing under ``/root/reference`` is
modified or copied.

``/root/reference`` exists only in the build container.  On the GPU box
``available()`` returns False and everything that needs the live reference is
skipped; the committed fixtures under ``tests/golden/`` (made by
``oracle/make_golden.py`` with this module) carry the reference's outputs there.

Only ``tests/`` and ``oracle/make_golden.py`` may import this file.
"""
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('PROTEUS_REFERENCE_ROOT', '/root/reference')
_REF = None


def available():
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, 'src', 'proteus', 'dswx_hls.py'))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Return the reference module object (cached)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    if 'osgeo' not in sys.modules:
        osgeo = _stub('osgeo')
        osgeo.gdal = _stub('osgeo.gdal', GDT_Byte=1, GDT_UInt16=2,
                           GDT_Float32=6)
        osgeo.osr = _stub('osgeo.osr')
        osgeo.ogr = _stub('osgeo.ogr')
        osgeo.gdalconst = _stub('osgeo.gdalconst', GDT_Float32=6, GDT_Byte=1)
    if 'yamale' not in sys.modules:
        _stub('yamale')
    if 'ruamel.yaml' not in sys.modules:
        ruamel = sys.modules.get('ruamel') or _stub('ruamel')
        ruamel.yaml = _stub('ruamel.yaml', YAML=object)
    src = os.path.join(REFERENCE_ROOT, 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    import proteus.dswx_hls as ref
    logging.getLogger('dswx_hls').setLevel(logging.CRITICAL)
    _REF = ref
    return ref


def default_runconfig_groups():
    """The 'groups' dict of the reference's packaged default runconfig."""
    import yaml
    path = os.path.join(REFERENCE_ROOT, 'src', 'proteus', 'defaults',
                        'dswx_hls.yaml')
    with open(path) as f:
        return yaml.safe_load(f)['runconfig']['groups']


class _Numpy1ScalarShim:
    """numpy 1.x value-based casting for the scalar x array expressions of _compute_opera_shadow_layer, obtained from
    the installed numpy >= 2: sin / cos / radians of a SCALAR return a Python float (same float64 value).  Under NEP 50
    a Python float is a weak scalar, so `float32_array * python_float` stays float32 with the scalar rounded to float32
    first - exactly what numpy 1.23.5 (the reference's pin, setup.py:78) does with a float64 scalar.  Scalar x scalar
    products stay float64 in both.  Everything else is numpy itself."""

    def __getattr__(self, name):
        import numpy
        return getattr(numpy, name)

    @staticmethod
    def _scalar(fn):
        import numpy

        def wrapped(x):
            r = fn(x)
            return float(r) if numpy.ndim(r) == 0 else r
        return wrapped

    def __init__(self):
        import numpy
        self.sin, self.cos, self.radians = (self._scalar(numpy.sin), self._scalar(numpy.cos),
                                            self._scalar(numpy.radians))


def live_shadow_layer_numpy1(dem, sun_azimuth_angle, sun_elevation_angle, min_slope_angle, max_sun_local_inc_angle,
                             **kw):
    """The UNMODIFIED code object of ``_compute_opera_shadow_layer`` (dswx_hls.py:4215-4283) executed with the module
    global ``np`` bound to the shim above: the reference under its own pinned numpy promotion rules."""
    ref = load()
    import numpy
    if numpy.lib.NumpyVersion(numpy.__version__) < '2.0.0':
        return ref._compute_opera_shadow_layer(dem, sun_azimuth_angle, sun_elevation_angle, min_slope_angle,
                                               max_sun_local_inc_angle, **kw)
    fn = ref._compute_opera_shadow_layer
    g = dict(fn.__globals__)
    g['np'] = _Numpy1ScalarShim()
    legacy = types.FunctionType(fn.__code__, g, fn.__name__, fn.__defaults__, fn.__closure__)
    return legacy(dem, sun_azimuth_angle, sun_elevation_angle, min_slope_angle, max_sun_local_inc_angle, **kw)


def live_create_landcover_mask(worldcover_up_3, copernicus, forest_classes, year, mask_type='standard'):
    """Run the UNMODIFIED ``create_landcover_mask`` (dswx_hls.py:911-1130) on in-memory rasters: its two
    ``_warp`` calls (GDAL) return the given arrays and the WorldCover metadata read returns ``year``;
    every numpy statement of the function (D:1003-1115) executes as written."""
    import tempfile
    ref = load()
    calls = []

    def fake_warp(*args, **kwargs):
        calls.append(1)
        return copernicus if len(calls) == 1 else worldcover_up_3

    class _Dataset:
        def GetMetadata(self):
            return {'time_start': f'{year}-01-01T00:00:00Z', 'time_end': f'{year}-12-31T23:59:59Z'}

    class _ColorTable:
        def SetColorEntry(self, *a):
            pass

    saved = ref._warp
    ref._warp = fake_warp
    ref.gdal.Open = lambda *a, **k: _Dataset()
    ref.gdal.ColorTable = _ColorTable
    ref.gdal.GA_ReadOnly = 0
    try:
        with tempfile.TemporaryDirectory() as d:
            f1, f2 = os.path.join(d, 'cgls.tif'), os.path.join(d, 'worldcover.tif')
            open(f1, 'w').close()
            open(f2, 'w').close()
            h, w = copernicus.shape
            return ref.create_landcover_mask(f1, f2, 'description', None, d, mask_type, (0, 30, 0, 0, 0, -30),
                                             'projection', h, w, forest_classes, temp_files_list=[])
    finally:
        ref._warp = saved


def live_save_output_rgb(red, green, blue, offset_dict, scale_dict, invalid_ind=None, flag_infrared=False):
    """Run the UNMODIFIED ``_save_output_rgb_file`` (dswx_hls.py:2960-3053) with a GDAL driver stand-in that
    captures the three arrays it writes: the float32 offset-and-scale statements (D:3024-3036) execute as written."""
    ref = load()
    written = []

    class _Band:
        def WriteArray(self, a):
            written.append(a.copy())

    class _Dataset:
        def SetMetadata(self, *a): pass
        def SetGeoTransform(self, *a): pass
        def SetProjection(self, *a): pass
        def GetRasterBand(self, i): return _Band()
        def FlushCache(self): pass

    class _Driver:
        def Create(self, *a, **k): return _Dataset()

    saved = (getattr(ref.gdal, 'GetDriverByName', None), ref.save_as_cog, ref._makedirs)
    ref.gdal.GetDriverByName = lambda name: _Driver()
    ref.save_as_cog = lambda *a, **k: None
    ref._makedirs = lambda *a, **k: None
    try:
        ref._save_output_rgb_file(red, green, blue, 'rgb.tif', offset_dict, scale_dict, False, {}, (0, 30, 0, 0, 0, -30),
                                  'projection', invalid_ind=invalid_ind, output_files_list=None,
                                  flag_infrared=flag_infrared)
    finally:
        if saved[0] is not None:
            ref.gdal.GetDriverByName = saved[0]
        ref.save_as_cog, ref._makedirs = saved[1], saved[2]
    return written
