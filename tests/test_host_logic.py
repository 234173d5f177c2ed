"""Host-side logic and the C-ABI surface (no GPU needed, no compute calls)."""
import ctypes as C
import math
import os
import re
from fractions import Fraction

import numpy as np
import pytest

from conftest import ROOT, has_cuda

HEADER = os.path.join(ROOT, 'include', 'proteus_b200.h')


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pb200_[a-z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_every_declared_symbol():
    from proteus_b200 import _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert sorted(_lib.EXPORTS) == declared, 'ctypes binding list drifted from the header'
    assert lib.pb200_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    """Compile a tiny C program against the header and compare sizeof/offsetof
    with the ctypes mirrors."""
    import subprocess
    import tempfile
    from proteus_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "proteus_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pb200_thresholds), sizeof(pb200_params),
         sizeof(pb200_tile), offsetof(pb200_params, aerosol_class_bits),
         offsetof(pb200_params, min_slope_angle), offsetof(pb200_tile, sun_terms),
         offsetof(pb200_tile, diag), offsetof(pb200_tile, counters));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    exp = [C.sizeof(_lib.Thresholds), C.sizeof(_lib.Params), C.sizeof(_lib.Tile),
           _lib.Params.aerosol_class_bits.offset, _lib.Params.min_slope_angle.offset,
           _lib.Tile.sun_terms.offset, _lib.Tile.diag.offset, _lib.Tile.counters.offset]
    assert got == exp


def _exact_bound(t, is_less):
    nb = np.nextafter(t, -np.inf if is_less else np.inf)
    mid = (Fraction(float(t)) + Fraction(float(nb))) / 2
    best = None
    for q in range(1, 32769):
        if not is_less:
            a = math.floor(mid * q) + 1
            if a > 32768:
                continue
            f = Fraction(max(a, -32768), q)
            if best is None or f < best:
                best = f
        else:
            a = math.ceil(mid * q) - 1
            if a < -32768:
                continue
            f = Fraction(min(a, 32768), q)
            if best is None or f > best:
                best = f
    return best


@pytest.mark.parametrize('t,is_less', [
    (0.124, 0), (-0.44, 0), (-0.5, 0), (0.7, 1), (0.0, 0), (0.0, 1), (1.0, 0), (-1.0, 1),
    (0.25, 0), (0.25, 1), (1 / 3, 0), (1 / 3, 1), (5.5, 1), (1e-300, 0), (-3e-20, 1),
    (32767.5, 0), (-32768.0, 0), (123.456, 0)])
def test_ratio_bound_is_the_exact_neighbour(t, is_less):
    """pb200_ratio_bound returns the fraction adjacent to the rounding
    midpoint among all p/q, |p| <= 32768, 1 <= q <= 32768 (exact rationals)."""
    from proteus_b200 import _lib
    a, b = _lib.ratio_bound(t, is_less)
    assert b > 0
    assert Fraction(a, b) == _exact_bound(t, is_less)


def test_ratio_bound_equals_numpy_division_on_random_pairs():
    from proteus_b200 import _lib
    rng = np.random.default_rng(0)
    n = rng.integers(-32768, 32768, 2_000_000).astype(np.int16)
    d = rng.integers(-32768, 32768, 2_000_000).astype(np.int16)
    d[:1000] = 0
    n[:10] = 0
    # ratios sitting exactly on the thresholds
    n[1000:1100], d[1000:1100] = 31, 250          # 0.124
    n[1100:1200], d[1100:1200] = -11, 25          # -0.44
    n[1200:1300], d[1200:1300] = 7, 10            # 0.7
    with np.errstate(all='ignore'):
        q = n / d                                  # int16 / int16 -> float64 (dswx_hls.py:1872)
    p = np.where(d < 0, -n.astype(np.int64), n.astype(np.int64))
    qq = np.abs(d.astype(np.int64))
    for t, less in ((0.124, 0), (-0.44, 0), (-0.5, 0), (0.7, 1), (0.0, 0), (0.5, 1)):
        a, b = _lib.ratio_bound(t, less)
        got = ((p * b <= a * qq) if less else (p * b >= a * qq)) & ((p | qq) != 0)
        ref = (q < t) if less else (q > t)
        assert np.array_equal(got, ref), (t, less)


def test_ratio_bound_degenerate_thresholds():
    from proteus_b200 import _lib
    assert _lib.ratio_bound(float('nan'), 0) == (1, 0)       # never
    assert _lib.ratio_bound(float('nan'), 1) == (-1, 0)
    assert _lib.ratio_bound(1e9, 0) == (1, 0)                # never  > 1e9
    assert _lib.ratio_bound(1e9, 1) == (1, 0)                # always < 1e9
    assert _lib.ratio_bound(-1e9, 0) == (-1, 0)              # always > -1e9
    assert _lib.ratio_bound(-1e9, 1) == (-1, 0)              # never  < -1e9


def test_angle_thresholds_are_numpy_decision_boundaries():
    from proteus_b200.params import angle_thresholds
    for min_slope, max_inc in ((-5, 40), (-5.0, 40.0), (0, 90), (-12.5, 33.3), (3, 10)):
        c, t = angle_thresholds(min_slope, max_inc)
        with np.errstate(invalid='ignore'):
            assert np.degrees(np.arccos(c)) <= max_inc
            assert not np.degrees(np.arccos(np.nextafter(c, -np.inf))) <= max_inc
            assert np.degrees(np.arctan(t)) <= min_slope
            assert not np.degrees(np.arctan(np.nextafter(t, np.inf))) <= min_slope
    c, t = angle_thresholds(-91, 181)
    assert c == -1.0 and math.isnan(t)
    c, t = angle_thresholds(90, -1)
    assert c == 2.0 and t == float('inf')


def test_float32_angle_thresholds_for_numpy1_promotion():
    """numpy 1.x promotion: the angle tests run on float32 arrays; the thresholds are float32 decision boundaries."""
    from proteus_b200 import _lib
    from proteus_b200.params import angle_thresholds, make_params, numpy1_promotion_default, set_numpy1_promotion
    f32 = np.float32
    for min_slope, max_inc in ((-5, 40), (-12.5, 33.3), (0, 90)):
        c, t = angle_thresholds(min_slope, max_inc, True)
        assert f32(c) == c and f32(t) == t
        with np.errstate(invalid='ignore'):
            arr = lambda x: np.array([x], dtype=f32)
            assert (np.degrees(np.arccos(arr(c))) <= max_inc)[0]
            assert not (np.degrees(np.arccos(arr(np.nextafter(f32(c), f32(-np.inf))))) <= max_inc)[0]
            assert (np.degrees(np.arctan(arr(t))) <= min_slope)[0]
            assert not (np.degrees(np.arctan(arr(np.nextafter(f32(t), f32(np.inf))))) <= min_slope)[0]
    c64, t64 = angle_thresholds(-5, 40)
    c32, t32 = angle_thresholds(-5, 40, True)
    assert abs(c32 - c64) < 1e-6 and abs(t32 - t64) < 1e-7
    # default follows the installed numpy (>= 2 here: float64), the override and the explicit argument win
    assert numpy1_promotion_default() == (np.lib.NumpyVersion(np.__version__) < '2.0.0')
    assert make_params().numpy1_promotion == int(numpy1_promotion_default())
    prev = set_numpy1_promotion(True)
    try:
        p = make_params()
        assert p.numpy1_promotion == 1 and (p.cos_inc_threshold, p.tan_slope_threshold) == (c32, t32)
        assert make_params(numpy1_promotion=False).numpy1_promotion == 0
    finally:
        set_numpy1_promotion(prev)
    # the library's own (libm float32) derivation lands within 2 float32 ulps of numpy's
    lib = _lib.load()
    q = _lib.Params()
    lib.pb200_params_default(C.byref(q))
    q.numpy1_promotion = 1
    c, t = C.c_double(), C.c_double()
    assert lib.pb200_angle_thresholds(C.byref(q), C.byref(c), C.byref(t)) == 0
    assert f32(c.value) == c.value and f32(t.value) == t.value
    assert abs(c.value - c32) <= 2 * np.spacing(f32(c32)) and abs(t.value - t32) <= 2 * np.spacing(f32(abs(t32)))


def test_library_libm_thresholds_agree_with_numpy_ones_here():
    """Not required for parity (Python always passes numpy's), but the libm
    derivation inside the library should land on the same float64."""
    from proteus_b200 import _lib
    from proteus_b200.params import angle_thresholds
    lib = _lib.load()
    p = _lib.Params()
    lib.pb200_params_default(C.byref(p))
    c, t = C.c_double(), C.c_double()
    assert lib.pb200_angle_thresholds(C.byref(p), C.byref(c), C.byref(t)) == 0
    nc, nt = angle_thresholds(-5, 40)
    assert abs(c.value - nc) <= 2 * np.spacing(nc) and abs(t.value - nt) <= 2 * np.spacing(abs(nt))


def test_params_default_matches_python_defaults():
    from proteus_b200 import _lib
    from proteus_b200.params import make_params, THRESHOLD_FIELDS
    lib = _lib.load()
    p = _lib.Params()
    lib.pb200_params_default(C.byref(p))
    q = make_params()
    for f in THRESHOLD_FIELDS:
        assert getattr(p.th, f) == getattr(q.th, f)
    assert list(p.band_fill) == list(q.band_fill) == [-9999] * 6
    assert p.fmask_fill == q.fmask_fill == 255
    assert bytes(p.aerosol_class_bits) == bytes(q.aerosol_class_bits)
    assert (p.adjacent_mode, p.apply_aerosol_class_remapping, p.collapse_wtr_classes) == \
           (q.adjacent_mode, q.apply_aerosol_class_remapping, q.collapse_wtr_classes)


def test_params_validation():
    from proteus_b200.params import make_params, HlsThresholds, aerosol_class_bits
    with pytest.raises(Exception, match='ERROR mask adjacent'):
        make_params(mask_adjacent_to_cloud_mode='bogus')
    with pytest.raises(OverflowError):
        make_params(HlsThresholds(lcmask_nir=70000))
    make_params(HlsThresholds(lcmask_nir=70000.0))            # floats compare in float64: fine
    bits = aerosol_class_bits([224, 300, -1], [224], [], [96])
    assert bits[224] == 0b00101 and bits[96] == 0b10000 and bits.sum() == 0b00101 + 0b10000
    p = make_params(band_fill=[-9999, -9999.5, 1e9, None, -1, 0], fmask_fill=None)
    from proteus_b200 import _lib
    assert list(p.band_fill) == [-9999, _lib.NO_FILL, _lib.NO_FILL, _lib.NO_FILL, -1, 0]
    assert p.fmask_fill == _lib.NO_FILL


def test_sun_terms_are_the_reference_expressions():
    from proteus_b200.params import sun_terms
    az, el = 157.25, 33.5
    a, z = np.radians(az), np.radians(90 - el)
    assert sun_terms(az, el) == (float(np.sin(a) * np.sin(z)), float(np.cos(a) * np.sin(z)),
                                 float(np.cos(z)), float(np.sin(a)), float(np.cos(a)))


def test_coverage_percentages_match_oracle():
    from proteus_b200.engine import counters_to_dict
    from oracle import dswx_oracle as O
    for c in ([0, 0, 0], [99, 98, 100], [1, 1, 3], [13395600, 5, 13395600]):
        total = max(c[2], 100)
        d = counters_to_dict(c + [0] * 9, total, has_ocean=True)
        assert (d['SPATIAL_COVERAGE'], d['SPATIAL_COVERAGE_EXCLUDING_MASKED_OCEAN'],
                d['CLOUD_COVERAGE']) == O.coverage_percentages(c[0], c[1], c[2], total)


@pytest.mark.skipif(has_cuda(), reason='checks the no-GPU failure mode')
def test_no_gpu_means_loud_failure_not_fallback():
    import proteus_b200
    from proteus_b200 import synth
    t = synth.make_tile(1, 16, 16)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        proteus_b200.classify_tile(t['bands'], t['fmask'])
    with pytest.raises(RuntimeError):
        proteus_b200.dswx_hls._get_binary_water_layer(t['fmask'])


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under proteus_b200/ may
    import or execute it."""
    pkg = os.path.join(ROOT, 'proteus_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'dswx_oracle' not in text, f


def test_install_swaps_exactly_the_replaced_functions_on_the_live_reference():
    """INTEGRATION level 2 in the build container: install() on the LIVE reference module rebinds exactly the names
    of REPLACED_FUNCTIONS, every one with the reference's own signature, and uninstall() puts the originals back."""
    import inspect
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('needs /root/reference (build container)')
    import proteus_b200
    import proteus_b200.dswx_hls as G
    ref = ref_import.load()
    originals = {n: getattr(ref, n) for n in G.REPLACED_FUNCTIONS}
    other = {n: v for n, v in vars(ref).items() if callable(v) and n not in G.REPLACED_FUNCTIONS}
    for n, fn in originals.items():
        assert inspect.signature(fn) == inspect.signature(getattr(G, n)), n
    try:
        assert proteus_b200.install(ref) is ref
        for n in G.REPLACED_FUNCTIONS:
            assert getattr(ref, n) is getattr(G, n), n
        for n, v in other.items():
            assert getattr(ref, n) is v, f'{n} must not be touched'
        # generate_dswx_layers resolves its helpers as module globals at call time (D:5089, 5161, 5225-5368)
        assert ref.generate_dswx_layers.__globals__['_compute_diagnostic_tests'] is G._compute_diagnostic_tests
    finally:
        proteus_b200.uninstall(ref)
    for n, fn in originals.items():
        assert getattr(ref, n) is fn, n
