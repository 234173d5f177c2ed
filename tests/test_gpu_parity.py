"""GPU parity: the CUDA path (through the C ABI) against the committed
reference fixtures and against the CPU oracle on seeded inputs.

Bar: bit-exact (np.array_equal) on every uint8 / uint16 layer and on the three
coverage counters - the reference's own comparison rule for integer layers
(compare_dswx_hls_products, dswx_hls.py:753-755)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import dswx_oracle as O
from proteus_b200 import synth

pytestmark = pytest.mark.gpu

GPU_CASES = ('full_default', 'full_adversarial', 'ignore_noaerosol',
             'l30_minimal', 'ragged_adversarial', 'shadow_only', 'guard_band')
FUSED_LAYERS = ('DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD', 'SHAD', 'WTR',
                'BWTR', 'CONF')


def _assert_layers(got, ref, names, what):
    for name in names:
        if name not in ref:
            continue
        g, r = got[name], ref[name]
        assert g.dtype == r.dtype, (what, name, g.dtype, r.dtype)
        if not np.array_equal(g, r):
            bad = np.argwhere(g != r)
            y, x = bad[0]
            raise AssertionError(
                f'{what}: layer {name}: {len(bad)} of {g.size} pixels differ; first at '
                f'({y},{x}): got {g[y, x]}, reference {r[y, x]}')


def _classify_host(pb, ins, collapse, **kw):
    return pb.classify_tile(
        ins['bands'], ins['fmask'], ins['dem'], ins['land'], ins['ocean'],
        ins['sun_azimuth'], ins['sun_elevation'],
        mask_adjacent_to_cloud_mode=ins['mode'],
        apply_aerosol_class_remapping=ins['aerosol'],
        collapse_wtr_classes=collapse, dem_margin=ins['dem_margin'], **kw)


@pytest.mark.parametrize('case', GPU_CASES)
def test_fused_host_path_matches_reference_fixture(pb, case):
    ins, ref = load_golden(case)
    got = _classify_host(pb, ins, collapse=False, class_histogram=True)
    _assert_layers(got, ref, FUSED_LAYERS, case)
    assert np.array_equal(got['counters'][:3], ref['counters'])
    cov = got['coverage']
    assert [cov['SPATIAL_COVERAGE'], cov['SPATIAL_COVERAGE_EXCLUDING_MASKED_OCEAN'],
            cov['CLOUD_COVERAGE']] == ref['percentages'].tolist()
    hist = np.bincount(ref['WTR'].ravel(), minlength=256)
    assert [cov['class_histogram'][c] for c in (0, 1, 2, 3, 4, 252, 253, 254, 255)] == \
        [int(hist[c]) for c in (0, 1, 2, 3, 4, 252, 253, 254, 255)]
    # collapsed variants as saved by save_dswx_product (dswx_hls.py:2688-2689)
    got_c = _classify_host(pb, ins, collapse=True)
    for name, key in (('WTR', 'WTR_COLLAPSED'), ('WTR1', 'WTR1_COLLAPSED'), ('WTR2', 'WTR2_COLLAPSED')):
        assert np.array_equal(got_c[name], ref[key]), (case, key)
    for name in ('BWTR', 'CONF', 'DIAG', 'CLOUD'):
        assert np.array_equal(got_c[name], ref[name]), (case, name)


@pytest.mark.parametrize('case', GPU_CASES)
def test_fused_device_path_matches_reference_fixture(pb, case):
    import torch
    ins, ref = load_golden(case)
    dev = 'cuda'
    tile = dict(bands=[torch.from_numpy(b).to(dev) for b in ins['bands']],
                fmask=torch.from_numpy(ins['fmask']).to(dev),
                dem=torch.from_numpy(ins['dem']).to(dev) if ins['dem'] is not None else None,
                land=torch.from_numpy(ins['land']).to(dev) if ins['land'] is not None else None,
                ocean=torch.from_numpy(ins['ocean']).to(dev) if ins['ocean'] is not None else None,
                sun_azimuth=ins['sun_azimuth'], sun_elevation=ins['sun_elevation'],
                dem_margin=ins['dem_margin'])
    params = pb.make_params(mask_adjacent_to_cloud_mode=ins['mode'],
                            apply_aerosol_class_remapping=ins['aerosol'],
                            collapse_wtr_classes=False)
    layers = [n for n in FUSED_LAYERS if n != 'SHAD' or ins['dem'] is not None]
    plan = pb.Plan([tile, tile], params, layers)       # two descriptors, same inputs
    plan.run()
    for i in (0, 1):
        got = plan.results(i)
        _assert_layers(got, ref, FUSED_LAYERS, f'{case}[{i}]')
        assert np.array_equal(got['counters'][:3], ref['counters'])
    # a second run ADDS to the counters (documented contract)
    plan.run()
    assert np.array_equal(plan.results(0)['counters'][:3], 2 * ref['counters'])


@pytest.mark.parametrize('seed,h,w,kw', [
    (11, 96, 128, {}),                                       # vector path, exact tile multiple
    (12, 70, 132, dict(with_ocean=False)),                   # W % 4 == 0, ragged tile edges
    (13, 33, 129, dict(adversarial=True)),                   # generic path (W % 4 != 0)
    (14, 1, 5, dict(adversarial=True, with_dem=False)),      # single row
    (15, 130, 1, dict(with_land=False)),                     # single column
    (16, 64, 260, dict(adversarial=True, with_land=False, with_ocean=False)),
])
def test_fused_matches_oracle_on_seeded_tiles(pb, seed, h, w, kw):
    t = synth.make_tile(seed, h, w, **kw)
    for mode, aerosol in (('mask', True), ('ignore', False)):
        ref = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                                t['sun_azimuth'], t['sun_elevation'],
                                processing=dict(mask_adjacent_to_cloud_mode=mode,
                                                apply_aerosol_class_remapping=aerosol))
        got = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                               t['sun_azimuth'], t['sun_elevation'],
                               mask_adjacent_to_cloud_mode=mode,
                               apply_aerosol_class_remapping=aerosol,
                               collapse_wtr_classes=False)
        _assert_layers(got, ref, FUSED_LAYERS, f'seed {seed} {mode}')
        assert np.array_equal(got['counters'][:3], ref['counters'])


def test_non_default_thresholds_and_fills(pb):
    """Thresholds that sit exactly on representable ratios (0.25, 1/3, 0.5),
    non-integral band thresholds and a different fill value."""
    t = synth.make_tile(21, 128, 256, adversarial=True)
    # plant exact ratios: green=5, swir1=3 -> mndwi = 0.25 exactly
    t['bands'][1][0, :64] = 5
    t['bands'][4][0, :64] = 3
    t['bands'][3][1, :64] = 2       # nir=2, red=1 -> ndvi = 1/3
    t['bands'][2][1, :64] = 1
    for b in t['bands']:
        b[5, 5] = -32768
    kw = dict(wigt=0.25, awgt=-0.25, pswt_1_mndwi=1 / 3, pswt_1_nir=1500.5,
              pswt_1_swir1=899.99, pswt_1_ndvi=1 / 3, pswt_2_mndwi=-0.5,
              pswt_2_blue=1000.0, pswt_2_nir=2500, pswt_2_swir1=3000,
              pswt_2_swir2=1000, lcmask_nir=1199.5)
    ref = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                            t['sun_azimuth'], t['sun_elevation'],
                            thresholds=O.HlsThresholds(**kw), band_fill=-32768, fmask_fill=7)
    got = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                           t['sun_azimuth'], t['sun_elevation'],
                           hls_thresholds=pb.HlsThresholds(**kw), band_fill=-32768,
                           fmask_fill=7, collapse_wtr_classes=False)
    _assert_layers(got, ref, FUSED_LAYERS, 'custom thresholds')
    assert np.array_equal(got['counters'][:3], ref['counters'])


def test_random_parameter_sets_through_the_dynamic_kernel(pb):
    """Random thresholds (simple fractions that select the FAST8 variant and arbitrary decimals that do not), fills,
    adjacent-cloud modes and angle limits over adversarial tiles (40 % of the pixels wrap an int16 sum) and regular
    ones, through the device-resident plan (dswx_fused_stream_dyn_kernel) - against the oracle."""
    import torch
    from proteus_b200 import _lib
    rng = np.random.default_rng(4242)
    seen_fast8 = set()
    for trial in range(14):
        simple = trial % 2 == 0
        frac = lambda lo, hi: float(rng.integers(int(lo * 16), int(hi * 16) + 1)) / 16.0     # k / 16: byte-sized bounds
        dec = lambda lo, hi: float(np.round(rng.uniform(lo, hi), 3))
        pick = frac if simple else dec
        kw = dict(wigt=pick(0.0, 0.5), awgt=float(rng.choice([0.0, 0.25, -0.5, 1.75])),
                  pswt_1_mndwi=pick(-0.75, -0.25), pswt_1_nir=float(rng.integers(800, 2500)),
                  pswt_1_swir1=float(rng.integers(500, 1500)), pswt_1_ndvi=pick(0.5, 0.875),
                  pswt_2_mndwi=pick(-0.875, -0.25), pswt_2_blue=float(rng.integers(500, 2000)),
                  pswt_2_nir=float(rng.integers(1500, 3500)), pswt_2_swir1=float(rng.integers(2000, 4000)),
                  pswt_2_swir2=float(rng.integers(500, 2000)), lcmask_nir=float(rng.integers(800, 2000)))
        mode = str(rng.choice(['mask', 'ignore']))
        aerosol = bool(rng.random() < 0.7)
        slope, inc = float(rng.choice([-5, -8, -2.5])), float(rng.choice([40, 35, 55]))
        fill = int(rng.choice([-9999, -32768, -1000]))
        tiles, refs = [], []
        for j in range(3):
            t = synth.make_tile(3000 + 10 * trial + j, 4 * int(rng.integers(8, 48)), 4 * int(rng.integers(9, 70)),
                                adversarial=(j != 1))
            for b in t['bands']:
                b[b == -9999] = fill
            refs.append(O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                                          t['sun_elevation'], thresholds=O.HlsThresholds(**kw), band_fill=fill,
                                          processing=dict(mask_adjacent_to_cloud_mode=mode, apply_aerosol_class_remapping=aerosol,
                                                          min_slope_angle=slope, max_sun_local_inc_angle=inc)))
            dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in t.items() if k != 'bands'}
            dev['bands'] = [torch.from_numpy(b).cuda() for b in t['bands']]
            tiles.append(dev)
        params = pb.make_params(pb.HlsThresholds(**kw), mask_adjacent_to_cloud_mode=mode, apply_aerosol_class_remapping=aerosol,
                                min_slope_angle=slope, max_sun_local_inc_angle=inc, band_fill=fill, collapse_wtr_classes=False)
        plan = pb.Plan(tiles, params, pb.GRADED_LAYERS)
        assert plan.kernels & _lib.KERNEL_STREAM_DYN
        seen_fast8.add(bool(plan.kernels & _lib.KERNEL_FAST8))
        plan.run()
        for i, ref in enumerate(refs):
            res = plan.results(i)
            _assert_layers(res, ref, ('DIAG', 'WTR', 'BWTR', 'CONF'), f'trial {trial} {kw} tile {i}')
            assert np.array_equal(res['counters'][:3], ref['counters']), (trial, i)
    assert seen_fast8 == {True, False}, 'both kernel variants should have been exercised'


@pytest.mark.parametrize('t,is_less', [(0.124, 0), (-0.44, 0), (-0.5, 0), (0.7, 1),
                                       (0.0, 0), (0.25, 1), (1 / 3, 0), (-1.0, 1)])
def test_ratio_test_exhaustive_int16_pairs(pb, t, is_less):
    """All 2^32 (n, d) int16 pairs: the integer form of float64(n)/float64(d)
    {>,<} t equals IEEE float64 division (incl. d == 0 -> inf / nan)."""
    import ctypes as C
    ctx = pb.get_context()
    bad = C.c_uint64(123)
    from proteus_b200 import _lib
    _lib.check(ctx._lib.pb200_ratio_sweep(ctx.handle, float(t), int(is_less), C.byref(bad)))
    assert bad.value == 0


def test_fast8_integer_forms_exhaustive(pb):
    """The FAST8 kernel variant evaluates each rational test as ONE IDP.2A on a per-pixel pack and repairs pixels whose
    int16 sums wrapped by sign bookkeeping (no division anywhere).  Checked here against numpy's arithmetic - wrapping
    int16 sums, IEEE float64 quotient (D:1872-1914) - over EVERY clipped (green, swir1) and (nir, red) pair in
    [1, 32767]^2 (2^30 pairs each; half of them wrap), and 4*awesh on all those pairs with pseudo-random other bands."""
    import ctypes as C
    from proteus_b200 import _lib
    ctx = pb.get_context()
    for th in (None, pb.HlsThresholds(wigt=0.2, awgt=0.25, pswt_1_mndwi=-0.3, pswt_1_ndvi=0.65, pswt_2_mndwi=-0.55)):
        params = pb.make_params(th)
        counts = (C.c_uint64 * 6)()
        rc = ctx._lib.pb200_fast8_sweep(ctx.handle, C.byref(params), counts)
        if th is not None and rc == _lib.E_UNSUPPORTED:
            continue                                   # these thresholds do not have byte-sized bounds: no FAST8 kernel
        _lib.check(rc)
        assert [int(c) for c in counts[:5]] == [0, 0, 0, 0, 0], [int(c) for c in counts]
        assert int(counts[5]) > 2 ** 28               # wrapped pairs were visited
    # parameters that are not FAST8-shaped are refused, not silently approximated
    params = pb.make_params(pb.HlsThresholds(wigt=0.123456789))
    counts = (C.c_uint64 * 6)()
    assert ctx._lib.pb200_fast8_sweep(ctx.handle, C.byref(params), counts) == _lib.E_UNSUPPORTED


def test_shadow_shortcuts_never_decide_wrongly(pb):
    """the float32 shadow shortcuts proved the way the ratio test was - by brute force on the GPU.  For
    five sun geometries: 2^32 random DEM neighbourhoods + 2^30 / 2^28 planted on the decision boundaries (slope:
    diff = +-e; incidence: D = +-eg, on rays that face away from the sun) + special values, through shadow_fast
    (compares), shadow_fast1 / shadow_fast2 (sign bits, FAST8) and the exact float64 sequence (D:4264-4281).  A decided
    sample must never differ from the exact one."""
    import ctypes as C
    import json
    import os
    from proteus_b200 import _lib
    from proteus_b200.params import sun_terms
    ctx = pb.get_context()
    params = pb.make_params()
    report = {}
    for az, el in ((150.0, 45.0), (10.0, 5.0), (359.0, 89.0), (200.0, 15.0), (135.0, 70.0)):
        terms = (C.c_double * 5)(*sun_terms(az, el))
        for mode, n in ((0, 2 ** 32), (1, 2 ** 30), (2, 2 ** 28), (3, 2 ** 24)):
            counts = (C.c_uint64 * 8)()
            _lib.check(ctx._lib.pb200_shadow_sweep(ctx.handle, C.byref(params), az, el, terms, mode, 1234 + mode, n, counts))
            n_s, shadow, dec_c, bad_c, dec_s, bad_s, dec_s2, bad_s2 = [int(v) for v in counts]
            assert n_s >= n
            assert bad_c == 0 and bad_s == 0 and bad_s2 == 0, (az, el, mode, bad_c, bad_s, bad_s2)
            report[f'az {az} el {el} mode {mode}'] = dict(samples=n_s, shadow_share=shadow / n_s, undecided_compare=1 - dec_c / n_s,
                                                         undecided_sign=1 - dec_s / n_s, undecided_sign_packed=1 - dec_s2 / n_s)
        r0, r1 = report[f'az {az} el {el} mode 0'], report[f'az {az} el {el} mode 1']
        # the generators do what they say: random gradients are almost always decided, with both outcomes; the ones
        # planted on the slope boundary often fall inside the guard band; special values mostly go to the exact sequence
        assert r0['undecided_compare'] < 1e-3 and r0['undecided_sign'] < 1e-3
        assert 0.02 < r0['shadow_share'] < 0.98
        assert r1['undecided_sign'] > 0.01 and r1['undecided_compare'] > 0.01
        assert report[f'az {az} el {el} mode 3']['undecided_sign'] > 0.3
    # the incidence boundary can only be met by a back slope when the sun stands high (zenith < 40 degrees)
    assert report['az 359.0 el 89.0 mode 2']['undecided_sign'] > 0.01
    assert report['az 135.0 el 70.0 mode 2']['undecided_sign'] > 0.01
    os.makedirs(os.path.join(os.path.dirname(__file__), '..', 'gpurun_out'), exist_ok=True)
    with open(os.path.join(os.path.dirname(__file__), '..', 'gpurun_out', 'shadow_sweep_report.json'), 'w') as f:
        json.dump(report, f, indent=1)


def test_guard_band_fixture_under_both_numpy_promotions(pb):
    """VERDICT r1 #3 + ADVICE (numpy 1.23.5 pin): a DEM whose pixels sit ON the two decision boundaries of the
    terrain-shadow test (about half of them inside the guard bands of the float32 shortcut, so the exact sequence runs
    there) against the live reference's layers - under numpy >= 2 promotion (float64 dot product, the fixture's out_*)
    and under numpy 1.x promotion (float32, out_*_NUMPY1: the same reference code object with numpy-1 scalar casting)."""
    import torch
    ins, ref = load_golden('guard_band')
    assert int((ref['SHAD'] != ref['SHAD_NUMPY1']).sum()) > 500          # the two modes really differ on this DEM
    for np1, sfx in ((False, ''), (True, '_NUMPY1')):
        got = _classify_host(pb, ins, collapse=False, numpy1_promotion=np1)           # fast kernel, all layers
        for name in ('SHAD', 'WTR2', 'CLOUD', 'WTR', 'BWTR', 'CONF'):
            assert np.array_equal(got[name], ref[name + sfx]), (name, np1, int((got[name] != ref[name + sfx]).sum()))
        assert np.array_equal(got['DIAG'], ref['DIAG']) and np.array_equal(got['WTR1'], ref['WTR1'])
        # graded layers through the TMA-fed stream kernel (device-resident plan)
        tile = dict(bands=[torch.from_numpy(b).cuda() for b in ins['bands']], fmask=torch.from_numpy(ins['fmask']).cuda(),
                    dem=torch.from_numpy(ins['dem']).cuda(), land=torch.from_numpy(ins['land']).cuda(),
                    ocean=torch.from_numpy(ins['ocean']).cuda(), sun_azimuth=ins['sun_azimuth'],
                    sun_elevation=ins['sun_elevation'], dem_margin=ins['dem_margin'])
        params = pb.make_params(mask_adjacent_to_cloud_mode=ins['mode'], apply_aerosol_class_remapping=ins['aerosol'],
                                collapse_wtr_classes=False, numpy1_promotion=np1)
        plan = pb.Plan([tile], params, pb.GRADED_LAYERS)
        assert 'stream' in plan.kernel_name
        plan.run()
        res = plan.results(0)
        for name in ('WTR', 'BWTR', 'CONF'):
            assert np.array_equal(res[name], ref[name + sfx]), (name, np1)
        # the function-level drop-in follows the override
        import proteus_b200.dswx_hls as G
        from proteus_b200 import params as PP
        prev = PP.set_numpy1_promotion(np1)
        try:
            shad_m = G._compute_opera_shadow_layer(ins['dem'], ins['sun_azimuth'], ins['sun_elevation'], -5, 40)
        finally:
            PP.set_numpy1_promotion(prev)
        assert shad_m.dtype == np.bool_ and np.array_equal(shad_m.astype(np.uint8), ref['SHAD_WITH_MARGIN' + sfx])


def test_shadow_shortcuts_never_decide_wrongly_under_numpy1_promotion(pb):
    """The sweep of test_shadow_shortcuts_never_decide_wrongly with the float32 (numpy 1.x) sequence as the exact side:
    the guard bands of the shortcuts also cover the reference's own float32 rounding."""
    import ctypes as C
    from proteus_b200 import _lib
    from proteus_b200.params import sun_terms
    ctx = pb.get_context()
    params = pb.make_params(numpy1_promotion=True)
    for az, el in ((150.0, 45.0), (10.0, 5.0), (359.0, 89.0), (135.0, 70.0)):
        terms = (C.c_double * 5)(*sun_terms(az, el))
        for mode, n in ((0, 2 ** 30), (1, 2 ** 28), (2, 2 ** 28), (3, 2 ** 22)):
            counts = (C.c_uint64 * 8)()
            _lib.check(ctx._lib.pb200_shadow_sweep(ctx.handle, C.byref(params), az, el, terms, mode, 4321 + mode, n, counts))
            n_s, shadow, dec_c, bad_c, dec_s, bad_s, dec_s2, bad_s2 = [int(v) for v in counts]
            assert n_s >= n and bad_c == 0 and bad_s == 0 and bad_s2 == 0, (az, el, mode, bad_c, bad_s, bad_s2)
            if mode == 0:
                assert dec_s / n_s > 0.999 and 0.02 < shadow / n_s < 0.98


def test_device_division_matches_numpy_division():
    """The sweep above trusts __ddiv_rn == numpy true_divide; spot-check that
    link on the host side of the same inputs through the function-level kernel."""
    import proteus_b200.dswx_hls as G
    rng = np.random.default_rng(5)
    bands = [rng.integers(-32768, 32768, (64, 1024), dtype=np.int16) for _ in range(6)]
    bands[1][0, :32] = 0            # green = 0 and swir1 = 0 -> 0/0
    bands[4][0, :32] = 0
    bands[1][1, :32] = 7            # x/0
    bands[4][1, :32] = -7
    with np.errstate(all='ignore'):
        ref = O.compute_diagnostic_tests(*bands, O.default_thresholds())
    got = G._compute_diagnostic_tests(*bands, G.HlsThresholds())
    assert got.dtype == np.uint16
    assert np.array_equal(got, ref)


def test_function_level_dropins_match_oracle(pb):
    import proteus_b200.dswx_hls as G
    rng = np.random.default_rng(9)
    shape = (61, 257)
    u8 = lambda: rng.integers(0, 256, shape, dtype=np.uint8)
    fmask, land = u8(), u8()
    wtr_vals = np.array([0, 1, 2, 3, 4, 254, 255, 9, 100, 252, 253], dtype=np.uint8)
    wtr1 = wtr_vals[rng.integers(0, len(wtr_vals), shape)]
    cloud_vals = np.array(list(range(16)) + [255, 254, 77], dtype=np.uint8)
    cloud = cloud_vals[rng.integers(0, len(cloud_vals), shape)]
    nir = rng.integers(-32768, 32768, shape, dtype=np.int16)
    nir[::2] = rng.integers(900, 1300, nir[::2].shape, dtype=np.int16)
    shad = rng.integers(0, 2, shape).astype(bool)
    th = O.default_thresholds()
    gth = G.HlsThresholds()

    diag = rng.integers(0, 70, shape).astype(np.uint16)
    assert np.array_equal(G.generate_interpreted_layer(diag), O.generate_interpreted_layer(diag))
    big = rng.integers(-5, 200000, shape)
    assert np.array_equal(G.generate_interpreted_layer(big), O.generate_interpreted_layer(big))
    assert np.array_equal(G._get_binary_representation(diag), O.get_binary_representation(diag))
    for mode in ('mask', 'ignore', 'cover'):
        assert np.array_equal(G._compute_preliminary_cloud_layer(fmask, mode),
                              O.compute_preliminary_cloud_layer(fmask, mode))
    with pytest.raises(Exception, match='ERROR mask adjacent to cloud/cloud-shadow mode'):
        G._compute_preliminary_cloud_layer(fmask, 'bogus')

    lists = ([224, 160, 96], [224, 160, 96, 1], [224, 192, 160, 128, 96], [0, 255, 96])
    w_ref, c_ref = wtr1.copy(), cloud.copy()
    assert O.apply_aerosol_class_remapping(w_ref, nir, c_ref, fmask, *lists) is None
    w_got, c_got = wtr1.copy(), cloud.copy()
    assert G._apply_aerosol_class_remapping(w_got, nir, c_got, fmask, *lists) is None
    assert np.array_equal(w_got, w_ref) and np.array_equal(c_got, c_ref)

    for l, s in ((land, shad), (None, shad), (land, None), (None, None)):
        assert np.array_equal(G._apply_landcover_and_shadow_masks(wtr1, nir, l, s, gth),
                              O.apply_landcover_and_shadow_masks(wtr1, nir, l, s, th))
    # nir is not looked at without a land-cover raster (D:1354-1362): None must not be dereferenced
    big = wtr_vals[rng.integers(0, len(wtr_vals), (700, 1000))]
    big_shad = rng.integers(0, 2, big.shape).astype(bool)
    assert np.array_equal(G._apply_landcover_and_shadow_masks(big, None, None, big_shad, gth),
                          O.apply_landcover_and_shadow_masks(big, np.zeros(big.shape, np.int16), None, big_shad, th))
    for mode in ('mask', 'ignore'):
        c_ref = O.add_snow_to_cloud_layer(wtr1, cloud.copy(), fmask, mode)
        c_in = cloud.copy()
        c_got = G._add_snow_to_cloud_layer(wtr1, c_in, fmask, mode)
        assert c_got is c_in and np.array_equal(c_got, c_ref)
    # 'cover': masked dilations (scipy semantics) - clouds restricted to the values the chain produces
    cloud_c = np.array([0, 1, 4, 5, 8, 9, 12, 13], dtype=np.uint8)[rng.integers(0, 8, shape)]
    cloud_c[rng.random(shape) < 0.6] = 0
    fm_c = (fmask & 0x14) | (rng.random(shape) < 0.5).astype(np.uint8) * 4
    c_ref = O.add_snow_to_cloud_layer(wtr1, cloud_c.copy(), fm_c, 'cover')
    c_in = cloud_c.copy()
    c_got = G._add_snow_to_cloud_layer(wtr1, c_in, fm_c, 'cover')
    assert c_got is c_in and np.array_equal(c_got, c_ref)
    assert np.array_equal(G._apply_cloud_masking(wtr1, cloud), O.apply_cloud_masking(wtr1, cloud))
    assert np.array_equal(G._get_binary_water_layer(wtr1), O.get_binary_water_layer(wtr1))
    assert np.array_equal(G._get_confidence_layer(wtr1, cloud), O.get_confidence_layer(wtr1, cloud))
    allv = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(G._collapse_wtr_classes(allv), O.collapse_wtr_classes(allv))
    # empty rasters
    e = np.zeros((0, 7), np.uint8)
    assert G._get_binary_water_layer(e).shape == (0, 7)


def test_shadow_function_matches_oracle_including_borders(pb):
    import proteus_b200.dswx_hls as G
    for seed, shape, sigma in ((1, (90, 140), 400.0), (2, (2, 2), 50.0), (3, (3, 301), 900.0)):
        rng = np.random.default_rng(seed)
        dem = synth._smooth_field(rng, shape[0] + 8, shape[1] + 8, 12.0)[:shape[0], :shape[1]]
        dem = np.ascontiguousarray(dem * np.float32(sigma), dtype=np.float32)
        for az, el in ((150.0, 45.0), (10.0, 5.0), (359.0, 89.0)):
            ref = O.compute_opera_shadow_layer(dem, az, el, -5, 40)
            got = G._compute_opera_shadow_layer(dem, az, el, -5, 40)
            assert got.dtype == np.bool_
            assert np.array_equal(got, ref), (seed, az, el, int((got != ref).sum()))
    # flat DEM, NaN and inf samples: NaN compares false -> "not shadow" 
    dem = np.zeros((40, 40), np.float32)
    dem[10, 10] = np.nan
    dem[20, 20] = np.inf
    dem[30, 5] = 1e30
    with np.errstate(all='ignore'):
        ref = O.compute_opera_shadow_layer(dem, 150.0, 45.0, -5, 40)
    assert np.array_equal(G._compute_opera_shadow_layer(dem, 150.0, 45.0, -5, 40), ref)
    # float64 and integer DEMs: np.gradient promotes an integer DEM to float64, everything then runs in float64
    rng = np.random.default_rng(8)
    d64 = (synth._smooth_field(rng, 120, 150, 10.0) * 300.0 + 500.0).astype(np.float64)
    for az, el in ((150.0, 45.0), (300.0, 80.0)):
        assert np.array_equal(G._compute_opera_shadow_layer(d64, az, el, -5, 40), O.compute_opera_shadow_layer(d64, az, el, -5, 40))
        for dt in (np.int16, np.int32, np.uint16):
            di = d64.astype(dt)
            assert np.array_equal(G._compute_opera_shadow_layer(di, az, el, -5, 40), O.compute_opera_shadow_layer(di, az, el, -5, 40)), dt
    with pytest.raises(NotImplementedError):
        G._compute_opera_shadow_layer(dem.astype(np.float16), 150.0, 45.0, -5, 40)


def test_error_paths(pb):
    t = synth.make_tile(3, 32, 32)
    with pytest.raises(Exception, match='ERROR mask adjacent to cloud/cloud-shadow mode'):
        pb.classify_tile(t['bands'], t['fmask'], mask_adjacent_to_cloud_mode='nope')
    with pytest.raises(NotImplementedError):
        pb.classify_tile([b.astype(np.float32) for b in t['bands']], t['fmask'])
    with pytest.raises(ValueError):
        pb.classify_tile(t['bands'], t['fmask'], t['dem'][:-1], dem_margin=50)
    with pytest.raises(OverflowError):
        pb.classify_tile(t['bands'], t['fmask'], hls_thresholds=pb.HlsThresholds(pswt_1_nir=40000))


def test_full_size_tile_properties(pb):
    """BASELINE config 2 size (3660 x 3660, full product) without running the
    8 s/tile numpy chain: size-independent properties + exact parity on a
    row band that the oracle can afford."""
    t = synth.make_tile(0)
    got = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                           t['sun_azimuth'], t['sun_elevation'], collapse_wtr_classes=False,
                           class_histogram=True)
    h, w = t['fmask'].shape
    # (1) exact parity on rows [1700, 1956) - the DEM window carries its own margin
    r0, r1, m = 1700, 1956, t['dem_margin']
    sub = dict(bands=[b[r0:r1] for b in t['bands']], fmask=t['fmask'][r0:r1],
               dem=t['dem'][r0:r1 + 2 * m], land=t['land'][r0:r1], ocean=t['ocean'][r0:r1])
    ref = O.reference_chain(sub['bands'], sub['fmask'], sub['dem'], sub['land'], sub['ocean'],
                            t['sun_azimuth'], t['sun_elevation'])
    for name in FUSED_LAYERS:
        assert np.array_equal(got[name][r0:r1], ref[name]), name
    # (2) idempotence / functional relations between whole layers
    assert np.array_equal(O.get_binary_water_layer(got['WTR']), got['BWTR'])
    assert np.array_equal(O.apply_cloud_masking(got['WTR2'], got['CLOUD']), got['WTR'])
    assert np.array_equal(O.get_confidence_layer(got['WTR2'], got['CLOUD']), got['CONF'])
    inv = got['DIAG'] == 65535
    assert np.array_equal(inv, got['WTR'] == 255)
    assert np.array_equal(got['CLOUD'] == 255, inv)
    # (3) counters are a checksum of the layers
    cov = got['coverage']
    assert cov['n_not_ocean'] == int(t['ocean'].sum())
    assert cov['n_valid'] == int((~inv & (t['ocean'] != 0)).sum())
    hist = np.bincount(got['WTR'].ravel(), minlength=256)
    assert sum(cov['class_histogram'].values()) == h * w
    assert all(cov['class_histogram'][c] == hist[c] for c in cov['class_histogram'])
    # (4) strip size must not change anything (host pipeline)
    got2 = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                            t['sun_azimuth'], t['sun_elevation'], collapse_wtr_classes=False,
                            strip_rows=96, outputs=('WTR', 'CONF', 'DIAG', 'SHAD'))
    for name in ('WTR', 'CONF', 'DIAG', 'SHAD'):
        assert np.array_equal(got2[name], got[name]), name
    assert np.array_equal(got2['counters'][:3], got['counters'][:3])


def test_mosaic_row_strips_match_whole_raster(pb):
    """BASELINE config 5 in miniature: one raster split into 3 row strips, each
    strip holding only its own DEM rows + one halo row per neighbour, equals
    the reference chain on the whole raster (bit-exact, counters summed)."""
    import torch
    from proteus_b200 import mosaic
    h, w, m = 96 + 64 + 40, 264, 50
    t = synth.make_tile(31, h, w)
    ref = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                            t['sun_azimuth'], t['sun_elevation'])
    world = 3
    bounds = mosaic.strip_bounds(h, world)
    dem_full = torch.from_numpy(t['dem']).cuda()
    got = {k: np.zeros((h, w), ref[k].dtype) for k in ('WTR', 'BWTR', 'CONF', 'DIAG', 'SHAD', 'WTR2', 'CLOUD')}
    counters = np.zeros(3, np.uint64)
    params = pb.make_params(collapse_wtr_classes=False)
    for rank, (r0, r1) in enumerate(bounds):
        d0, d1 = mosaic.dem_rows_for_strip(r0, r1, h, m)
        strip = mosaic.MosaicStrip(
            [torch.from_numpy(b[r0:r1]).cuda() for b in t['bands']],
            torch.from_numpy(t['fmask'][r0:r1]).cuda(), dem_full[d0:d1].clone(),
            torch.from_numpy(t['land'][r0:r1]).cuda(), torch.from_numpy(t['ocean'][r0:r1]).cuda(),
            r0, r1, h, sun_azimuth=t['sun_azimuth'], sun_elevation=t['sun_elevation'],
            params=params, outputs=tuple(got), rank=rank, world=world)
        halo = (dem_full[m + r0 - 1] if rank > 0 else None,
                dem_full[m + r1] if rank < world - 1 else None)
        strip.run(halo=halo)
        res = strip.results()
        for k in got:
            got[k][r0:r1] = res[k]
        counters += res['counters'][:3]
    for k in got:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(counters, ref['counters'])


def test_cover_mode_matches_reference_fixture(pb):
    """mask_adjacent_to_cloud_mode='cover' (dswx_hls.py:2055-2078): the three-step device flow
    against the live-reference fixture, uncollapsed and collapsed."""
    ins, ref = load_golden('cover_mode')
    assert ins['mode'] == 'cover'
    got = _classify_host(pb, ins, collapse=False, class_histogram=True)
    _assert_layers(got, ref, FUSED_LAYERS, 'cover_mode')
    assert np.array_equal(got['counters'][:3], ref['counters'])
    hist = np.bincount(ref['WTR'].ravel(), minlength=256)
    assert [got['coverage']['class_histogram'][c] for c in (0, 1, 2, 3, 4, 252, 253, 254, 255)] == \
        [int(hist[c]) for c in (0, 1, 2, 3, 4, 252, 253, 254, 255)]
    got_c = _classify_host(pb, ins, collapse=True)
    for name, key in (('WTR', 'WTR_COLLAPSED'), ('WTR1', 'WTR1_COLLAPSED'), ('WTR2', 'WTR2_COLLAPSED')):
        assert np.array_equal(got_c[name], ref[key]), key
    for name in ('BWTR', 'CONF', 'DIAG', 'CLOUD'):
        assert np.array_equal(got_c[name], ref[name]), name
    # a second, adversarial tile against the oracle
    t = synth.make_tile(41, 150, 212, adversarial=True)
    o = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                          t['sun_elevation'], processing=dict(mask_adjacent_to_cloud_mode='cover'))
    g = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                         t['sun_elevation'], mask_adjacent_to_cloud_mode='cover', collapse_wtr_classes=False)
    _assert_layers(g, o, FUSED_LAYERS, 'cover adversarial')


def test_cover_mode_dilates_through_band_fill_pixels(pb):
    """D:2084 (cloud[wtr2 == 255] = 255) runs AFTER the dilations of D:2057-2078, which test cloud == 0 on
    the preliminary values.  A pixel that is fill in a band but carries a plain Fmask value (adjacent bit only) is
    part of areas_to_dilate in the reference: the snow front must pass through it."""
    t = synth.make_tile(77, 96, 256, with_dem=False, with_land=False, with_ocean=False)
    for b, v in zip(t['bands'], (300, 500, 400, 3000, 2000, 1000)):      # left half: land (class 0) - no dilation back
        b[...] = v
    for b, v in zip(t['bands'], (400, 500, 300, 200, 100, 50)):          # right half: water (class 1) - D:2070-2076 apply
        b[:, 128:] = v
    t['fmask'][...] = 4                                                  # adjacent to cloud everywhere: the dilation area
    t['fmask'][10, 5] = t['fmask'][10, 200] = 4 | 16                     # one snow seed per half ...
    for b in t['bands']:
        b[10, 8] = b[10, 203] = -9999                                    # ... a band-only fill pixel 3 px further on
        b[12:14, 0:40] = -9999                                           # and a fill band across the front's way
    ref = O.reference_chain(t['bands'], t['fmask'], None, None, None, t['sun_azimuth'], t['sun_elevation'],
                            processing=dict(mask_adjacent_to_cloud_mode='cover'))
    assert ref['WTR1'][10, 5] == 0 and ref['WTR1'][10, 201] == 1
    assert ref['CLOUD'][10, 8] == 255 and ref['CLOUD'][10, 11] == 2      # snow reached the far side of the fill pixel
    assert ref['CLOUD'][15, 5] == 2                                      # and crossed the fill band
    got = pb.classify_tile(t['bands'], t['fmask'], None, None, None, t['sun_azimuth'], t['sun_elevation'],
                           mask_adjacent_to_cloud_mode='cover', collapse_wtr_classes=False)
    _assert_layers(got, ref, FUSED_LAYERS, 'cover through fill')
    # the advisor's search: adversarial tiles (full-range bands: many band-only fill pixels with arbitrary Fmask)
    for seed in range(100, 130):
        t = synth.make_tile(seed, 150, 212, adversarial=True)
        o = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                              t['sun_elevation'], processing=dict(mask_adjacent_to_cloud_mode='cover'))
        g = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                             t['sun_elevation'], mask_adjacent_to_cloud_mode='cover', collapse_wtr_classes=False)
        _assert_layers(g, o, FUSED_LAYERS, f'cover adversarial seed {seed}')


def test_masked_dilation_matches_scipy(pb):
    import ctypes as C
    import torch
    from scipy.ndimage import binary_dilation
    from proteus_b200 import _lib
    ctx = pb.get_context()
    rng = np.random.default_rng(3)
    # one launch runs up to 16 steps in shared memory (96 x 64 tiles + halo): sizes around the tile edges, iteration
    # counts around the per-launch limit, sparse seeds in a dense mask so that fronts cross several tiles
    for shape, iters in (((37, 53), 1), ((64, 200), 10), ((1, 40), 7), ((50, 1), 3), ((129, 131), 4),
                         ((64, 96), 16), ((65, 97), 17), ((200, 300), 40), ((130, 200), 33)):
        a = rng.random(shape) < (0.05 if iters < 16 else 0.002)
        m = rng.random(shape) < (0.7 if iters < 16 else 0.97)
        ref = binary_dilation(a, iterations=iters, mask=m)
        da, dm = torch.from_numpy(a.view(np.uint8)).cuda(), torch.from_numpy(m.view(np.uint8)).cuda()
        out, scr = torch.empty_like(da), torch.empty_like(da)
        _lib.check(ctx._lib.pb200_masked_dilation(
            ctx.handle, da.data_ptr(), dm.data_ptr(), shape[0], shape[1], iters, out.data_ptr(),
            scr.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert np.array_equal(out.cpu().numpy().astype(bool), ref), (shape, iters)


def test_landcover_aggregate_matches_reference_fixture(pb):
    """SURVEY 8f next #1 on the GPU: the numpy tail of create_landcover_mask."""
    import os
    from conftest import GOLDEN_DIR
    import proteus_b200.dswx_hls as G
    z = np.load(os.path.join(GOLDEN_DIR, 'landcover.npz'))
    forest = z['forest_classes'].tolist()
    for i in (0, 1):            # 152 columns: vector path; 93 columns: generic path
        got = G.landcover_aggregate(z[f'wc{i}'], z[f'cop{i}'], forest, int(z[f'year{i}']), str(z[f'type{i}']))
        assert got.dtype == np.uint8 and np.array_equal(got, z[f'land{i}']), i
    rng = np.random.default_rng(12)
    wc = rng.integers(0, 256, (3 * 64, 3 * 200), dtype=np.uint8)
    wc[rng.random(wc.shape) < 0.5] = 80
    cop = rng.integers(0, 256, (64, 200), dtype=np.uint8)
    assert np.array_equal(G.landcover_aggregate(wc, cop, forest, 2030, 'water heavy'),
                          O.landcover_aggregate(wc, cop, forest, 30, 'water heavy'))
    assert np.array_equal(G.landcover_aggregate(wc, cop, None, 2000), O.landcover_aggregate(wc, cop, None, 0))
    with pytest.raises(ValueError):
        G.landcover_aggregate(wc[:-1], cop, forest, 2021)


def test_float32_diagnostic_tests_match_reference_fixture(pb):
    """_compute_diagnostic_tests on float32 bands (--offset-and-scale-inputs): float32 arithmetic in
    numpy's operation order, IEEE division, no FMA contraction -> bit-exact codes."""
    import os
    from conftest import GOLDEN_DIR
    import proteus_b200.dswx_hls as G
    z = np.load(os.path.join(GOLDEN_DIR, 'float_diag.npz'))
    for key in ('scaled', 'dn'):
        got = G._compute_diagnostic_tests(*[z[f'{key}{k}'] for k in range(6)], G.HlsThresholds())
        assert got.dtype == np.uint16 and np.array_equal(got, z[f'diag_{key}']), key
    rng = np.random.default_rng(8)
    bands = [(rng.standard_normal((50, 300)) * 3000).astype(np.float32) for _ in range(6)]
    bands[1][0, :8] = 0.0
    bands[4][0, :8] = 0.0                                  # 0/0 and x/0
    bands[3][1, :8] = np.inf
    with np.errstate(all='ignore'):
        ref = O.compute_diagnostic_tests(*bands, O.default_thresholds())
    assert np.array_equal(G._compute_diagnostic_tests(*bands, G.HlsThresholds()), ref)


def test_browse_relabel_and_scaling(pb):
    """_compute_browse_array (all flag combinations, odd sizes) and the float32
    offset-and-scale of the reflectance bands."""
    import itertools
    import proteus_b200.dswx_hls as G
    rng = np.random.default_rng(21)
    w = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    w[rng.random(w.shape) < 0.7] = 3
    for flags in itertools.product([False, True], repeat=6):
        got = G._compute_browse_array(w, *flags)
        assert got.dtype == np.uint8 and np.array_equal(got, O.compute_browse_array(w, *flags)), flags
    big = rng.integers(0, 5, (512, 1024), dtype=np.uint8)
    assert np.array_equal(G._compute_browse_array(big), O.compute_browse_array(big))
    assert G._compute_browse_array(np.zeros((0, 4), np.uint8)).shape == (0, 4)
    x = np.arange(-32768, 32768, dtype=np.int16).reshape(256, 256)
    inv = rng.random(x.shape) < 0.05
    for scale, offset in ((0.0001, -0.01), (2.75e-5, 123.5), (1.0, 0.0)):
        for invalid in (None, inv, np.where(inv)):
            got = G.scale_and_offset_band(x, scale, offset, invalid)
            ref = O.scale_and_offset_band(x, scale, offset, inv if invalid is not None else None)
            assert got.dtype == np.float32 and np.array_equal(got, ref, equal_nan=True)


def test_batch_of_mixed_tiles_through_the_item_pipeline(pb):
    """The fast kernel's item pipeline (double-buffered DEM tile, full/empty barriers, row requests across items)
    on a batch whose tiles differ in size and in which rasters they carry: tiles with and without DEM alternate,
    some tiles are smaller than one item, some make a CTA change tile after every item.  Both kernel variants
    (graded layers only / all layers), two launches of the same plan."""
    import torch
    specs = [(21, 256, 384, {}), (22, 128, 132, dict(with_dem=False)), (23, 512, 640, dict(adversarial=True)),
             (24, 32, 128, dict(with_land=False, with_ocean=False)), (25, 300, 260, dict(with_dem=False, with_land=False)),
             (26, 64, 1028, {}), (27, 4, 4, {}), (28, 700, 128, dict(adversarial=True))]
    tiles, refs = [], []
    for seed, h, w, kw in specs:
        t = synth.make_tile(seed, h, w, **kw)
        refs.append(O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                                      t['sun_azimuth'], t['sun_elevation']))
        dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in t.items() if k != 'bands'}
        dev['bands'] = [torch.from_numpy(b).cuda() for b in t['bands']]
        tiles.append(dev)
    for layers, names in ((pb.GRADED_LAYERS, ('DIAG', 'WTR', 'BWTR', 'CONF')), (pb.ALL_LAYERS, FUSED_LAYERS)):
        plan = pb.Plan(tiles, pb.make_params(collapse_wtr_classes=False), layers)
        for launch in range(2):
            plan.zero_counters()
            plan.run()
            for i, ref in enumerate(refs):
                res = plan.results(i)
                _assert_layers(res, ref, names, f'tile {i} launch {launch} {len(layers)} layers')
                assert np.array_equal(res['counters'][:3], ref['counters']), (i, launch)


def test_tma_fed_stream_kernel_matches_oracle(pb):
    """The product configuration (four graded layers + counters) on tiles whose planes TMA can address as 4-row
    super-rows runs dswx_fused_stream_dyn_kernel (producer warp + shared-memory ring + row tickets; the static
    dswx_fused_stream_kernel with PB200_STREAM_DYNAMIC=0): tiles smaller than an item, ragged
    right / bottom edges, with and without DEM / LAND / ocean, adversarial data, several tiles per launch (a CTA changes
    tile between items), two launches of the same plan - against the oracle; and the same batch through the
    direct-load kernel (PB200_NO_STREAM) gives the same bits."""
    import os
    import torch
    from proteus_b200 import _lib
    specs = [(51, 96, 128, {}), (52, 100, 132, dict(with_ocean=False)), (53, 4, 36, dict(adversarial=True)),
             (54, 700, 128, dict(adversarial=True)), (55, 256, 384, dict(with_dem=False)),
             (56, 192, 260, dict(with_land=False, with_ocean=False)), (57, 388, 1028, {}), (58, 96, 3660, {}),
             (59, 8, 40, dict(with_dem=False, with_land=False, with_ocean=False))]
    tiles, refs = [], []
    for seed, h, w, kw in specs:
        t = synth.make_tile(seed, h, w, **kw)
        refs.append(O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                                      t['sun_azimuth'], t['sun_elevation']))
        dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in t.items() if k != 'bands'}
        dev['bands'] = [torch.from_numpy(b).cuda() for b in t['bands']]
        tiles.append(dev)
    for params, what in ((pb.make_params(collapse_wtr_classes=False), 'defaults (FAST8)'),
                         (pb.make_params(pb.HlsThresholds(wigt=0.13, pswt_1_ndvi=0.71), collapse_wtr_classes=False,
                                         min_slope_angle=3), 'general parameters')):
        oth = O.default_thresholds()
        oth.wigt, oth.pswt_1_ndvi = 0.13, 0.71
        kw = {} if what.startswith('defaults') else dict(thresholds=oth, processing=dict(min_slope_angle=3))
        if kw:
            refs_p = []
            for seed, h, w, tkw in specs:
                t = synth.make_tile(seed, h, w, **tkw)
                refs_p.append(O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                                                t['sun_azimuth'], t['sun_elevation'], **kw))
        else:
            refs_p = refs
        # both TMA-fed kernels: rows handed to free warps (default), and row w of every chunk to warp w
        for dyn in (True, False):
            os.environ['PB200_STREAM_DYNAMIC'] = '1' if dyn else '0'
            try:
                plan = pb.Plan(tiles, params, pb.GRADED_LAYERS)
            finally:
                del os.environ['PB200_STREAM_DYNAMIC']
            assert plan.kernels & _lib.KERNEL_STREAM, 'the batch should qualify for the TMA-fed kernel'
            assert bool(plan.kernels & _lib.KERNEL_STREAM_DYN) == dyn and ('stream_dyn' in plan.kernel_name) == dyn
            assert bool(plan.kernels & _lib.KERNEL_FAST8) == what.startswith('defaults')
            for launch in range(2):
                plan.zero_counters()
                plan.run()
                for i, ref in enumerate(refs_p):
                    res = plan.results(i)
                    _assert_layers(res, ref, ('DIAG', 'WTR', 'BWTR', 'CONF'), f'{what}: dyn {dyn} tile {i} launch {launch}')
                    assert np.array_equal(res['counters'][:3], ref['counters']), (what, dyn, i, launch)
        os.environ['PB200_NO_STREAM'] = '1'
        try:
            direct = pb.Plan(tiles, params, pb.GRADED_LAYERS)
        finally:
            del os.environ['PB200_NO_STREAM']
        assert not (direct.kernels & _lib.KERNEL_STREAM)
        direct.run()
        for i in range(len(tiles)):
            a, b = plan.results(i), direct.results(i)
            for name in ('DIAG', 'WTR', 'BWTR', 'CONF'):
                assert np.array_equal(a[name], b[name]), (what, i, name)


def test_dynamic_row_kernel_on_many_small_tiles(pb):
    """dswx_fused_stream_dyn_kernel hands rows to whichever warp is free and keeps two tile descriptors in flight: a
    batch of 240 small tiles of random shapes (1 - 6 items each, so that a CTA meets a new tile with nearly every item;
    tiles with and without DEM / LAND / ocean interleaved) against the oracle, every layer and every counter, twice."""
    import torch
    from proteus_b200 import _lib
    rng = np.random.default_rng(77)
    tiles, refs = [], []
    for i in range(240):
        h, w = 4 * int(rng.integers(1, 40)), 4 * int(rng.integers(9, 80))
        kw = dict(with_dem=bool(rng.random() < 0.8), with_land=bool(rng.random() < 0.7), with_ocean=bool(rng.random() < 0.6),
                  adversarial=bool(rng.random() < 0.2))
        t = synth.make_tile(2000 + i, h, w, **kw)
        refs.append(O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'], t['sun_elevation']))
        dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in t.items() if k != 'bands'}
        dev['bands'] = [torch.from_numpy(b).cuda() for b in t['bands']]
        tiles.append(dev)
    plan = pb.Plan(tiles, pb.make_params(collapse_wtr_classes=False), pb.GRADED_LAYERS)
    assert plan.kernels & _lib.KERNEL_STREAM_DYN
    for launch in range(2):
        plan.zero_counters()
        plan.run()
        for i, ref in enumerate(refs):
            res = plan.results(i)
            _assert_layers(res, ref, ('DIAG', 'WTR', 'BWTR', 'CONF'), f'tile {i} launch {launch}')
            assert np.array_equal(res['counters'][:3], ref['counters']), (i, launch)


def test_otsu_threshold_dropin(pb):
    """_compute_otsu_threshold on uint8 hillshades (GPU histogram + compare, host threshold)."""
    import proteus_b200.dswx_hls as G
    from test_oracle_golden import _otsu_images
    imgs = _otsu_images()[:18]
    rng = np.random.default_rng(3)
    imgs.append(np.clip(rng.normal(181, 35, (1237, 2051)), 0, 255).astype(np.uint8))      # odd sizes, > 1 block
    for img in imgs:
        for norm in (True, False):
            got = G._compute_otsu_threshold(img, norm)
            assert got.dtype == np.bool_ and np.array_equal(got, O.compute_otsu_threshold(img, norm))
    _, counts = G._otsu_counts(imgs[-1])
    assert np.array_equal(counts.cpu().numpy(), np.bincount(imgs[-1].ravel(), minlength=256))
    with pytest.raises(NotImplementedError):
        G._compute_otsu_threshold(imgs[0].astype(np.float32))


def test_non_square_pixels_run_the_exact_shadow_sequence(pb):
    """The float32 shadow shortcut assumes square pixels; any other spacing sends every pixel through the exact
    float64 sequence of the fused kernel (the reference itself always uses 30 m x 30 m, dswx_hls.py:5161)."""
    t = synth.make_tile(31, 96, 256)
    m = t['dem_margin']
    for spacing in ((30, 20), (10, 30)):
        params = pb.make_params(pixel_spacing=spacing, collapse_wtr_classes=False)
        got = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'],
                               t['sun_elevation'], params=params, outputs=('SHAD', 'WTR2', 'WTR'))
        shad = O.compute_opera_shadow_layer(t['dem'], t['sun_azimuth'], t['sun_elevation'], -5, 40, *spacing)
        shad = O.crop_2d_array_all_sides(shad, m)
        assert np.array_equal(got['SHAD'].astype(bool), shad), spacing
        ref = O.reference_chain(t['bands'], t['fmask'], None, t['land'], t['ocean'], t['sun_azimuth'], t['sun_elevation'])
        w1 = ref['WTR1_REMAPPED']
        w2 = O.apply_landcover_and_shadow_masks(w1, np.clip(t['bands'][3], 1, None), t['land'], shad, O.default_thresholds())
        assert np.array_equal(got['WTR2'], w2), spacing


def test_full_size_mosaic_strips_equal_the_whole_raster(pb):
    """BASELINE config 5 at full size (21 960 x 21 960 px, DEM 22 060 x 22 060): the raster classified as 8 row strips
    (each with its own DEM rows and one halo row per neighbour, as 8 ranks hold it) equals the raster classified as
    one tile - every graded layer bit for bit, counters summed; plus layer relations the oracle defines."""
    import torch
    from proteus_b200 import mosaic
    size, m, world = 21960, 50, 8
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip('needs 40 GB of free device memory')
    t = synth.make_device_batch(1, size, size, device='cuda', seed=77, n_distinct=1)[0]
    params = pb.make_params(collapse_wtr_classes=True)
    whole = pb.Plan([t], params, pb.GRADED_LAYERS)
    whole.run()
    torch.cuda.synchronize()
    ref = whole.outputs[0]
    ref_counters = whole.counters[0].clone()
    summed = torch.zeros_like(ref_counters)
    bounds = mosaic.strip_bounds(size, world)
    half = 32                                      # rows kept on either side of a seam for the oracle
    kept = {}                                      # row -> {layer: numpy row}, taken from the STRIP outputs
    for rank, (r0, r1) in enumerate(bounds):
        d0, d1 = mosaic.dem_rows_for_strip(r0, r1, size, m)
        strip = mosaic.MosaicStrip([b[r0:r1] for b in t['bands']], t['fmask'][r0:r1], t['dem'][d0:d1].clone(),
                                   t['land'][r0:r1], t['ocean'][r0:r1], r0, r1, size,
                                   sun_azimuth=t['sun_azimuth'], sun_elevation=t['sun_elevation'],
                                   params=params, outputs=pb.GRADED_LAYERS, rank=rank, world=world)
        strip.run(halo=(t['dem'][m + r0 - 1] if rank > 0 else None, t['dem'][m + r1] if rank < world - 1 else None))
        torch.cuda.synchronize()
        for name in pb.GRADED_LAYERS:
            assert torch.equal(strip.outputs[name], ref[name][r0:r1]), (name, rank)
        for a, b in ((0, half), (r1 - r0 - half, r1 - r0)):
            block = {name: strip.outputs[name][a:b].cpu().numpy() for name in pb.GRADED_LAYERS}
            for i in range(b - a):
                kept[r0 + a + i] = {name: block[name][i] for name in pb.GRADED_LAYERS}
        summed += strip.counters.reshape(-1)[:summed.numel()]
        del strip
    # the seams against the ORACLE: 32 rows on either side of each of the 7 strip boundaries, the rows whose DEM
    # stencil (D:4255) reaches into the neighbouring rank's rows - taken from the strips' own outputs
    keymap = {'WTR': 'WTR_COLLAPSED', 'BWTR': 'BWTR', 'CONF': 'CONF', 'DIAG': 'DIAG'}
    for (_, seam) in bounds[:-1]:
        a, b = seam - half, seam + half
        o = O.reference_chain([x[a:b].cpu().numpy() for x in t['bands']], t['fmask'][a:b].cpu().numpy(),
                              t['dem'][a:b + 2 * m].cpu().numpy(), t['land'][a:b].cpu().numpy(),
                              t['ocean'][a:b].cpu().numpy(), t['sun_azimuth'], t['sun_elevation'])
        for name, key in keymap.items():
            got = np.stack([kept[r][name] for r in range(a, b)])
            got = got.view(np.uint16) if name == 'DIAG' else got
            assert np.array_equal(got, o[key]), (name, seam, int((got != o[key]).sum()))
    assert torch.equal(summed[:3], ref_counters[:3])
    wtr, bwtr = ref['WTR'], ref['BWTR']
    assert torch.equal(torch.where((wtr >= 1) & (wtr <= 2), torch.ones_like(wtr), wtr), bwtr)      # D:1727 on collapsed WTR
    assert int(ref_counters[0]) == int(((ref['DIAG'].view(torch.int16) != -1) & (t['ocean'] != 0)).sum())


def test_full_size_l30_minimal_fused_equals_function_chain(pb):
    """BASELINE config 1 at full size (3660 x 3660, no DEM / LAND / ocean, outputs DIAG + WTR): the fused fast kernel
    against the chain of function-granular kernels called in the reference's order (D:5088-5369) - two independent
    code paths over the same 13.4 Mpixel."""
    import torch
    import proteus_b200.dswx_hls as G
    t = synth.make_device_batch(1, 3660, 3660, device='cuda', seed=5, n_distinct=1, full_product=False)[0]
    assert t.get('dem') is None and t.get('land') is None and t.get('ocean') is None
    plan = pb.Plan([t], pb.make_params(collapse_wtr_classes=False), ('DIAG', 'WTR'))
    plan.run()
    got = plan.results(0)
    raw = [b.cpu().numpy() for b in t['bands']]
    fmask = t['fmask'].cpu().numpy()
    invalid = fmask == 255
    for b in raw:
        invalid |= b == -9999                                                       # D:2203-2209
    bands = [np.clip(b, 1, None) for b in raw]                                      # D:2298-2299
    th = G.HlsThresholds()
    diag = G._compute_diagnostic_tests(*bands, th)
    diag[invalid] = 32                                                              # D:5227
    wtr1 = G.generate_interpreted_layer(diag)
    wtr1[invalid] = 255                                                             # D:5249
    cloud = G._compute_preliminary_cloud_layer(fmask, 'mask')
    pr = O.default_processing()
    G._apply_aerosol_class_remapping(
        wtr1, bands[3], cloud, fmask,
        pr['aerosol_not_water_to_high_conf_water_fmask_values'],
        pr['aerosol_water_moderate_conf_to_high_conf_water_fmask_values'],
        pr['aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values'],
        pr['aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values'])
    wtr2 = G._apply_landcover_and_shadow_masks(wtr1, bands[3], None, None, th)
    cloud = G._add_snow_to_cloud_layer(wtr2, cloud, fmask, 'mask')
    wtr = G._apply_cloud_masking(wtr2, cloud)
    assert np.array_equal(got['DIAG'], G._get_binary_representation(diag))
    assert np.array_equal(got['WTR'], wtr)
    cov = got['coverage']
    assert cov['n_valid'] == int((~invalid).sum())
    # ... and the whole tile against the ORACLE (the numpy chain without DEM / LAND / ocean runs in a few seconds)
    ref = O.reference_chain(raw, fmask, None, None, None, 150.0, 45.0)
    assert np.array_equal(got['DIAG'], ref['DIAG'])
    assert np.array_equal(got['WTR'], ref['WTR'])
    assert np.array_equal(got['counters'][:3], ref['counters'])


def test_host_path_reuses_resident_ancillary_rasters(pb):
    """Time series (BASELINE configs[3]): the second acquisition of a tile reuses the DEM / LAND / ocean rasters the
    first call left on the device (reuse_ancillary=True) and still matches the oracle; a call that does not fit the
    resident rasters is refused."""
    from proteus_b200 import _lib
    a = synth.make_tile(41, 200, 264)
    b = synth.make_tile(42, 200, 264)
    b['dem'], b['land'], b['ocean'] = a['dem'], a['land'], a['ocean']            # same MGRS tile, other acquisition
    first = pb.classify_tile(a['bands'], a['fmask'], a['dem'], a['land'], a['ocean'], a['sun_azimuth'],
                             a['sun_elevation'], collapse_wtr_classes=False)
    ref_a = O.reference_chain(a['bands'], a['fmask'], a['dem'], a['land'], a['ocean'], a['sun_azimuth'], a['sun_elevation'])
    _assert_layers(first, ref_a, FUSED_LAYERS, 'first acquisition')
    # poison the host copies: a re-upload would now change the result
    poisoned = dict(dem=np.zeros_like(a['dem']), land=np.full_like(a['land'], 255), ocean=np.ones_like(a['ocean']))
    second = pb.classify_tile(b['bands'], b['fmask'], poisoned['dem'], poisoned['land'], poisoned['ocean'],
                              b['sun_azimuth'], b['sun_elevation'], collapse_wtr_classes=False, reuse_ancillary=True)
    ref_b = O.reference_chain(b['bands'], b['fmask'], a['dem'], a['land'], a['ocean'], b['sun_azimuth'], b['sun_elevation'])
    _assert_layers(second, ref_b, FUSED_LAYERS, 'second acquisition, resident ancillary')
    assert np.array_equal(second['counters'][:3], ref_b['counters'])
    c = synth.make_tile(43, 96, 128)
    with pytest.raises(_lib.Pb200Error):
        pb.classify_tile(c['bands'], c['fmask'], c['dem'], c['land'], c['ocean'], c['sun_azimuth'], c['sun_elevation'],
                         reuse_ancillary=True)
    # a refused call uploads nothing and leaves the resident rasters usable
    third = pb.classify_tile(b['bands'], b['fmask'], poisoned['dem'], poisoned['land'], poisoned['ocean'],
                             b['sun_azimuth'], b['sun_elevation'], collapse_wtr_classes=False, reuse_ancillary=True)
    _assert_layers(third, ref_b, FUSED_LAYERS, 'after a refused call')


def _stand_in_reference_module():
    """A module object shaped like proteus.dswx_hls as far as generate_dswx_layers' per-pixel statements go: the
    constants they use, the one helper that is not replaced, and - for every function install() must rebind - a stub
    that fails when called."""
    import types
    import proteus_b200.dswx_hls as G
    mod = types.ModuleType('proteus_dswx_hls_stand_in')
    for name in ('DIAGNOSTIC_LAYER_NO_DATA_DECIMAL', 'WTR_OCEAN_MASKED', 'UINT8_FILL_VALUE'):
        setattr(mod, name, getattr(G, name))
    mod._crop_2d_array_all_sides = lambda a, margin: a[margin:-margin, margin:-margin]      # D:4320-4337, not replaced

    def make_stub(name):
        def stub(*args, **kwargs):
            raise AssertionError(f'{name}: the CPU function was called - install() did not rebind it')
        stub.__name__ = name
        return stub
    for name in G.REPLACED_FUNCTIONS:
        setattr(mod, name, make_stub(name))
    return mod


@pytest.mark.parametrize('case', ('full_default', 'ignore_noaerosol', 'cover_mode', 'ragged_adversarial'))
def test_install_rebinds_the_reference_call_order_onto_the_gpu(pb, case):
    """INTEGRATION level 2 executed.  install(module) rebinds the 14 per-pixel functions; the statement
    order of generate_dswx_layers (D:5088-5369, oracle/make_golden.py:reference_chain - the very code that produced
    the fixtures from the live reference) then runs on the GPU entry points and must reproduce every fixture layer;
    uninstall() restores the module."""
    import json
    import os
    from conftest import GOLDEN_DIR
    from oracle import make_golden
    import proteus_b200.dswx_hls as G
    mod = _stand_in_reference_module()
    before = {name: getattr(mod, name) for name in G.REPLACED_FUNCTIONS}
    assert pb.install(mod) is mod
    for name in G.REPLACED_FUNCTIONS:
        assert getattr(mod, name) is getattr(G, name), name
    with open(os.path.join(GOLDEN_DIR, 'reference_tables.json')) as f:
        tables = json.load(f)
    ins, ref = load_golden(case)
    t = dict(bands=ins['bands'], fmask=ins['fmask'], dem=ins['dem'], land=ins['land'], ocean=ins['ocean'],
             sun_azimuth=ins['sun_azimuth'], sun_elevation=ins['sun_elevation'], dem_margin=ins['dem_margin'])
    got = make_golden.reference_chain(mod, t, tables['processing'], pb.HlsThresholds(**tables['hls_thresholds']),
                                      ins['mode'], ins['aerosol'])
    for name, r in ref.items():
        assert name in got, name
        g = np.asarray(got[name])
        assert np.array_equal(g.astype(r.dtype) if g.dtype != r.dtype and g.dtype == np.bool_ else g, r), (case, name)
    pb.uninstall(mod)
    for name in G.REPLACED_FUNCTIONS:
        assert getattr(mod, name) is before[name], name


def test_one_shot_classify_and_invalid_and_clip_exports(pb):
    """The two exports nothing else calls: pb200_classify (plan build + launch + release in one asynchronous call on
    the caller's stream) and pb200_invalid_and_clip (D:2203-2209, D:2298-2299)."""
    import ctypes as C
    import torch
    from proteus_b200 import _lib, engine
    ctx = pb.get_context()
    t = synth.make_tile(61, 200, 264)
    ref = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'], t['sun_elevation'])
    dev = {k: torch.from_numpy(t[k]).cuda() for k in ('fmask', 'dem', 'land', 'ocean')}
    bands = [torch.from_numpy(b).cuda() for b in t['bands']]
    outs = {n: torch.empty((200, 264), device='cuda', dtype=torch.int16 if n == 'DIAG' else torch.uint8) for n in FUSED_LAYERS}
    counters = torch.zeros(12, dtype=torch.int64, device='cuda')
    tile = _lib.Tile()
    engine._fill_tile(tile, height=200, width=264, band_ptrs=[b.data_ptr() for b in bands], fmask_ptr=dev['fmask'].data_ptr(),
                      dem_ptr=dev['dem'].data_ptr(), dem_shape=dev['dem'].shape, dem_off=(50, 50),
                      land_ptr=dev['land'].data_ptr(), ocean_ptr=dev['ocean'].data_ptr(),
                      sun=(t['sun_azimuth'], t['sun_elevation']), out_ptrs={k: v.data_ptr() for k, v in outs.items()},
                      counters_ptr=counters.data_ptr())
    params = pb.make_params(collapse_wtr_classes=False)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    for _ in range(3):                                   # back-to-back one-shot calls on a non-default stream
        counters.zero_()
        side.wait_stream(torch.cuda.current_stream())
        _lib.check(ctx._lib.pb200_classify(ctx.handle, C.byref(tile), 1, C.byref(params), C.c_void_p(side.cuda_stream)))
        side.synchronize()
    for n in FUSED_LAYERS:
        a = outs[n].cpu().numpy()
        assert np.array_equal(a.view(np.uint16) if n == 'DIAG' else a, ref[n]), n
    assert np.array_equal(counters.cpu().numpy()[:3].astype(np.uint64), ref['counters'])
    # invalid mask + clip, custom fills, with and without the optional outputs
    n = 200 * 264
    raw = (C.c_void_p * 6)(*[b.data_ptr() for b in bands])
    clipped = [torch.empty_like(b) for b in bands]
    cl = (C.c_void_p * 6)(*[b.data_ptr() for b in clipped])
    invalid = torch.empty((200, 264), dtype=torch.uint8, device='cuda')
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(ctx._lib.pb200_invalid_and_clip(ctx.handle, raw, dev['fmask'].data_ptr(), C.byref(params), n, cl,
                                               invalid.data_ptr(), stream))
    inv_ref, clip_ref = O.invalid_mask_and_clip(t['bands'], t['fmask'])
    assert np.array_equal(invalid.cpu().numpy().astype(bool), inv_ref)
    for got, want in zip(clipped, clip_ref):
        assert np.array_equal(got.cpu().numpy(), want)
    _lib.check(ctx._lib.pb200_invalid_and_clip(ctx.handle, raw, None, C.byref(params), n, (C.c_void_p * 6)(), invalid.data_ptr(), stream))
    inv_nofmask = np.zeros_like(inv_ref)
    for b in t['bands']:
        inv_nofmask |= b == -9999
    assert np.array_equal(invalid.cpu().numpy().astype(bool), inv_nofmask)


def test_hillshade_of_the_otsu_algorithm_matches_the_gdaldem_restatement(pb):
    """SURVEY 8f next #3, second half (PARITY UNPINNED: GDAL's own arithmetic is not available; the target is the
    oracle's restatement of the published gdaldem Horn formula): the fused hillshade + histogram kernel, TMA path and
    generic path, several sun geometries, and the whole 'otsu' shadow layer."""
    import proteus_b200.dswx_hls as G
    for seed, shape in ((1, (300, 412)), (2, (97, 131)), (3, (3, 3)), (4, (40, 1000)), (5, (2, 50)), (6, (3760, 3760))):
        rng = np.random.default_rng(seed)
        dem = synth._smooth_field(rng, shape[0] + 8, shape[1] + 8, 15.0)[:shape[0], :shape[1]]
        dem = np.ascontiguousarray(dem * np.float32(350.0) + np.float32(700.0), dtype=np.float32)
        for az, el in (((150.0, 45.0), (315.0, 45.0), (10.0, 5.0), (200.0, 89.0)) if shape[0] < 1000 else ((150.0, 45.0),)):
            ref = O.compute_hillshade_gdal(dem, az, el)
            got = G.compute_hillshade(dem, az, el)
            assert got.dtype == np.uint8 and np.array_equal(got, ref), (shape, az, el, int((got != ref).sum()))
            hill, counts = G.compute_hillshade(dem, az, el, return_counts=True)
            assert np.array_equal(counts.cpu().numpy(), np.bincount(ref.ravel(), minlength=256))
            if min(shape) >= 3:
                assert np.array_equal(G.compute_otsu_shadow_layer(dem, az, el), O.compute_otsu_shadow_layer(dem, az, el))
    flat = np.full((64, 64), 123.0, np.float32)
    flat[10, 10] = np.nan
    with np.errstate(all='ignore'):
        assert np.array_equal(G.compute_hillshade(flat, 150.0, 45.0), O.compute_hillshade_gdal(flat, 150.0, 45.0))


def test_two_host_tiles_in_flight_give_the_same_layers(pb):
    """TilePipeline (two alternating host pipelines, pb200_classify_host_ex with PB200_HOST_ASYNC + pb200_host_wait):
    a stream of different tiles, some sharing their ancillary rasters with reuse_ancillary, equals the oracle tile by
    tile; results come back in submission order."""
    tiles = [synth.make_tile(70 + i, 200 + 32 * (i % 3), 264) for i in range(7)]
    refs = [O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'], t['sun_elevation'])
            for t in tiles]
    pipe = pb.TilePipeline()
    got = []
    for t in tiles:
        r = pipe.submit(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], t['sun_azimuth'], t['sun_elevation'],
                        collapse_wtr_classes=False)
        if r is not None:
            got.append(r)
    got += pipe.flush()
    assert len(got) == len(tiles)
    for i, (g, ref) in enumerate(zip(got, refs)):
        assert '_pending' not in g
        _assert_layers(g, ref, FUSED_LAYERS, f'tile {i} of the pipeline')
        assert np.array_equal(g['counters'][:3], ref['counters']), i
        assert g['coverage']['n_valid'] == int(ref['counters'][0])
    # time series on both slots: every slot keeps its own resident ancillary rasters
    a = tiles[0]
    series = [synth.make_tile(90 + i, 200, 264) for i in range(5)]
    pipe = pb.TilePipeline()
    out = []
    for i, t in enumerate(series):
        r = pipe.submit(t['bands'], t['fmask'], a['dem'], a['land'], a['ocean'], t['sun_azimuth'], t['sun_elevation'],
                        collapse_wtr_classes=False, reuse_ancillary=i >= 2)
        if r is not None:
            out.append(r)
    out += pipe.flush()
    for t, g in zip(series, out):
        ref = O.reference_chain(t['bands'], t['fmask'], a['dem'], a['land'], a['ocean'], t['sun_azimuth'], t['sun_elevation'])
        _assert_layers(g, ref, FUSED_LAYERS, 'series')
