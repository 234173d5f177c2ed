"""The CPU oracle against the reference's own vectors (no GPU).

Pins oracle/dswx_oracle.py to (1) the reference's known-answer table
(tests/test_dswx_hls_units.py:7-28 of the reference, exported to
tests/golden/reference_tables.json) and (2) the outputs of the LIVE reference
on seeded tiles (tests/golden/*.npz, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_DIR, load_golden
from oracle import dswx_oracle as O

LAYERS = ('DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD', 'WTR', 'BWTR',
          'CONF', 'WTR_COLLAPSED', 'WTR1_COLLAPSED', 'WTR2_COLLAPSED')


def _tables():
    with open(os.path.join(GOLDEN_DIR, 'reference_tables.json')) as f:
        return json.load(f)


def test_interpreted_layer_known_answers():
    """Same construction as the reference's unit test: every key of
    interpreted_dswx_band_dict plus an out-of-table value (111111)."""
    table = {int(k): v for k, v in _tables()['interpreted_dswx_band_dict'].items()}
    assert len(table) == 33
    width = len(table) + 1
    inp = np.full((1, width), 111111)
    exp = np.full((1, width), 255)
    for i, (k, v) in enumerate(table.items()):
        inp[0, i] = k
        exp[0, i] = v
    out = O.generate_interpreted_layer(inp)
    assert out.dtype == np.uint8
    assert np.array_equal(out, exp)


def test_collapse_table_and_defaults_match_reference():
    t = _tables()
    for k, v in t['collapse_wtr_classes_dict'].items():
        assert O.COLLAPSE_LUT[int(k)] == v
    others = [i for i in range(256) if str(i) not in t['collapse_wtr_classes_dict']]
    assert np.all(O.COLLAPSE_LUT[others] == 255)
    th = O.default_thresholds()
    for k, v in t['hls_thresholds'].items():
        assert getattr(th, k) == v
    pr = O.default_processing()
    for k, v in t['processing'].items():
        assert pr[k] == v
    assert O.AEROSOL_REMAPPING_MAX_NIR == t['constants']['AEROSOL_REMAPPING_MAX_NIR'] == 1000.0
    assert O.DEM_MARGIN_IN_PIXELS == t['constants']['DEM_MARGIN_IN_PIXELS']


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_chain_matches_reference_fixture(case):
    ins, ref = load_golden(case)
    got = O.reference_chain(
        ins['bands'], ins['fmask'], ins['dem'], ins['land'], ins['ocean'],
        ins['sun_azimuth'], ins['sun_elevation'],
        processing=dict(mask_adjacent_to_cloud_mode=ins['mode'],
                        apply_aerosol_class_remapping=ins['aerosol']),
        dem_margin=ins['dem_margin'])
    for name in LAYERS + (('SHAD',) if ins['dem'] is not None else ()):
        assert got[name].dtype == ref[name].dtype, name
        assert np.array_equal(got[name], ref[name]), name
    assert np.array_equal(got['counters'], ref['counters'])
    assert np.array_equal(got['percentages'], ref['percentages'])


def test_guard_band_fixture_under_numpy1_promotion():
    """The reference pins numpy 1.23.5 (setup.py:78), where the terrain-shadow test runs in float32 from the dot product
    on.  The guard_band fixture holds the live reference's layers under both promotion rules (oracle/make_golden.py);
    on its DEM - every pixel on a decision boundary - the two differ, and the oracle follows either."""
    ins, ref = load_golden('guard_band')
    assert int((ref['SHAD'] != ref['SHAD_NUMPY1']).sum()) > 500 and int((ref['WTR'] != ref['WTR_NUMPY1']).sum()) > 100
    got = O.reference_chain(ins['bands'], ins['fmask'], ins['dem'], ins['land'], ins['ocean'], ins['sun_azimuth'],
                            ins['sun_elevation'], processing=dict(mask_adjacent_to_cloud_mode=ins['mode'],
                                                                  apply_aerosol_class_remapping=ins['aerosol']),
                            dem_margin=ins['dem_margin'], numpy1_promotion=True)
    for name in ('SHAD', 'WTR2', 'CLOUD', 'WTR', 'BWTR', 'CONF', 'WTR_COLLAPSED'):
        assert np.array_equal(got[name], ref[name + '_NUMPY1']), name
    shad_m = O.compute_opera_shadow_layer(ins['dem'], ins['sun_azimuth'], ins['sun_elevation'], -5, 40, numpy1_promotion=True)
    assert np.array_equal(shad_m.astype(np.uint8), ref['SHAD_WITH_MARGIN_NUMPY1'])
    # a float64 DEM is float64 throughout under both rules
    d64 = ins['dem'].astype(np.float64)
    assert np.array_equal(O.compute_opera_shadow_layer(d64, 150.0, 56.0, -5, 40, numpy1_promotion=True),
                          O.compute_opera_shadow_layer(d64, 150.0, 56.0, -5, 40))


@pytest.mark.parametrize('case', ('full_default', 'full_adversarial', 'shadow_only'))
def test_functions_match_reference_fixture(case):
    """Function-granular: feed each oracle function the reference's
    intermediate layers."""
    ins, ref = load_golden(case)
    th = O.default_thresholds()
    inv, clipped = O.invalid_mask_and_clip(ins['bands'], ins['fmask'])
    assert np.array_equal(inv, ref['INVALID'])
    diag = O.compute_diagnostic_tests(*clipped, th)
    diag[inv] = 32
    assert np.array_equal(diag, ref['DIAG_DECIMAL'])
    assert np.array_equal(O.get_binary_representation(ref['DIAG_DECIMAL']), ref['DIAG'])
    assert np.array_equal(O.compute_preliminary_cloud_layer(ins['fmask'], ins['mode']),
                          ref['PRELIM_CLOUD'])
    if ins['dem'] is not None:
        shad_m = O.compute_opera_shadow_layer(ins['dem'], ins['sun_azimuth'],
                                              ins['sun_elevation'], -5, 40)
        assert shad_m.dtype == np.bool_
        assert np.array_equal(shad_m.astype(np.uint8), ref['SHAD_WITH_MARGIN'])
    shad = ref['SHAD'].astype(bool) if 'SHAD' in ref else None
    w2 = O.apply_landcover_and_shadow_masks(ref['WTR1_REMAPPED'], clipped[3],
                                            ins['land'], shad, th)
    assert np.array_equal(w2, ref['WTR2'])
    cloud = O.add_snow_to_cloud_layer(ref['WTR2'], ref['PRELIM_CLOUD_AFTER_AEROSOL'].copy(),
                                      ins['fmask'], ins['mode'])
    assert np.array_equal(cloud, ref['CLOUD'])
    assert np.array_equal(O.apply_cloud_masking(ref['WTR2'], ref['CLOUD']), ref['WTR'])
    assert np.array_equal(O.get_binary_water_layer(ref['WTR']), ref['BWTR'])
    assert np.array_equal(O.get_confidence_layer(ref['WTR2'], ref['CLOUD']), ref['CONF'])


def test_central_differences_equal_numpy_gradient():
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        f = rng.normal(size=(37, 53)).astype(dt) * 100
        g0, g1 = np.gradient(f)
        assert np.array_equal(O._central_differences(f, 0), g0)
        assert np.array_equal(O._central_differences(f, 1), g1)
        assert O._central_differences(f, 0).dtype == dt


def test_bad_adjacent_mode_raises_like_reference():
    with pytest.raises(Exception, match='ERROR mask adjacent to cloud/cloud-shadow mode'):
        O.compute_preliminary_cloud_layer(np.zeros((2, 2), np.uint8), 'bogus')


def test_coverage_percentages_edge_cases():
    assert O.coverage_percentages(0, 0, 0, 100) == (0, 0, 0)
    assert O.coverage_percentages(99, 98, 100, 100) == (99, 99, 98)
    assert O.coverage_percentages(1, 1, 3, 3) == (33, 33, 100)


def test_landcover_aggregate_matches_reference_fixture():
    """oracle vs the live create_landcover_mask outputs (tests/golden/landcover.npz,
    made with the function's GDAL calls replaced, oracle/ref_import.live_create_landcover_mask)."""
    z = np.load(os.path.join(GOLDEN_DIR, 'landcover.npz'))
    forest = z['forest_classes'].tolist()
    for i in (0, 1):
        got = O.landcover_aggregate(z[f'wc{i}'], z[f'cop{i}'], forest, int(z[f'year{i}']) - 2000, str(z[f'type{i}']))
        assert got.dtype == np.uint8 and np.array_equal(got, z[f'land{i}']), i
    a = np.arange(36, dtype=np.uint8).reshape(6, 6)
    assert np.array_equal(O.decimate_by_summation(a, 3, 3), [[a[:3, :3].sum(), a[:3, 3:].sum() % 256],
                                                             [a[3:, :3].sum() % 256, a[3:, 3:].sum() % 256]])


def test_float32_diagnostic_tests_match_reference_fixture():
    """The --offset-and-scale-inputs flavour of _compute_diagnostic_tests (float32 bands)."""
    z = np.load(os.path.join(GOLDEN_DIR, 'float_diag.npz'))
    th = O.default_thresholds()
    with np.errstate(all='ignore'):
        for key in ('scaled', 'dn'):
            got = O.compute_diagnostic_tests(*[z[f'{key}{k}'] for k in range(6)], th)
            assert np.array_equal(got, z[f'diag_{key}']), key


def test_browse_relabel_and_scaling_restatements():
    """SURVEY 8f next #4 (D:3057-3129, D:3024-3036): all 64 flag combinations on every byte value; the float32
    offset-and-scale expression on the full int16 range.  (Live-reference comparison: test_oracle_vs_reference.py.)"""
    import itertools
    allv = np.arange(256, dtype=np.uint8).reshape(16, 16)
    for flags in itertools.product([False, True], repeat=6):
        b = O.compute_browse_array(allv, *flags)
        collapse, no_agg, nw, cl, sn, oc = flags
        exp = allv.copy()
        if no_agg:
            exp[exp == 4] = 0
        if collapse:
            exp = O.collapse_wtr_classes(exp)
        for on, v in ((nw, 0), (cl, 253), (sn, 252), (oc, 254)):
            if on:
                exp[exp == v] = 255
        assert np.array_equal(b, exp) and b.dtype == np.uint8
    x = np.arange(-32768, 32768, dtype=np.int16)
    out = O.scale_and_offset_band(x, 0.0001, -0.01)
    assert out.dtype == np.float32
    assert np.array_equal(out, np.float32(0.0001) * (x.astype(np.float32) - np.float32(-0.01)))


def _otsu_images():
    rng = np.random.default_rng(0)
    out = []
    for k in range(60):
        kind = k % 6
        if kind == 0:
            img = rng.integers(0, 256, (40, 50), dtype=np.uint8)
        elif kind == 1:
            img = np.clip(rng.normal(180, 30, (64, 64)), 0, 255).astype(np.uint8)
        elif kind == 2:
            img = rng.choice([0, 255], size=(30, 30)).astype(np.uint8)
        elif kind == 3:
            img = np.full((8, 8), rng.integers(0, 256), np.uint8)
        elif kind == 4:
            lo = rng.integers(0, 200)
            img = rng.integers(lo, lo + rng.integers(2, 56), (50, 20), dtype=np.uint8)
        else:
            img = np.clip(np.concatenate([rng.normal(60, 10, 2000), rng.normal(200, 20, 3000)]),
                          0, 255).astype(np.uint8).reshape(50, 100)
        out.append(img)
    return out


def test_otsu_threshold_host_function_equals_the_numpy_restatement():
    """pb200_otsu_threshold (C, from exact per-value counts) reproduces numpy's histogram binning
    and float64 arithmetic to the last bit, NaN cases (constant images) included."""
    from proteus_b200.dswx_hls import otsu_threshold_from_counts
    for img in _otsu_images():
        for norm in (True, False):
            t_c = otsu_threshold_from_counts(np.bincount(img.ravel(), minlength=256), norm)
            t_np = O.otsu_threshold(img, norm)
            assert t_c == t_np, (img.min(), img.max(), norm, t_c, t_np)


def test_hillshade_restatement_known_answers():
    """The gdaldem Horn hillshade restatement (PARITY UNPINNED, see its docstring) on cases with a closed form: a flat
    DEM shades 1 + 254 sin(alt) everywhere but the border; a plane tilted towards the sun is brighter than one tilted
    away; the border is 0 (no data)."""
    flat = np.full((8, 9), 500.0, np.float32)
    h = O.compute_hillshade_gdal(flat, 315.0, 45.0)
    assert h.dtype == np.uint8 and h.shape == flat.shape
    assert (h[1:-1, 1:-1] == int(np.float32(1.0 + 254.0 * np.sin(np.radians(45.0))) + np.float32(0.5))).all()
    assert (h[0] == 0).all() and (h[-1] == 0).all() and (h[:, 0] == 0).all() and (h[:, -1] == 0).all()
    yy, xx = np.mgrid[0:32, 0:32].astype(np.float32)
    east_up = (xx * 10.0).astype(np.float32)            # rises towards the east: faces west
    h_w = O.compute_hillshade_gdal(east_up, 270.0, 30.0)[5, 5]       # sun in the west: lit
    h_e = O.compute_hillshade_gdal(east_up, 90.0, 30.0)[5, 5]        # sun in the east: darker than flat ground
    h_flat = int(np.float32(1.0 + 254.0 * 0.5) + np.float32(0.5))          # flat ground under a sun at 30 degrees
    assert h_w > h_flat > h_e >= 1, (h_w, h_flat, h_e)
    north_up = ((31 - yy) * 10.0).astype(np.float32)    # rises towards the north (row 0): faces south
    assert O.compute_hillshade_gdal(north_up, 180.0, 30.0)[5, 5] > h_flat > O.compute_hillshade_gdal(north_up, 0.0, 30.0)[5, 5]
    assert O.compute_hillshade_gdal(np.zeros((2, 5), np.float32), 150.0, 45.0).sum() == 0
