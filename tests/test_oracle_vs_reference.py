"""The oracle against the LIVE reference imported from /root/reference (build container only; skipped
on the GPU box, where the committed fixtures of tests/golden/ stand in).  Function by function, on
seeds and shapes that differ from the fixtures."""
import numpy as np
import pytest

from oracle import dswx_oracle as O
from oracle import ref_import
from proteus_b200 import synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason='/root/reference is not present')


@pytest.fixture(scope='module')
def ref():
    return ref_import.load()


@pytest.fixture(scope='module')
def ref_thresholds(ref):
    th = ref.HlsThresholds()
    for k, v in ref_import.default_runconfig_groups()['hls_thresholds'].items():
        setattr(th, k, v)
    return th


def test_reference_unit_test_runs_against_the_oracle(ref):
    """The reference's own test (tests/test_dswx_hls_units.py:7-28), with the oracle in place of
    generate_interpreted_layer - and the reference function itself for comparison."""
    width = len(ref.interpreted_dswx_band_dict) + 1
    inp = np.full((1, width), 111111)
    exp = np.full((1, width), 255)
    for i, (k, v) in enumerate(ref.interpreted_dswx_band_dict.items()):
        inp[0, i], exp[0, i] = k, v
    assert np.array_equal(ref.generate_interpreted_layer(inp), exp)
    assert np.array_equal(O.generate_interpreted_layer(inp), exp)


@pytest.mark.parametrize('seed,adversarial', [(101, False), (102, True)])
def test_functions_against_live_reference(ref, ref_thresholds, seed, adversarial):
    t = synth.make_tile(seed, 96, 140, adversarial=adversarial)
    th, oth = ref_thresholds, O.default_thresholds()
    clipped = [np.clip(b, 1, None) for b in t['bands']]
    with np.errstate(all='ignore'):
        d_ref = ref._compute_diagnostic_tests(*clipped, th)
        assert np.array_equal(O.compute_diagnostic_tests(*clipped, oth), d_ref)
        raw = t['bands']                                   # unclipped: zero denominators occur
        assert np.array_equal(O.compute_diagnostic_tests(*raw, oth), ref._compute_diagnostic_tests(*raw, th))
        f32 = [b.astype(np.float32) * np.float32(1e-4) for b in clipped]
        assert np.array_equal(O.compute_diagnostic_tests(*f32, oth), ref._compute_diagnostic_tests(*f32, th))
    assert np.array_equal(O.get_binary_representation(d_ref), ref._get_binary_representation(d_ref))
    for mode in ('mask', 'ignore', 'cover'):
        assert np.array_equal(O.compute_preliminary_cloud_layer(t['fmask'], mode),
                              ref._compute_preliminary_cloud_layer(t['fmask'], mode))
    shad_ref = ref._compute_opera_shadow_layer(t['dem'], 133.0, 27.0, -5, 40)
    assert np.array_equal(O.compute_opera_shadow_layer(t['dem'], 133.0, 27.0, -5, 40), shad_ref)
    rng = np.random.default_rng(seed)
    wtr = np.array([0, 1, 2, 3, 4, 254, 255], np.uint8)[rng.integers(0, 7, t['fmask'].shape)]
    cloud = np.array([0, 1, 4, 5, 8, 9, 12, 13], np.uint8)[rng.integers(0, 8, t['fmask'].shape)]
    lists = ([224, 160, 96], [224, 160, 96], [224, 192, 160, 128, 96], [224, 192, 160, 128, 96])
    w1, c1, w2, c2 = wtr.copy(), cloud.copy(), wtr.copy(), cloud.copy()
    ref._apply_aerosol_class_remapping(w1, clipped[3], c1, t['fmask'], *lists)
    O.apply_aerosol_class_remapping(w2, clipped[3], c2, t['fmask'], *lists)
    assert np.array_equal(w1, w2) and np.array_equal(c1, c2)
    shad = ref._crop_2d_array_all_sides(shad_ref, 50)
    for land, sh in ((t['land'], shad), (None, shad), (t['land'], None)):
        assert np.array_equal(O.apply_landcover_and_shadow_masks(wtr, clipped[3], land, sh, oth),
                              ref._apply_landcover_and_shadow_masks(wtr, clipped[3], land, sh, th))
    for mode in ('mask', 'ignore', 'cover'):
        assert np.array_equal(O.add_snow_to_cloud_layer(wtr, cloud.copy(), t['fmask'], mode),
                              ref._add_snow_to_cloud_layer(wtr, cloud.copy(), t['fmask'], mode))
    full = ref._add_snow_to_cloud_layer(wtr, cloud.copy(), t['fmask'], 'mask')
    assert np.array_equal(O.apply_cloud_masking(wtr, full), ref._apply_cloud_masking(wtr, full))
    assert np.array_equal(O.get_binary_water_layer(wtr), ref._get_binary_water_layer(wtr))
    assert np.array_equal(O.get_confidence_layer(wtr, full), ref._get_confidence_layer(wtr, full))
    allv = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(O.collapse_wtr_classes(allv), ref._collapse_wtr_classes(allv))


def test_landcover_tail_against_live_reference():
    from oracle.make_golden import make_landcover_inputs
    forest = ref_import.default_runconfig_groups()['processing']['forest_mask_landcover_classes']
    wc, cop = make_landcover_inputs(77, 40, 52)
    for year, mt in ((2021, 'standard'), (2099, 'water heavy')):
        live = ref_import.live_create_landcover_mask(wc, cop, forest, year, mt)
        assert np.array_equal(O.landcover_aggregate(wc, cop, forest, year - 2000, mt), live)
    assert np.array_equal(O.landcover_aggregate(wc, cop, None, 21),
                          ref_import.live_create_landcover_mask(wc, cop, None, 2021))


def test_browse_and_rgb_scaling_against_live_reference(ref):
    import itertools
    rng = np.random.default_rng(5)
    w = np.array([0, 1, 2, 3, 4, 252, 253, 254, 255, 7, 100], np.uint8)[rng.integers(0, 11, (48, 64))]
    for flags in itertools.product([False, True], repeat=6):
        assert np.array_equal(O.compute_browse_array(w, *flags), ref._compute_browse_array(w, *flags)), flags
    bands = [rng.integers(-2000, 12000, (40, 50)).astype(np.int16) for _ in range(3)]
    inv = rng.random((40, 50)) < 0.1
    off = dict(red=-0.01, green=0.0, blue=123.5, swir1=0.25, nir=-3.0)
    sc = dict(red=1e-4, green=0.0001, blue=2.75e-5, swir1=1e-4, nir=0.5)
    for infrared in (False, True):
        out = ref_import.live_save_output_rgb(bands[0], bands[1], bands[2], off, sc, np.where(inv), infrared)
        keys = ('swir1', 'nir', 'red') if infrared else ('red', 'green', 'blue')
        for a, band, kk in zip(out, bands, keys):
            assert np.array_equal(a, O.scale_and_offset_band(band, sc[kk], off[kk], inv), equal_nan=True), kk


def test_otsu_threshold_against_live_reference(ref):
    from test_oracle_golden import _otsu_images
    from proteus_b200.dswx_hls import otsu_threshold_from_counts
    for img in _otsu_images():
        for norm in (True, False):
            with np.errstate(all='ignore'):
                live = ref._compute_otsu_threshold(img, norm)
            assert np.array_equal(live, O.compute_otsu_threshold(img, norm))
            assert np.array_equal(live, img > otsu_threshold_from_counts(np.bincount(img.ravel(), minlength=256), norm))


@pytest.mark.parametrize('dtype', [np.float64, np.int16, np.int32, np.uint16])
def test_shadow_layer_on_float64_and_integer_dems_against_live_reference(ref, dtype):
    """np.gradient promotes an integer DEM to float64 (D:4255); the oracle follows the live reference there too."""
    from proteus_b200 import synth
    rng = np.random.default_rng(8)
    dem = (synth._smooth_field(rng, 90, 110, 9.0) * 300.0 + 500.0).astype(dtype)
    for az, el in ((150.0, 45.0), (300.0, 80.0), (20.0, 8.0)):
        assert np.array_equal(O.compute_opera_shadow_layer(dem, az, el, -5, 40),
                              ref._compute_opera_shadow_layer(dem, az, el, -5, 40))


@pytest.mark.parametrize('az,el,seed', [(150.0, 56.0, 1), (150.0, 45.0, 2), (10.0, 70.0, 3), (220.0, 58.5, 4)])
def test_shadow_layer_under_numpy1_promotion_against_the_live_code_object(ref, az, el, seed):
    """The reference pins numpy 1.23.5 (setup.py:78).  ref_import.live_shadow_layer_numpy1 runs the UNMODIFIED code
    object of _compute_opera_shadow_layer with numpy-1 scalar casting (float32_array * float64_scalar stays float32);
    the oracle's numpy1_promotion flag follows it - on DEMs built to sit on the decision boundaries, where the two
    promotion rules decide differently, and on a smooth one, where they agree."""
    from oracle import ref_import
    from proteus_b200 import synth
    dem = synth.make_guard_band_dem(96, 120, az, el, seed=seed)
    live1 = ref_import.live_shadow_layer_numpy1(dem, az, el, -5, 40)
    live2 = ref._compute_opera_shadow_layer(dem, az, el, -5, 40)
    assert live1.dtype == np.bool_
    assert np.array_equal(O.compute_opera_shadow_layer(dem, az, el, -5, 40, numpy1_promotion=True), live1)
    assert np.array_equal(O.compute_opera_shadow_layer(dem, az, el, -5, 40), live2)
    if el == 56.0:
        assert int((live1 != live2).sum()) > 100
    rng = np.random.default_rng(seed)
    smooth = (synth._smooth_field(rng, 90, 110, 9.0) * 300.0 + 500.0).astype(np.float32)
    smooth[3, 4], smooth[10, 10], smooth[20, 5] = np.nan, np.inf, -np.inf
    assert np.array_equal(O.compute_opera_shadow_layer(smooth, az, el, -5, 40, numpy1_promotion=True),
                          ref_import.live_shadow_layer_numpy1(smooth, az, el, -5, 40))
