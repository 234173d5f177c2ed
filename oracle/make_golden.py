"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz from the LIVE reference.

Runs the unmodified functions of /root/reference/src/proteus/dswx_hls.py
(imported through oracle/ref_import.py) in the order generate_dswx_layers
calls them (D:5088-5369) on small seeded synthetic tiles, and stores inputs
and every output layer.  The fixtures travel to the GPU box, the reference
does not.

    python oracle/make_golden.py            # build container only

numpy version used is recorded in each file (the reference pins 1.23.5; the
container has 2.x - see SURVEY.md section 8c for what that changes: nothing on
the integer layers).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from proteus_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# name -> (tile_id, height, width, make_tile kwargs, mode, aerosol, sun)
CASES = {
    'full_default':     (0, 256, 320, {}, 'mask', True, None),
    'full_adversarial': (1, 192, 256, dict(adversarial=True), 'mask', True, (135.0, 30.0)),
    'ignore_noaerosol': (2, 160, 200, {}, 'ignore', False, (170.0, 62.0)),
    'l30_minimal':      (4, 200, 264, dict(with_dem=False, with_land=False, with_ocean=False), 'mask', True, None),
    'ragged_adversarial': (7, 97, 131, dict(adversarial=True, with_ocean=False), 'mask', True, (121.0, 21.0)),
    'shadow_only':      (8, 130, 150, dict(with_land=False, with_ocean=False), 'ignore', True, (200.0, 15.0)),
    'cover_mode':       (5, 160, 192, {}, 'cover', True, None),
    # DEM on the two decision boundaries of the terrain-shadow test (synth.make_guard_band_dem): about half of the
    # pixels inside the guard bands of the float32 shortcut; also stored under numpy 1.x promotion (out_*_NUMPY1)
    'guard_band':       (11, 128, 160, dict(guard_band_dem=True), 'mask', True, (150.0, 56.0)),
}


def reference_chain(ref, t, processing, thresholds, mode, aerosol, shadow_fn=None):
    """generate_dswx_layers' per-pixel statements, calling the reference.  ``shadow_fn`` replaces
    ref._compute_opera_shadow_layer (ref_import.live_shadow_layer_numpy1: the same code object under numpy 1.x
    promotion rules)."""
    shadow_fn = shadow_fn or ref._compute_opera_shadow_layer
    fmask = t['fmask']
    invalid = fmask == 255                                     # D:2204 (Fmask first in v2? order is irrelevant: OR)
    clipped = []
    for raw in t['bands']:
        invalid = np.logical_or(invalid, raw == -9999)         # D:2206-2207
        clipped.append(np.clip(raw, 1, None))                  # D:2299
    blue, green, red, nir, swir1, swir2 = clipped
    invalid_ind = np.where(invalid)                            # D:5040
    valid = ~invalid                                           # D:5041
    prelim = ref._compute_preliminary_cloud_layer(fmask, mode)  # D:5089
    total = fmask.size
    ocean = t['ocean']
    if ocean is not None:
        valid = np.logical_and(valid, ocean)                   # D:5104
        n_not_ocean = np.sum(ocean)                            # D:5105
    else:
        n_not_ocean = total
    n_valid = np.sum(valid)                                    # D:5110
    n_cloud_and_valid = np.sum((prelim != 0) & valid)          # D:5111
    spatial = int(100 * float(n_valid) / total)                # D:5115
    cloud_cov = 0 if n_valid == 0 else int(100 * float(n_cloud_and_valid) / n_valid)
    spatial_no = 0 if n_not_ocean == 0 else int(100 * float(n_valid) / n_not_ocean)
    shad = None
    if t['dem'] is not None:
        shad_m = shadow_fn(                                    # D:5161
            t['dem'], t['sun_azimuth'], t['sun_elevation'],
            processing['min_slope_angle'], processing['max_sun_local_inc_angle'])
        shad = ref._crop_2d_array_all_sides(shad_m, t['dem_margin'])   # D:5166
    diag_dec = ref._compute_diagnostic_tests(blue, green, red, nir, swir1, swir2, thresholds)
    diag_dec[invalid_ind] = ref.DIAGNOSTIC_LAYER_NO_DATA_DECIMAL       # D:5227
    diag_dec_saved = diag_dec.copy()
    wtr1 = ref.generate_interpreted_layer(diag_dec)            # D:5229
    diag = ref._get_binary_representation(diag_dec)            # D:5231
    if ocean is not None:
        wtr1[ocean == 0] = ref.WTR_OCEAN_MASKED                # D:5245
    wtr1[invalid_ind] = ref.UINT8_FILL_VALUE                   # D:5249
    wtr1_saved = wtr1.copy()
    if aerosol:                                                # D:5260-5266
        ref._apply_aerosol_class_remapping(
            wtr1, nir, prelim, fmask,
            processing['aerosol_not_water_to_high_conf_water_fmask_values'],
            processing['aerosol_water_moderate_conf_to_high_conf_water_fmask_values'],
            processing['aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values'],
            processing['aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values'])
    prelim_after_aerosol = prelim.copy()
    wtr2 = ref._apply_landcover_and_shadow_masks(wtr1, nir, t['land'], shad, thresholds)
    cloud = ref._add_snow_to_cloud_layer(wtr2, prelim, fmask, mode)    # D:5282
    wtr = ref._apply_cloud_masking(wtr2, cloud)                # D:5286
    bwtr = ref._get_binary_water_layer(wtr)                    # D:5358
    conf = ref._get_confidence_layer(wtr2, cloud)              # D:5368
    out = dict(
        DIAG_DECIMAL=diag_dec_saved, DIAG=diag, WTR1=wtr1_saved, WTR1_REMAPPED=wtr1,
        PRELIM_CLOUD=ref._compute_preliminary_cloud_layer(fmask, mode),
        PRELIM_CLOUD_AFTER_AEROSOL=prelim_after_aerosol,
        WTR2=wtr2, CLOUD=cloud, WTR=wtr, BWTR=bwtr, CONF=conf,
        WTR_COLLAPSED=ref._collapse_wtr_classes(wtr),          # D:2688-2689
        WTR1_COLLAPSED=ref._collapse_wtr_classes(wtr1_saved),
        WTR2_COLLAPSED=ref._collapse_wtr_classes(wtr2),
        INVALID=invalid,
        counters=np.array([n_valid, n_cloud_and_valid, n_not_ocean], dtype=np.uint64),
        percentages=np.array([spatial, spatial_no, cloud_cov], dtype=np.int64))
    if shad is not None:
        out['SHAD'] = shad.astype(np.uint8)
        out['SHAD_WITH_MARGIN'] = shad_m.astype(np.uint8)
    return out


def main():
    ref = ref_import.load()
    groups = ref_import.default_runconfig_groups()
    thresholds = ref.HlsThresholds()
    for k, v in groups['hls_thresholds'].items():
        setattr(thresholds, k, v)
    processing = groups['processing']
    os.makedirs(GOLDEN_DIR, exist_ok=True)

    # the reference's own known-answer table (tests/test_dswx_hls_units.py:7-28)
    table = {str(k): int(v) for k, v in ref.interpreted_dswx_band_dict.items()}
    collapse = {str(k): int(v) for k, v in ref.collapse_wtr_classes_dict.items()}
    with open(os.path.join(GOLDEN_DIR, 'reference_tables.json'), 'w') as f:
        json.dump(dict(
            interpreted_dswx_band_dict=table, collapse_wtr_classes_dict=collapse,
            hls_thresholds={k: v for k, v in groups['hls_thresholds'].items()},
            processing={k: processing[k] for k in (
                'apply_aerosol_class_remapping',
                'aerosol_not_water_to_high_conf_water_fmask_values',
                'aerosol_water_moderate_conf_to_high_conf_water_fmask_values',
                'aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values',
                'aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values',
                'min_slope_angle', 'max_sun_local_inc_angle', 'mask_adjacent_to_cloud_mode')},
            constants=dict(AEROSOL_REMAPPING_MAX_NIR=ref.AEROSOL_REMAPPING_MAX_NIR,
                           DEM_MARGIN_IN_PIXELS=ref.DEM_MARGIN_IN_PIXELS,
                           SOFTWARE_VERSION=ref.SOFTWARE_VERSION),
            numpy_version=np.__version__), f, indent=1, sort_keys=True)

    for name, (tid, h, w, kw, mode, aerosol, sun) in CASES.items():
        kw = dict(kw)
        guard_band = kw.pop('guard_band_dem', False)
        t = synth.make_tile(tid, h, w, sun=sun, **kw)
        if guard_band:
            t['dem'] = synth.make_guard_band_dem(h, w, *sun, min_slope_angle=processing['min_slope_angle'],
                                                 max_sun_local_inc_angle=processing['max_sun_local_inc_angle'])
        out = reference_chain(ref, t, processing, thresholds, mode, aerosol)
        if guard_band:
            np1 = reference_chain(ref, t, processing, thresholds, mode, aerosol,
                                  shadow_fn=ref_import.live_shadow_layer_numpy1)
            for key in ('SHAD', 'SHAD_WITH_MARGIN', 'WTR2', 'CLOUD', 'WTR', 'BWTR', 'CONF', 'WTR_COLLAPSED'):
                out[key + '_NUMPY1'] = np1[key]
            print(f'{name}: numpy 1.x promotion changes SHAD on {int((np1["SHAD"] != out["SHAD"]).sum())} px, '
                  f'WTR on {int((np1["WTR"] != out["WTR"]).sum())} px')
        arrays = {f'in_band{k}': b for k, b in enumerate(t['bands'])}
        arrays['in_fmask'] = t['fmask']
        for key in ('dem', 'land', 'ocean'):
            if t[key] is not None:
                arrays[f'in_{key}'] = t[key]
        arrays['in_sun'] = np.array([t['sun_azimuth'], t['sun_elevation']])
        arrays['in_mode'] = np.array(mode)
        arrays['in_aerosol'] = np.array(aerosol)
        arrays['in_dem_margin'] = np.array(t['dem_margin'])
        arrays['numpy_version'] = np.array(np.__version__)
        for k, v in out.items():
            arrays[f'out_{k}'] = v
        path = os.path.join(GOLDEN_DIR, f'{name}.npz')
        np.savez_compressed(path, **arrays)
        print(f'{name}: {h}x{w} mode={mode} aerosol={aerosol} -> '
              f'{os.path.getsize(path) / 1024:.0f} KiB; '
              f'WTR classes {np.unique(out["WTR"]).tolist()}; '
              f'counters {out["counters"].tolist()}')


def make_landcover_inputs(seed, h, w):
    rng = np.random.default_rng(seed)
    vals = np.array([10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100, 0], np.uint8)
    blob = np.kron(rng.integers(0, len(vals), (-(-h // 5), -(-w // 5))), np.ones((15, 15), int))[:3 * h, :3 * w]
    noise = vals[rng.integers(0, len(vals), (3 * h, 3 * w))]
    wc = np.where(rng.random((3 * h, 3 * w)) < 0.6, vals[blob], noise).astype(np.uint8)
    cop = np.array([20, 50, 111, 113, 115, 116, 121, 123, 125, 126, 30, 40, 200, 0],
                   np.uint8)[rng.integers(0, 14, (h, w))]
    return wc, cop


def extra_fixtures():
    """LAND aggregation (live create_landcover_mask with its GDAL calls replaced) and the
    float32 'scaled' diagnostic tests (live _compute_diagnostic_tests on float32 bands)."""
    ref = ref_import.load()
    groups = ref_import.default_runconfig_groups()
    forest = groups['processing']['forest_mask_landcover_classes']
    arrays = {'forest_classes': np.array(forest)}
    for i, (seed, h, w, year, mt) in enumerate(((5, 120, 152, 2021, 'standard'), (6, 67, 93, 2020, 'water heavy'))):
        wc, cop = make_landcover_inputs(seed, h, w)
        arrays[f'wc{i}'], arrays[f'cop{i}'] = wc, cop
        arrays[f'year{i}'], arrays[f'type{i}'] = np.array(year), np.array(mt)
        arrays[f'land{i}'] = ref_import.live_create_landcover_mask(wc, cop, forest, year, mt)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'landcover.npz'), **arrays)
    print('landcover:', {k: np.unique(v).tolist() for k, v in arrays.items() if k.startswith('land')})

    thresholds = ref.HlsThresholds()
    for k, v in groups['hls_thresholds'].items():
        setattr(thresholds, k, v)
    t = synth.make_tile(9, 96, 160)
    # D:2298-2302: clip, then scale_factor * (float32(image) - offset) with the HLS scale 1e-4
    scaled = [0.0001 * (np.asarray(np.clip(b, 1, None), dtype=np.float32) - 0.0) for b in t['bands']]
    scaled = [np.asarray(b, dtype=np.float32) for b in scaled]
    # thresholds sit in unscaled units (D:44): also exercise float bands in DN units
    dn = [np.asarray(np.clip(b, 1, None), dtype=np.float32) for b in t['bands']]
    out = {f'scaled{k}': b for k, b in enumerate(scaled)}
    out.update({f'dn{k}': b for k, b in enumerate(dn)})
    with np.errstate(all='ignore'):
        out['diag_scaled'] = ref._compute_diagnostic_tests(*scaled, thresholds)
        out['diag_dn'] = ref._compute_diagnostic_tests(*dn, thresholds)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'float_diag.npz'), **out)
    print('float_diag: codes', np.unique(out['diag_scaled']).tolist(), np.unique(out['diag_dn']).size)


if __name__ == '__main__':
    if '--extra-only' not in sys.argv:
        main()
    extra_fixtures()
