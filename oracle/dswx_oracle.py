"""TEST INFRASTRUCTURE ONLY - CPU (numpy) restatement of the per-pixel
DSWx-HLS classification path of nasa/PROTEUS v1.0.2.

This file is the *checker* for the CUDA path, never the thing shipped or
measured: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``proteus_b200/`` imports it.

Every function restates one reference function; ``D:`` abbreviates
``/root/reference/src/proteus/dswx_hls.py``.  The restatement keeps the
reference's numpy *numerics* (int16 wrap-around sums, int16/int16 true-divide
in float64, float32 gradient + float64 dot product under numpy >= 2) but is
written independently (table look-ups instead of per-class passes, explicit
finite differences instead of ``np.gradient``).

Pinning (see DESIGN.md "Oracle"):
  * ``generate_interpreted_layer`` against the reference's own known-answer
    test ``tests/test_dswx_hls_units.py:7-28`` (the 33-entry table);
  * every function, and the whole chain, against the *live* reference imported
    from /root/reference (``tests/test_oracle_vs_reference.py``, build container
    only) and against the committed outputs of that reference in
    ``tests/golden/*.npz`` (made by ``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# constants (values restated from D:46-58, D:94-95, D:146-185, D:252-264)
# --------------------------------------------------------------------------
UINT8_FILL_VALUE = 255                      # D:50
DEM_MARGIN_IN_PIXELS = 50                   # D:58
AEROSOL_REMAPPING_MAX_NIR = 0.1 / 0.0001    # D:45-46 (== 1000.0)
DIAG_FILL_DECIMAL = 0b100000                # D:94
DIAG_FILL_BINARY_REPR = 65535               # D:95
WTR_SNOW_MASKED, WTR_CLOUD_MASKED, WTR_OCEAN_MASKED = 252, 253, 254  # D:162-164
LAND_WATER, LAND_EVERGREEN = 200, 201       # D:260-261

# D:97-143 - diagnostic code (5 test bits) -> WTR-1 class.  Written here as
# "class: codes" and inverted below.
_CLASS_TO_CODES = {
    0: (0b00000, 0b00001, 0b00010, 0b00100, 0b01000),
    1: (0b01111, 0b10111, 0b11011, 0b11101, 0b11110, 0b11111),
    2: (0b00111, 0b01011, 0b01101, 0b01110, 0b10011, 0b10101, 0b10110,
        0b11001, 0b11010, 0b11100),
    3: (0b11000,),
    4: (0b00011, 0b00101, 0b00110, 0b01001, 0b01010, 0b01100, 0b10000,
        0b10001, 0b10010, 0b10100),
}
DIAG_TO_WTR1 = np.full(64, UINT8_FILL_VALUE, dtype=np.uint8)
for _cls, _codes in _CLASS_TO_CODES.items():
    for _c in _codes:
        DIAG_TO_WTR1[_c] = _cls
DIAG_TO_WTR1[DIAG_FILL_DECIMAL] = UINT8_FILL_VALUE
assert sorted(c for cs in _CLASS_TO_CODES.values() for c in cs) == list(range(32))

# D:201-213 - uncollapsed -> collapsed classes; anything else -> 255
COLLAPSE_LUT = np.full(256, UINT8_FILL_VALUE, dtype=np.uint8)
COLLAPSE_LUT[[0, 1, 2, 3, 4]] = [0, 1, 1, 2, 2]
COLLAPSE_LUT[[252, 253, 254, 255]] = [252, 253, 254, 255]

BAND_NAMES = ('blue', 'green', 'red', 'nir', 'swir1', 'swir2')


class HlsThresholds:
    """Same attribute names as the reference's holder (D:274-318)."""
    _FIELDS = ('wigt', 'awgt', 'pswt_1_mndwi', 'pswt_1_nir', 'pswt_1_swir1',
               'pswt_1_ndvi', 'pswt_2_mndwi', 'pswt_2_blue', 'pswt_2_nir',
               'pswt_2_swir1', 'pswt_2_swir2', 'lcmask_nir')

    def __init__(self, **kw):
        for f in self._FIELDS:
            setattr(self, f, kw.get(f))


def default_thresholds():
    """Values of /root/reference/src/proteus/defaults/dswx_hls.yaml:176-212."""
    return HlsThresholds(
        wigt=0.124, awgt=0.0, pswt_1_mndwi=-0.44, pswt_1_nir=1500,
        pswt_1_swir1=900, pswt_1_ndvi=0.7, pswt_2_mndwi=-0.5,
        pswt_2_blue=1000, pswt_2_nir=2500, pswt_2_swir1=3000,
        pswt_2_swir2=1000, lcmask_nir=1200)


def default_processing():
    """Values of defaults/dswx_hls.yaml:64-109 that reach the hot path."""
    return dict(
        apply_aerosol_class_remapping=True,
        aerosol_not_water_to_high_conf_water_fmask_values=[224, 160, 96],
        aerosol_water_moderate_conf_to_high_conf_water_fmask_values=[224, 160, 96],
        aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values=[224, 192, 160, 128, 96],
        aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values=[224, 192, 160, 128, 96],
        min_slope_angle=-5, max_sun_local_inc_angle=40,
        mask_adjacent_to_cloud_mode='mask')


# --------------------------------------------------------------------------
# a1 / a2: load-stage statements  (D:2203-2209, D:2298-2299)
# --------------------------------------------------------------------------
def invalid_mask_and_clip(raw_bands, fmask, band_fill=-9999, fmask_fill=255):
    """invalid = OR over the 7 rasters of (raw == fill) evaluated on RAW
    values (D:2203-2209); then each reflectance band := max(band, 1), dtype
    kept (D:2298-2299).  Fmask is not clipped (D:2224-2226).

    ``band_fill`` may be a scalar or a sequence of 6 per-band fills."""
    fills = (list(band_fill) if np.ndim(band_fill) else [band_fill] * 6)
    invalid = (fmask == fmask_fill)
    clipped = []
    for raw, fill in zip(raw_bands, fills):
        invalid = np.logical_or(invalid, raw == fill)
        clipped.append(np.clip(raw, 1, None))
    return invalid, clipped


# --------------------------------------------------------------------------
# a6: diagnostic tests  (D:1840-1916)
# --------------------------------------------------------------------------
def compute_diagnostic_tests(blue, green, red, nir, swir1, swir2, th):
    """5 spectral tests -> uint16 code 0..31.  With int16 inputs the sums and
    differences wrap in int16 and ``/`` is int16/int16 -> float64 (D:1872,
    D:1884); the thresholds are Python scalars."""
    with np.errstate(divide='ignore', invalid='ignore'):
        mndwi = (green - swir1) / (green + swir1)            # D:1872
        ndvi = (nir - red) / (nir + red)                     # D:1884
    mbsrv = green + red                                      # D:1875
    mbsrn = nir + swir1                                      # D:1878
    awesh = blue + 2.5 * green - 1.5 * mbsrn - 0.25 * swir2  # D:1881
    t1 = mndwi > th.wigt                                     # D:1893
    t2 = mbsrv > mbsrn                                       # D:1896
    t3 = awesh > th.awgt                                     # D:1899
    t4 = ((mndwi > th.pswt_1_mndwi) & (swir1 < th.pswt_1_swir1) &
          (nir < th.pswt_1_nir) & (ndvi < th.pswt_1_ndvi))   # D:1902-1906
    t5 = ((mndwi > th.pswt_2_mndwi) & (blue < th.pswt_2_blue) &
          (swir1 < th.pswt_2_swir1) & (swir2 < th.pswt_2_swir2) &
          (nir < th.pswt_2_nir))                             # D:1909-1914
    diag = (t1.astype(np.uint16) | (t2.astype(np.uint16) << 1) |
            (t3.astype(np.uint16) << 2) | (t4.astype(np.uint16) << 3) |
            (t5.astype(np.uint16) << 4))
    return diag


# --------------------------------------------------------------------------
# a8: DIAG -> WTR-1  (D:1687-1707, table D:97-143)
# --------------------------------------------------------------------------
def generate_interpreted_layer(diag):
    """Every value not in the 33-entry table maps to 255 (D:1702)."""
    d = np.asarray(diag)
    in_table = (d >= 0) & (d < 64)
    idx = np.where(in_table, d, 63).astype(np.intp)   # 63 is not in the table
    return DIAG_TO_WTR1[idx]


# --------------------------------------------------------------------------
# a9: DIAG decimal -> "binary digits read as decimal"  (D:4286-4317)
# --------------------------------------------------------------------------
def get_binary_representation(diag, nbits=6):
    d = np.asarray(diag)
    out = np.zeros(d.shape, dtype=np.uint16)
    for i in range(min(nbits, 5)):
        out += (((d >> i) & 1) * (10 ** i)).astype(np.uint16)
    for i in range(5, nbits):                      # D:4312-4315
        out[((d >> i) & 1) != 0] = DIAG_FILL_BINARY_REPR
    return out


# --------------------------------------------------------------------------
# a3: Fmask -> preliminary CLOUD  (D:1919-1993)
# --------------------------------------------------------------------------
def compute_preliminary_cloud_layer(fmask, mask_adjacent_to_cloud_mode):
    if mask_adjacent_to_cloud_mode not in ('mask', 'ignore', 'cover'):
        raise Exception('ERROR mask adjacent to cloud/cloud-shadow mode:'
                        f' {mask_adjacent_to_cloud_mode}')    # D:1977-1981
    shadow_like = (fmask & 8) != 0                             # D:1984
    if mask_adjacent_to_cloud_mode == 'mask':
        shadow_like = shadow_like | ((fmask & 4) != 0)         # D:1986-1988
    cloud = shadow_like.astype(np.uint8)
    cloud += (((fmask & 2) != 0) * 4).astype(np.uint8)         # D:1991
    return cloud


# --------------------------------------------------------------------------
# a4: coverage counters and the three floor-percentages  (D:5092-5136)
# --------------------------------------------------------------------------
def coverage_counters(invalid, preliminary_cloud_layer, ocean_mask=None):
    valid = ~invalid
    total = int(invalid.size)
    if ocean_mask is not None:
        valid = np.logical_and(valid, ocean_mask)              # D:5104
        n_not_ocean = int(np.sum(ocean_mask))                  # D:5105
    else:
        n_not_ocean = total                                    # D:5107
    n_valid = int(np.sum(valid))                               # D:5110
    n_cloud_and_valid = int(np.sum((preliminary_cloud_layer != 0) & valid))
    return n_valid, n_cloud_and_valid, n_not_ocean, total


def coverage_percentages(n_valid, n_cloud_and_valid, n_not_ocean, total):
    """(SPATIAL_COVERAGE, SPATIAL_COVERAGE_EXCLUDING_MASKED_OCEAN,
    CLOUD_COVERAGE) exactly as D:5115-5124 (float division then int())."""
    spatial = int(100 * float(n_valid) / total)
    cloud = 0 if n_valid == 0 else int(100 * float(n_cloud_and_valid) / n_valid)
    spatial_no_ocean = (0 if n_not_ocean == 0
                        else int(100 * float(n_valid) / n_not_ocean))
    return spatial, spatial_no_ocean, cloud


# --------------------------------------------------------------------------
# a11: aerosol class remapping, IN PLACE on wtr1 and cloud  (D:1210-1302)
# --------------------------------------------------------------------------
def apply_aerosol_class_remapping(wtr_1_layer, nir, preliminary_cloud_layer,
                                  fmask, values_not_water, values_moderate,
                                  values_psw_conservative,
                                  values_psw_aggressive):
    """Classes 0, 2, 3, 4 become 1 where fmask is in the class' list and
    nir <= 1000.0; bit 3 of CLOUD is set there unless CLOUD is 255.  The
    reference loops over the classes in the order 0, 2, 3, 4 (D:1283-1302);
    because the output class (1) is never an input class the four steps are
    independent and are evaluated here from a snapshot."""
    snapshot = wtr_1_layer.copy()
    low_nir = nir <= AEROSOL_REMAPPING_MAX_NIR                 # D:1239
    remap = np.zeros(snapshot.shape, dtype=bool)
    for cls, values in ((0, values_not_water), (2, values_moderate),
                        (3, values_psw_conservative),
                        (4, values_psw_aggressive)):
        remap |= np.isin(fmask, values) & (snapshot == cls) & low_nir
    wtr_1_layer[remap] = 1                                     # D:1240
    flag = remap & (preliminary_cloud_layer != UINT8_FILL_VALUE)  # D:1243-1244
    preliminary_cloud_layer[flag] |= 8                         # D:1245-1246
    return None


# --------------------------------------------------------------------------
# a12: land-cover and terrain-shadow masking -> WTR-2  (D:1305-1378)
# --------------------------------------------------------------------------
def apply_landcover_and_shadow_masks(interpreted_layer, nir, landcover_mask,
                                     shadow_layer, th):
    w1 = interpreted_layer
    water = (w1 >= 1) & (w1 <= 4)
    psw = (w1 == 3) | (w1 == 4)
    kill = np.zeros(w1.shape, dtype=bool)
    if shadow_layer is not None:                               # D:1331-1344
        in_shadow = (shadow_layer == 0)
        if landcover_mask is None:
            kill |= in_shadow & water
        else:
            kill |= in_shadow & (landcover_mask != LAND_WATER) & water
    if landcover_mask is not None:                             # D:1349-1376
        bright = nir > th.lcmask_nir
        low_dev = (landcover_mask >= 0) & (landcover_mask < 100)
        high_dev = (landcover_mask >= 100) & (landcover_mask < 200)
        kill |= (landcover_mask == LAND_EVERGREEN) & bright & psw
        kill |= low_dev & bright & psw
        kill |= high_dev & water
    out = w1.copy()
    out[kill] = 0
    return out


# --------------------------------------------------------------------------
# a13: snow -> CLOUD, mutates and returns cloud_layer  (D:1996-2086)
# --------------------------------------------------------------------------
def add_snow_to_cloud_layer(wtr_2_layer, cloud_layer, fmask,
                            mask_adjacent_to_cloud_mode):
    snow = (fmask & 16) != 0                                   # D:2052
    if mask_adjacent_to_cloud_mode == 'cover':                 # D:2055-2078
        from scipy.ndimage import binary_dilation
        adjacent = (fmask & 4) != 0
        grow_area = adjacent & (cloud_layer == 0)
        snow = binary_dilation(snow, iterations=10, mask=grow_area)
        grow_area = grow_area & (wtr_2_layer >= 1) & (wtr_2_layer <= 4)
        clear = (~snow) & (cloud_layer == 0)
        clear = binary_dilation(clear, iterations=7, mask=grow_area)
        snow[clear] = False
    cloud_layer[snow] += 2                                     # D:2081
    cloud_layer[wtr_2_layer == UINT8_FILL_VALUE] = UINT8_FILL_VALUE  # D:2084
    return cloud_layer


# --------------------------------------------------------------------------
# a14: CLOUD masking -> WTR  (D:2089-2133)
# --------------------------------------------------------------------------
def apply_cloud_masking(wtr_2_layer, cloud_layer):
    wtr = wtr_2_layer.copy()
    wtr[(cloud_layer != 0) & (cloud_layer != 8)] = WTR_CLOUD_MASKED   # D:2119
    wtr[(cloud_layer == 2) | (cloud_layer == 10)] = WTR_SNOW_MASKED   # D:2124
    wtr[wtr_2_layer == WTR_OCEAN_MASKED] = WTR_OCEAN_MASKED           # D:2128
    wtr[wtr_2_layer == UINT8_FILL_VALUE] = UINT8_FILL_VALUE           # D:2131
    return wtr


# --------------------------------------------------------------------------
# a15: BWTR  (D:1710-1730)
# --------------------------------------------------------------------------
def get_binary_water_layer(wtr_layer):
    out = wtr_layer.copy()
    out[(wtr_layer >= 1) & (wtr_layer <= 4)] = 1
    return out


# --------------------------------------------------------------------------
# a16: CONF  (D:1733-1837)
# --------------------------------------------------------------------------
def get_confidence_layer(wtr_2_layer, cloud_layer):
    """cloud/shadow/adjacent (CLOUD in {1,3,4,5,6,7,9,11,12,13,14,15}) -> +10;
    otherwise snow only when CLOUD == 2 exactly (not 10) -> +20 (D:1793-1835).
    The snow pass runs after the cloud pass on the already updated layer, so a
    pixel moved to 10..14 is not touched again."""
    conf = wtr_2_layer.copy()
    is_class = wtr_2_layer <= 4
    cloudy = np.isin(cloud_layer, [1, 3, 4, 5, 6, 7, 9, 11, 12, 13, 14, 15])
    conf[is_class & cloudy] += 10
    snowy = (cloud_layer == 2)
    conf[is_class & snowy & ~cloudy] += 20
    return conf


# --------------------------------------------------------------------------
# a17: collapse 4 water classes into 2 at save time  (D:2578-2598)
# --------------------------------------------------------------------------
def collapse_wtr_classes(layer):
    return COLLAPSE_LUT[layer]


# --------------------------------------------------------------------------
# a5: terrain shadow from the DEM  (D:4215-4283) and crop (D:4320-4337)
# --------------------------------------------------------------------------
def _central_differences(f, axis):
    """What np.gradient(f)[axis] evaluates for unit spacing (D:4255): interior
    (f[i+1] - f[i-1]) / 2.0, one-sided first differences at the two ends; the
    dtype of a floating input is kept."""
    f = np.asarray(f)
    if not np.issubdtype(f.dtype, np.floating):
        f = f.astype(np.float64)
    g = np.empty_like(f)
    fm = np.moveaxis(f, axis, 0)
    gm = np.moveaxis(g, axis, 0)
    gm[1:-1] = (fm[2:] - fm[:-2]) / 2.0
    gm[0] = (fm[1] - fm[0]) / 1.0
    gm[-1] = (fm[-1] - fm[-2]) / 1.0
    return g


def compute_opera_shadow_layer(dem, sun_azimuth_angle, sun_elevation_angle,
                               min_slope_angle, max_sun_local_inc_angle,
                               pixel_spacing_x=30, pixel_spacing_y=30, numpy1_promotion=False):
    """Bool mask, True = NOT shadow.  Array arithmetic stays in the DEM's
    dtype (float32 for the warped DEM) up to the normalisation factor; the
    products with the float64 sun-vector scalars are float64 under numpy >= 2
    (NEP 50), which is the semantics the parity target uses.

    ``numpy1_promotion``: numpy 1.x value-based casting instead (the reference
    pins numpy 1.23.5, setup.py:78): `float32_array * float64_scalar` stays
    float32 - the scalar is rounded to float32 first - so the dot product, the
    division, arccos / arctan and degrees all run in float32."""
    az = np.radians(sun_azimuth_angle)                         # D:4245
    zen = np.radians(90 - sun_elevation_angle)                 # D:4246-4247
    sun = (np.sin(az) * np.sin(zen), np.cos(az) * np.sin(zen), np.cos(zen))
    sin_az, cos_az = np.sin(az), np.cos(az)
    if numpy1_promotion and np.asarray(dem).dtype == np.float32:
        sun = tuple(np.float32(v) for v in sun)                # the cast value-based promotion applies to the scalar
        sin_az, cos_az = np.float32(sin_az), np.float32(cos_az)
    g_row = _central_differences(dem, 0)                       # D:4255
    g_col = _central_differences(dem, 1)
    nx = -g_col / pixel_spacing_x                              # D:4260
    ny = -g_row / -abs(pixel_spacing_y)                        # D:4261
    norm = np.sqrt(nx ** 2 + ny ** 2 + 1)                      # D:4264-4265
    with np.errstate(invalid='ignore'):
        inc_deg = np.degrees(np.arccos(
            (nx * sun[0] + ny * sun[1] + 1 * sun[2]) / norm))  # D:4267-4273
        dir_deg = np.degrees(np.arctan(
            nx * sin_az + ny * cos_az))                        # D:4275-4277
        backslope = dir_deg <= min_slope_angle                 # D:4279
        low_inc = inc_deg <= max_sun_local_inc_angle           # D:4280
    return low_inc | ~backslope                                # D:4281


def crop_2d_array_all_sides(a, margin):
    return a[margin:-margin, margin:-margin]                   # D:4336


# --------------------------------------------------------------------------
# numpy tail of create_landcover_mask  (D:874-904, D:1003-1115)
# --------------------------------------------------------------------------
LANDCOVER_THRESHOLDS = {'standard': (6, 3, 7, 3), 'water heavy': (6, 3, 7, 1)}     # D:270-271


def decimate_by_summation(image, size_y, size_x):
    """Block sum over size_y x size_x windows (D:874-904); the dtype of the
    input is kept (uint8 sums of 0/1 masks stay <= 9)."""
    h, w = image.shape
    assert h % size_y == 0 and w % size_x == 0, 'the reference is only used on exact multiples'
    blocks = image.reshape(h // size_y, size_y, w // size_x, size_x)
    return blocks.sum(axis=(1, 3), dtype=image.dtype)


def landcover_aggregate(worldcover_up_3, copernicus_landcover, forest_mask_landcover_classes,
                        year_offset, mask_type='standard'):
    """LAND layer from the ESA WorldCover raster warped to 10 m (3x the product
    grid) and the CGLS land-cover raster on the product grid, i.e. the
    statements of create_landcover_mask after its two GDAL warps:
    water {80, 90, 95}, urban 50 and tree 10 counts per 3x3 block (D:1003-1031),
    tree counts kept only on CGLS forest classes (D:1033-1043), then the
    hierarchy 255 -> evergreen 201 -> low intensity -> high intensity -> water
    200, later rules overriding earlier ones (D:1045-1115)."""
    th = LANDCOVER_THRESHOLDS[mask_type.lower()]
    water = decimate_by_summation(np.isin(worldcover_up_3, [80, 90, 95]).astype(np.uint8), 3, 3)
    urban = decimate_by_summation((worldcover_up_3 == 50).astype(np.uint8), 3, 3)
    tree = decimate_by_summation((worldcover_up_3 == 10).astype(np.uint8), 3, 3)
    forest = np.zeros(tree.shape, dtype=bool)
    for cls in (forest_mask_landcover_classes or ()):
        forest |= (copernicus_landcover == cls)
    tree = np.where(forest, tree, 0)
    land = np.full(water.shape, UINT8_FILL_VALUE, dtype=np.uint8)
    land[tree >= th[0]] = LAND_EVERGREEN                                   # D:1062
    land[urban >= th[1]] = 0 + year_offset                                 # D:1096-1100
    land[urban >= th[2]] = 100 + year_offset                               # D:1102-1107
    land[water >= th[3]] = LAND_WATER                                      # D:1109-1113
    return land


# --------------------------------------------------------------------------
# the chain, in the order of generate_dswx_layers  (D:5088-5369)
# --------------------------------------------------------------------------
def compute_browse_array(masked_interpreted_water_layer, flag_collapse_wtr_classes=True,
                         exclude_psw_aggressive=False, set_not_water_to_nodata=False,
                         set_cloud_to_nodata=False, set_snow_to_nodata=False,
                         set_ocean_masked_to_nodata=True):
    """_compute_browse_array, dswx_hls.py:3057-3129."""
    b = np.array(masked_interpreted_water_layer, copy=True)
    if exclude_psw_aggressive:
        b[b == 4] = 0                                     # :3107-3110
    if flag_collapse_wtr_classes:
        b = collapse_wtr_classes(b)                       # :3112-3113
    if set_not_water_to_nodata:
        b[b == 0] = 255                                   # :3115-3116
    if set_cloud_to_nodata:
        b[b == 253] = 255                                 # :3118-3119
    if set_snow_to_nodata:
        b[b == 252] = 255                                 # :3121-3122
    if set_ocean_masked_to_nodata:
        b[b == 254] = 255                                 # :3124-3125
    return b


def scale_and_offset_band(image, scale_factor, offset, invalid_ind=None):
    """dswx_hls.py:2301-2302 and :3024-3036: float32 array arithmetic with Python-float scalars."""
    out = float(scale_factor) * (np.asarray(image, dtype=np.float32) - float(offset))
    if invalid_ind is not None:
        out[invalid_ind] = np.nan
    return out


def otsu_threshold(image, is_normalized=True):
    """The threshold of _compute_otsu_threshold, dswx_hls.py:1638-1684: 256-bin histogram
    between the image's extremes, class weights and means by cumulative sums, first maximum of the inter-class
    variance, bin centre below it."""
    counts, edges = np.histogram(image, bins=256)                         # :1663
    h = counts.ravel() / counts.max() if is_normalized else counts        # :1666-1667
    centres = (edges[:-1] + edges[1:]) / 2.                               # :1670
    with np.errstate(all='ignore'):
        w_lo, w_hi = np.cumsum(h), np.cumsum(h[::-1])[::-1]               # :1673-1674
        mu_lo = np.cumsum(h * centres) / w_lo                             # :1677
        mu_hi = (np.cumsum((h * centres)[::-1]) / w_hi[::-1])[::-1]       # :1679
        between = w_lo[:-1] * w_hi[1:] * (mu_lo[:-1] - mu_hi[1:]) ** 2    # :1681
    return centres[:-1][np.argmax(between)]                               # :1684-1686


def compute_otsu_threshold(image, is_normalized=True):
    """_compute_otsu_threshold: ``image > threshold`` (:1689)."""
    return image > otsu_threshold(image, is_normalized)


def compute_hillshade_gdal(dem, sun_azimuth_angle, sun_elevation_angle, ewres=30.0, nsres=-30.0):
    """PARITY UNPINNED restatement of what ``_compute_hillshade`` (D:4177-4212) gets back from
    ``gdal.DEMProcessing(..., "hillshade", azimuth, altitude)``: GDAL 3.6.2 is not vendored in the reference tree and is
    absent from this image, so neither its output nor a reference-owned vector can pin these numbers.  This is the
    published gdaldem algorithm (gdaldem_lib.cpp: GDALCreateHillshadeData + GDALHillshadeAlg, Horn gradient, defaults
    z = 1, scale = 1, no -compute_edges, Byte output) for a float32 band:

        x = ((w0 + w3 + w3 + w6) - (w2 + w5 + w5 + w8)) / ewres      (window sums in float32, left to right)
        y = ((w6 + w7 + w7 + w8) - (w0 + w1 + w1 + w2)) / nsres      (w0..w2 = northern row)
        c = (254 sin(alt) - (y * 254 cos(az) cos(alt) z/8 - x * 254 sin(az) cos(alt) z/8)) / sqrt(1 + (z/8)^2 (x^2 + y^2))
        shade = 1 if c <= 0 else 1 + c;  Byte = trunc(float32(shade) + 0.5);  border rows / columns = 0 (no data)

    SSE2 builds of GDAL approximate the division with rsqrt + one Newton step: a shade within ~1e-6 of a rounding
    boundary may differ there by one grey level."""
    d = np.asarray(dem)
    if d.dtype != np.float32 or d.ndim != 2:
        raise TypeError('compute_hillshade_gdal: a 2-D float32 DEM')
    rows, cols = d.shape
    out = np.zeros((rows, cols), np.uint8)
    if rows < 3 or cols < 3:
        return out
    w = [[d[i:rows - 2 + i, j:cols - 2 + j] for j in range(3)] for i in range(3)]      # w[row][col], float32 views
    xs = (((w[0][0] + w[1][0]) + w[1][0]) + w[2][0]) - (((w[0][2] + w[1][2]) + w[1][2]) + w[2][2])
    ys = (((w[2][0] + w[2][1]) + w[2][1]) + w[2][2]) - (((w[0][0] + w[0][1]) + w[0][1]) + w[0][2])
    assert xs.dtype == np.float32 and ys.dtype == np.float32
    deg2rad = 3.14159265358979323846 / 180.0
    z_scaled = 1.0 / 8.0
    cos_alt_z = np.cos(sun_elevation_angle * deg2rad) * z_scaled
    sin_alt_254 = 254.0 * np.sin(sun_elevation_angle * deg2rad)
    cos_az_254 = 254.0 * (np.cos(sun_azimuth_angle * deg2rad) * cos_alt_z)
    sin_az_254 = 254.0 * (np.sin(sun_azimuth_angle * deg2rad) * cos_alt_z)
    x = xs.astype(np.float64) * (1.0 / ewres)
    y = ys.astype(np.float64) * (1.0 / nsres)
    with np.errstate(all='ignore'):
        c254 = (sin_alt_254 - (y * cos_az_254 - x * sin_az_254)) / np.sqrt(1.0 + (z_scaled * z_scaled) * (x * x + y * y))
        cang = np.where(c254 <= 0.0, 1.0, 1.0 + c254)
        f = cang.astype(np.float32) + np.float32(0.5)
        b = np.where(f > 0, np.minimum(f, np.float32(255.0)), np.float32(0.0))      # NaN -> 0
        out[1:-1, 1:-1] = np.trunc(b).astype(np.uint8)
    return out


def compute_otsu_shadow_layer(dem, sun_azimuth_angle, sun_elevation_angle, ewres=30.0, nsres=-30.0):
    """The 'otsu' branch of D:5152-5157: hillshade (above, parity unpinned) -> _compute_otsu_threshold(normalized)."""
    return compute_otsu_threshold(compute_hillshade_gdal(dem, sun_azimuth_angle, sun_elevation_angle, ewres, nsres), True)


def reference_chain(raw_bands, fmask, dem_with_margin=None, landcover=None,
                    ocean_mask=None, sun_azimuth_angle=150.0,
                    sun_elevation_angle=45.0, thresholds=None,
                    processing=None, band_fill=-9999, fmask_fill=255,
                    dem_margin=DEM_MARGIN_IN_PIXELS, numpy1_promotion=False):
    """All layers of the hot path for one tile.  ``raw_bands`` = 6 int16
    arrays in BAND_NAMES order, straight from the file (before clipping)."""
    th = thresholds or default_thresholds()
    pr = dict(default_processing())
    pr.update(processing or {})
    mode = pr['mask_adjacent_to_cloud_mode']

    invalid, (blue, green, red, nir, swir1, swir2) = invalid_mask_and_clip(
        raw_bands, fmask, band_fill, fmask_fill)
    prelim = compute_preliminary_cloud_layer(fmask, mode)      # D:5089
    counters = coverage_counters(invalid, prelim, ocean_mask)  # D:5092-5111

    shad = None
    if dem_with_margin is not None:                            # D:5161-5167
        shad_m = compute_opera_shadow_layer(
            dem_with_margin, sun_azimuth_angle, sun_elevation_angle,
            pr['min_slope_angle'], pr['max_sun_local_inc_angle'], numpy1_promotion=numpy1_promotion)
        shad = (crop_2d_array_all_sides(shad_m, dem_margin)
                if dem_margin else shad_m)

    diag_dec = compute_diagnostic_tests(blue, green, red, nir, swir1, swir2, th)
    diag_dec[invalid] = DIAG_FILL_DECIMAL                      # D:5227
    wtr1 = generate_interpreted_layer(diag_dec)                # D:5229
    diag = get_binary_representation(diag_dec)                 # D:5231
    if ocean_mask is not None:
        wtr1[ocean_mask == 0] = WTR_OCEAN_MASKED               # D:5245
    wtr1[invalid] = UINT8_FILL_VALUE                           # D:5249
    wtr1_saved = wtr1.copy()          # what the WTR-1 file holds (D:5251-5258)

    cloud = prelim.copy()
    if pr['apply_aerosol_class_remapping']:                    # D:5260-5266
        apply_aerosol_class_remapping(
            wtr1, nir, cloud, fmask,
            pr['aerosol_not_water_to_high_conf_water_fmask_values'],
            pr['aerosol_water_moderate_conf_to_high_conf_water_fmask_values'],
            pr['aerosol_partial_surface_water_conservative_to_high_conf_water_fmask_values'],
            pr['aerosol_partial_surface_aggressive_to_high_conf_water_fmask_values'])
    wtr2 = apply_landcover_and_shadow_masks(wtr1, nir, landcover, shad, th)
    cloud = add_snow_to_cloud_layer(wtr2, cloud, fmask, mode)  # D:5282
    wtr = apply_cloud_masking(wtr2, cloud)                     # D:5286
    bwtr = get_binary_water_layer(wtr)                         # D:5358
    conf = get_confidence_layer(wtr2, cloud)                   # D:5368

    out = dict(
        DIAG=diag, WTR1=wtr1_saved, WTR1_REMAPPED=wtr1, WTR2=wtr2,
        CLOUD=cloud, WTR=wtr, BWTR=bwtr, CONF=conf,
        WTR_COLLAPSED=collapse_wtr_classes(wtr),               # D:2688-2689
        WTR1_COLLAPSED=collapse_wtr_classes(wtr1_saved),
        WTR2_COLLAPSED=collapse_wtr_classes(wtr2),
        counters=np.array(counters[:3], dtype=np.uint64),
        percentages=np.array(coverage_percentages(*counters), dtype=np.int64))
    if shad is not None:
        out['SHAD'] = shad.astype(np.uint8)
    return out
