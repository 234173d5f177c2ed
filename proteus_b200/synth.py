"""Deterministic synthetic HLS tiles.

There is no network and no GDAL in the build or GPU images, so every test and
benchmark runs on synthetic rasters shaped like one HLS tile: six int16
reflectance bands + uint8 Fmask (what ``gdal.ReadAsArray`` hands to
``generate_dswx_layers``, /root/reference/src/proteus/dswx_hls.py:2192), a
float32 DEM with the 50-px margin (D:5145-5150), a uint8 LAND raster
(D:5196-5201) and a uint8 {0,1} ocean mask (D:5097-5101).

Two generators:
  * ``make_tile``   - numpy, seeded with ``1000 + tile_id``; used for parity
                      tests, golden fixtures and the CPU baseline.
  * ``make_device_batch`` - torch, on the GPU; used only to fill large
                      device-resident benchmark batches quickly.  Parity on
                      those batches is checked by copying sampled tiles back.
"""
from __future__ import annotations

import numpy as np

BAND_NAMES = ('blue', 'green', 'red', 'nir', 'swir1', 'swir2')
DEM_MARGIN = 50
HLS_TILE = 3660

# per surface type: (lo, hi) per band in BAND_NAMES order
_TYPE_RANGES = {
    'water': ((300, 800), (300, 900), (200, 700), (50, 600), (10, 400), (5, 300)),
    'veg':   ((200, 700), (300, 900), (200, 800), (2500, 5000), (1000, 2500), (500, 1500)),
    'soil':  ((1000, 3500),) * 6,
    'snow':  ((5000, 12000),) * 6,
}
_TYPE_ORDER = ('water', 'veg', 'soil', 'snow')
_TYPE_QUANTILES = (0.25, 0.60, 0.90)     # cumulative shares of the first three


def _smooth_field(rng, height, width, cell):
    """Unit-variance smooth random field: coarse N(0,1) grid, lightly blurred,
    spline-zoomed to (height, width).  ``cell`` = feature size in pixels."""
    from scipy import ndimage
    ch = max(4, int(np.ceil(height / cell)) + 3)
    cw = max(4, int(np.ceil(width / cell)) + 3)
    coarse = ndimage.gaussian_filter(
        rng.standard_normal((ch, cw)).astype(np.float32), 1.0, mode='wrap')
    zoomed = ndimage.zoom(coarse, (height / (ch - 3), width / (cw - 3)),
                          order=3, mode='nearest', grid_mode=False)
    out = np.ascontiguousarray(zoomed[:height, :width], dtype=np.float32)
    out -= out.mean()
    out /= max(float(out.std()), 1e-6)
    return out


def make_tile(tile_id=0, height=HLS_TILE, width=HLS_TILE, *, with_dem=True,
              with_land=True, with_ocean=True, dem_sigma_m=400.0,
              dem_margin=DEM_MARGIN, sun=None, adversarial=False):
    """Return a dict of host arrays for one synthetic tile.

    keys: 'bands' (list of 6 int16 [H,W], RAW values incl. -9999 fill),
    'fmask' uint8, 'dem' float32 [(H+2m),(W+2m)] or None, 'land' uint8 or
    None, 'ocean' uint8 or None, 'sun_azimuth', 'sun_elevation'.

    adversarial=True replaces the band values by full-range int16 noise
    (forces int16 wrap-around in every sum) and Fmask / LAND by uniform
    bytes."""
    rng = np.random.default_rng(1000 + int(tile_id))
    h, w = int(height), int(width)
    cell = max(8.0, min(h, w) / 24.0)

    surf = _smooth_field(rng, h, w, cell)
    q = np.quantile(surf[::7, ::7], _TYPE_QUANTILES)
    type_idx = np.digitize(surf, q).astype(np.uint8)        # 0..3

    bands = []
    if adversarial:
        for _ in range(6):
            bands.append(rng.integers(-32768, 32768, (h, w), dtype=np.int16))
    else:
        u_noise = rng.random((h, w), dtype=np.float32)
        outlier = rng.random((h, w), dtype=np.float32)
        for b in range(6):
            lo = np.array([_TYPE_RANGES[t][b][0] for t in _TYPE_ORDER],
                          dtype=np.float32)[type_idx]
            hi = np.array([_TYPE_RANGES[t][b][1] for t in _TYPE_ORDER],
                          dtype=np.float32)[type_idx]
            per_band = rng.random((h, w), dtype=np.float32)
            val = lo + (hi - lo) * (0.5 * u_noise + 0.5 * per_band)
            band = val.astype(np.int16)
            # 10 % broadband noise incl. negative reflectances (clip path)
            m = outlier < 0.10
            band[m] = rng.integers(-200, 6001, int(m.sum()), dtype=np.int16)
            # 0.1 % saturated values: g + s1 etc. wrap in int16
            m = outlier > 0.999
            band[m] = rng.integers(16000, 32768, int(m.sum()), dtype=np.int16)
            bands.append(band)

    # fill wedge (~2 % of the tile, like an HLS swath edge): all rasters
    yy, xx = np.mgrid[0:h, 0:w]
    wedge = (xx + 0.35 * yy) < 0.08 * w * (1.0 - yy / max(h, 1))
    # a thin second wedge where only SWIR-2 is fill (cumulative invalid mask)
    single = ((w - 1 - xx) + 0.2 * yy) < 0.01 * w
    for band in bands:
        band[wedge] = -9999
    bands[5][single] = -9999

    # Fmask bit fields (D:1958-1972): 1 cloud, 2 adjacent, 3 shadow, 4 snow,
    # 5 water, 6-7 aerosol level
    if adversarial:
        fmask = rng.integers(0, 256, (h, w), dtype=np.uint8)
    else:
        from scipy import ndimage
        cloud_f = _smooth_field(rng, h, w, cell * 0.8)
        cloud = cloud_f > np.quantile(cloud_f[::7, ::7], 0.85)
        ring = ndimage.binary_dilation(cloud, iterations=3) & ~cloud
        shadow_f = np.roll(cloud_f, (int(cell // 2), int(cell // 2)), (0, 1))
        shadow = (shadow_f > np.quantile(shadow_f[::7, ::7], 0.95)) & ~cloud
        snow = (type_idx == 3) & (rng.random((h, w), dtype=np.float32) < 0.3)
        aerosol = rng.integers(0, 4, (h, w), dtype=np.uint8)
        fmask = ((cloud.astype(np.uint8) << 1) | (ring.astype(np.uint8) << 2) |
                 (shadow.astype(np.uint8) << 3) | (snow.astype(np.uint8) << 4) |
                 ((type_idx == 0).astype(np.uint8) << 5) |
                 (aerosol << 6)).astype(np.uint8)
    fmask[wedge] = 255

    out = dict(tile_id=int(tile_id), height=h, width=w, bands=bands,
               fmask=fmask, dem=None, land=None, ocean=None,
               dem_margin=int(dem_margin))

    if with_dem:
        m = int(dem_margin)
        dem = _smooth_field(rng, h + 2 * m, w + 2 * m, max(8.0, cell * 0.6))
        dem = (dem * np.float32(dem_sigma_m) + np.float32(800.0))
        # quantise like a cubic-warped metre DEM (float32, non-trivial bits)
        out['dem'] = np.ascontiguousarray(dem, dtype=np.float32)

    if with_land:
        if adversarial:
            land = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            lf = _smooth_field(rng, h, w, cell * 0.5)
            ql = np.quantile(lf[::7, ::7], (0.70, 0.78, 0.88, 0.94))
            land = np.array([255, 200, 201, 21, 121],
                            dtype=np.uint8)[np.digitize(lf, ql)]
        out['land'] = np.ascontiguousarray(land)

    if with_ocean:
        shore = 0.9 * w + 0.03 * w * np.sin(np.arange(h) / max(h, 1) * 9.0)
        ocean = (np.arange(w)[None, :] < shore[:, None]).astype(np.uint8)
        out['ocean'] = np.ascontiguousarray(ocean)

    if sun is None:
        # time-series style: seeded per acquisition, az 120-170, el 20-65
        sun = (float(rng.uniform(120.0, 170.0)), float(rng.uniform(20.0, 65.0)))
        if tile_id == 0:
            sun = (150.0, 45.0)
    out['sun_azimuth'], out['sun_elevation'] = float(sun[0]), float(sun[1])
    return out


def make_device_batch(n_tiles, height=HLS_TILE, width=HLS_TILE, *,
                      device='cuda', seed=1000, n_distinct=4,
                      shared_ancillary=False, dem_margin=DEM_MARGIN,
                      full_product=True, adversarial=False, wedge=0.08):
    """Fill ``n_tiles`` device-resident tiles with torch RNG (benchmark data).

    Cheap, blocky but representative: surface type per 64x64 cell, per-pixel
    uniform noise inside the type's range, the same outlier / wrap / fill
    shares as ``make_tile``.  Only ``n_distinct`` different tiles are drawn;
    the rest are rolled copies (distinct memory, so nothing is cache-hot).

    ``adversarial=True``: full-range int16 noise in every band (the int16 sums of about 40 % of the pixels wrap
    and take the kernel's scalar patch path) and uniform random Fmask / LAND bytes, like ``make_tile``'s - the
    data-dependent worst case.

    ``wedge``: width of the diagonal no-data wedge at the top-left corner as a share of the tile width (0.08 = the 2 %
    of fill pixels of SURVEY 8d; 0.8 = a tile at the edge of a swath, about a third of it fill).

    Returns a list of dicts of torch tensors with the ``make_tile`` keys."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    h, w, m = int(height), int(width), int(dem_margin)
    lo = torch.tensor([[r[0] for r in _TYPE_RANGES[t]] for t in _TYPE_ORDER],
                      dtype=torch.float32, device=device)
    hi = torch.tensor([[r[1] for r in _TYPE_RANGES[t]] for t in _TYPE_ORDER],
                      dtype=torch.float32, device=device)

    def coarse_to_full(c, hh, ww, cellsz):
        return c.repeat_interleave(cellsz, 0).repeat_interleave(cellsz, 1)[:hh, :ww]

    def draw_one():
        cs = 64
        ch, cw = -(-h // cs), -(-w // cs)
        t_c = torch.multinomial(
            torch.tensor([0.25, 0.35, 0.30, 0.10], device=device),
            ch * cw, True, generator=g).reshape(ch, cw)
        tix = coarse_to_full(t_c, h, w, cs)
        outlier = torch.rand((h, w), device=device, generator=g)
        yy = torch.arange(h, device=device)[:, None]
        xx = torch.arange(w, device=device)[None, :]
        wedge_mask = (xx + 0.35 * yy) < float(wedge) * w * (1.0 - yy / max(h, 1))
        bands = []
        for b in range(6):
            u = torch.rand((h, w), device=device, generator=g)
            val = (lo[:, b][tix] + (hi[:, b] - lo[:, b])[tix] * u).to(torch.int16)
            noise = torch.randint(-200, 6001, (h, w), device=device,
                                  generator=g, dtype=torch.int16)
            sat = torch.randint(16000, 32768, (h, w), device=device,
                                generator=g, dtype=torch.int16)
            val = torch.where(outlier < 0.10, noise, val)
            val = torch.where(outlier > 0.999, sat, val)
            if adversarial:
                val = torch.randint(-32768, 32768, (h, w), device=device, generator=g, dtype=torch.int16)
            val = torch.where(wedge_mask, torch.full_like(val, -9999), val)
            bands.append(val.contiguous())
        cl_c = torch.rand((ch, cw), device=device, generator=g)
        cloud = coarse_to_full(cl_c < 0.15, h, w, cs)
        adj = coarse_to_full((cl_c >= 0.15) & (cl_c < 0.20), h, w, cs)
        shd = coarse_to_full((cl_c >= 0.20) & (cl_c < 0.25), h, w, cs)
        snow = (tix == 3) & (torch.rand((h, w), device=device, generator=g) < 0.3)
        aer = torch.randint(0, 4, (h, w), device=device, generator=g,
                            dtype=torch.uint8)
        fmask = ((cloud.to(torch.uint8) << 1) | (adj.to(torch.uint8) << 2) |
                 (shd.to(torch.uint8) << 3) | (snow.to(torch.uint8) << 4) |
                 ((tix == 0).to(torch.uint8) << 5) | (aer << 6))
        if adversarial:
            fmask = torch.randint(0, 256, (h, w), device=device, generator=g, dtype=torch.uint8)
        fmask = torch.where(wedge_mask, torch.full_like(fmask, 255), fmask).contiguous()
        d = dict(height=h, width=w, bands=bands, fmask=fmask, dem=None,
                 land=None, ocean=None, dem_margin=m)
        if full_product:
            d.update(draw_ancillary())
        return d

    def draw_ancillary():
        hh, ww = h + 2 * m, w + 2 * m
        # smooth DEM: bilinear upsample of a coarse field (slopes up to ~0.5)
        c = torch.randn((1, 1, hh // 25 + 2, ww // 25 + 2), device=device,
                        generator=g)
        dem = torch.nn.functional.interpolate(
            c, size=(hh, ww), mode='bicubic', align_corners=True)[0, 0]
        dem = (dem * 400.0 + 800.0).to(torch.float32).contiguous()
        lc = torch.multinomial(
            torch.tensor([0.70, 0.08, 0.10, 0.06, 0.06], device=device),
            (-(-h // 32)) * (-(-w // 32)), True, generator=g
        ).reshape(-(-h // 32), -(-w // 32))
        land = torch.tensor([255, 200, 201, 21, 121], dtype=torch.uint8,
                            device=device)[coarse_to_full(lc, h, w, 32)].contiguous()
        if adversarial:
            land = torch.randint(0, 256, (h, w), device=device, generator=g, dtype=torch.uint8)
        shore = 0.9 * w + 0.03 * w * torch.sin(
            torch.arange(h, device=device) / max(h, 1) * 9.0)
        ocean = (torch.arange(w, device=device)[None, :] <
                 shore[:, None]).to(torch.uint8).contiguous()
        return dict(dem=dem, land=land, ocean=ocean)

    distinct = [draw_one() for _ in range(min(n_distinct, n_tiles))]
    shared = draw_ancillary() if (shared_ancillary and full_product) else None
    tiles = []
    for i in range(n_tiles):
        src = distinct[i % len(distinct)]
        if i < len(distinct):
            t = dict(src)
        else:
            sh = 8 * (i // len(distinct))
            t = dict(src)
            t['bands'] = [torch.roll(b, sh, 1).contiguous() for b in src['bands']]
            t['fmask'] = torch.roll(src['fmask'], sh, 1).contiguous()
            if full_product and not shared_ancillary:
                t['dem'] = src['dem'].clone()
                t['land'] = src['land'].clone()
                t['ocean'] = src['ocean'].clone()
        if shared is not None:
            t.update(shared)
        t['tile_id'] = i
        az = 120.0 + 50.0 * ((i * 37) % 101) / 100.0
        el = 20.0 + 45.0 * ((i * 53) % 101) / 100.0
        t['sun_azimuth'], t['sun_elevation'] = (150.0, 45.0) if i == 0 else (az, el)
        tiles.append(t)
    return tiles


def make_guard_band_dem(height, width, sun_azimuth=150.0, sun_elevation=56.0, *, min_slope_angle=-5.0,
                        max_sun_local_inc_angle=40.0, margin=DEM_MARGIN, period=16, seed=0):
    """A float32 DEM (with margin) whose pixels sit ON the two decision boundaries of the terrain-shadow test
    (dswx_hls.py:4264-4281) - the worst case for a float32 shortcut with guard bands, and the place where numpy 1.x
    (float32) and numpy >= 2 (float64) promotion can decide differently.

    Rows of the upper half: directional slope = min_slope (1 + delta) with a cross-slope that keeps the incidence angle
    above max_inc (so the slope test alone decides); lower half: a back slope whose incidence angle is
    max_inc (1 + delta) (needs 90 - sun_elevation + |slope| == max_inc: the defaults with elevation 56 and a 6 degree
    slope).  delta is constant per 16 x 16 block, drawn from {0, +-1e-7 ... +-1e-3}; the surface is a sawtooth of
    planes (period `period` pixels) so that heights stay small and float32 rounding of the heights - about one quantum of
    the gradient - scatters the pixels across the boundary.  Values at the sawtooth jumps are arbitrary slopes."""
    rng = np.random.default_rng(7000 + seed)
    h, w = height + 2 * margin, width + 2 * margin
    az = np.radians(sun_azimuth)
    zen = 90.0 - sun_elevation
    deltas = np.array([0.0, 0.0, 1e-7, -1e-7, 3e-7, -3e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-3, -1e-3])
    by, bx = -(-h // period), -(-w // period)
    delta = np.kron(deltas[rng.integers(0, len(deltas), (by, bx))], np.ones((period, period)))[:h, :w]
    yy, xx = np.mgrid[0:h, 0:w]
    upper = yy < h // 2
    # slope along the sun azimuth (nx, ny) = t (sin az, cos az) + c (cos az, -sin az)
    t_upper = np.tan(np.radians(min_slope_angle)) * (1.0 + delta)
    theta = -(max_sun_local_inc_angle * (1.0 + delta) - zen)          # incidence = zen - theta for a slope along the azimuth
    t_lower = np.tan(np.radians(theta))
    t = np.where(upper, t_upper, t_lower)
    c = np.where(upper, 0.5, 0.0)
    nx = t * np.sin(az) + c * np.cos(az)
    ny = t * np.cos(az) - c * np.sin(az)
    g_col, g_row = -30.0 * nx, 30.0 * ny                              # nx = -g_col / 30, ny = -g_row / -30 (D:4260-4261)
    dem = g_col * (xx % period) + g_row * (yy % period)
    return dem.astype(np.float32)
