"""World-size-2 tests of the multi-GPU host logic on CPU (gloo backend):
tile sharding, row-strip partitioning and the DEM halo exchange of the mosaic
path (proteus_b200/mosaic.py).  No CUDA involved."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, height, width, margin, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from proteus_b200 import mosaic
        rng = np.random.default_rng(77)                       # same full DEM on every rank
        dem_full = rng.normal(size=(height + 2 * margin, width + 2 * margin)).astype(np.float32)
        bounds = mosaic.strip_bounds(height, world)
        r0, r1 = bounds[rank]
        n = r1 - r0
        d0, d1 = mosaic.dem_rows_for_strip(r0, r1, height, margin)
        dem_local = torch.from_numpy(dem_full[d0:d1].copy())
        first = margin + r0 - d0
        dem_ext = torch.full((n + 2, dem_full.shape[1]), float('nan'))
        dem_ext[1:n + 1] = dem_local[first:first + n]
        if r0 == 0:
            dem_ext[0] = dem_local[first - 1]
        if r1 == height:
            dem_ext[n + 1] = dem_local[first + n]
        for req in mosaic.exchange_dem_halo(dem_ext, n, rank, world):
            req.wait()
        expect = dem_full[margin + r0 - 1: margin + r1 + 1]
        ok = np.array_equal(dem_ext.numpy(), expect)
        # tile sharding + a max-over-ranks reduction like bench.py's timing
        mine = mosaic.shard_tiles(7, rank, world)
        t = torch.tensor([float(len(mine))], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counters = torch.tensor([n, rank + 1, 1], dtype=torch.int64)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
        # Otsu threshold of a raster split in row strips: one all-reduce of the per-value counts
        hill = np.clip(np.random.default_rng(5).normal(170, 40, (height, width)), 0, 255).astype(np.uint8)
        thr = mosaic.otsu_threshold_of_strips(np.bincount(hill[r0:r1].ravel(), minlength=256))
        from oracle import dswx_oracle as O
        otsu_ok = thr == O.otsu_threshold(hill)
        np.save(os.path.join(out_dir, f'r{rank}.npy'),
                np.array([int(ok), r0, r1, len(mine), int(t.item()), *counters.tolist(), int(otsu_ok)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('height', [64, 200])
def test_halo_exchange_and_sharding_world2(tmp_path, height):
    world, width, margin = 2, 24, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, height, width, margin, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f'r{r}.npy') for r in range(world)]
    assert all(r[0] == 1 for r in res), 'halo rows do not match the full DEM'
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == height      # strips partition the rows
    assert res[0][2] % 32 == 0
    assert [r[3] for r in res] == [4, 3] and all(r[4] == 4 for r in res)         # 7 tiles round robin, max = 4
    assert all(r[5] == height and r[6] == 3 and r[7] == 2 for r in res)          # summed counters
    assert all(r[8] == 1 for r in res), 'threshold from all-reduced strip histograms != whole-raster Otsu threshold'


def test_strip_bounds_properties():
    from proteus_b200 import mosaic
    for h in (1, 31, 32, 33, 2745, 21960):
        for w in (1, 2, 3, 8):
            b = mosaic.strip_bounds(h, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == h
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(r0 % 32 == 0 for r0, r1 in b if r1 > r0)       # empty strips sit at the end
            rows = [d for r0, r1 in b for d in range(*mosaic.dem_rows_for_strip(r0, r1, h, 50)) if r1 > r0]
            if all(r1 > r0 for r0, r1 in b):
                assert rows == list(range(h + 100))                                  # DEM rows partitioned
    assert mosaic.shard_tiles(64, 3, 8) == list(range(3, 64, 8))
