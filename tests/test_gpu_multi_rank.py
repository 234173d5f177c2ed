"""Two ranks on two GPUs of one box: the row-stripped mosaic (BASELINE configs[4]) with the REAL halo exchange -
NCCL send / recv through torch.distributed and through the C ABI's own communicator (pb200_comm_init /
pb200_halo_exchange_dem) - against the oracle run on the whole raster.  The only non-local function of the path is the
one-row stencil of np.gradient (dswx_hls.py:4255); the seam between the strips is where a mosaic can go wrong.

Skipped on a box with fewer than two GPUs (NCCL refuses two ranks on one device)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

LAYERS = ('DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD', 'SHAD', 'WTR', 'BWTR', 'CONF')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, h, w, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    res = {}
    try:
        import proteus_b200 as pb
        from proteus_b200 import mosaic, synth
        from oracle import dswx_oracle as O
        m = 50
        t = synth.make_tile(91, h, w)                          # the same whole raster on every rank
        ref = O.reference_chain(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'],
                                t['sun_azimuth'], t['sun_elevation'])
        r0, r1 = mosaic.strip_bounds(h, world)[rank]
        d0, d1 = mosaic.dem_rows_for_strip(r0, r1, h, m)
        dev = f'cuda:{rank}'
        ctx = pb.get_context(rank)
        mosaic.init_library_comm(ctx, rank, world)
        params = pb.make_params(collapse_wtr_classes=False)

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        for exchange in ('torch', 'library'):
            for overlap in (True, False):
                strip = mosaic.MosaicStrip(
                    [up(b[r0:r1]) for b in t['bands']], up(t['fmask'][r0:r1]), up(t['dem'][d0:d1]),
                    up(t['land'][r0:r1]), up(t['ocean'][r0:r1]), r0, r1, h,
                    sun_azimuth=t['sun_azimuth'], sun_elevation=t['sun_elevation'], params=params,
                    outputs=LAYERS, rank=rank, world=world, ctx=ctx, exchange=exchange, overlap=overlap)
                # poison the halo rows the exchange has to fill (rank 0 keeps its top row, the last rank its bottom row)
                if rank > 0:
                    strip.dem_ext[0].fill_(float('nan'))
                if rank < world - 1:
                    strip.dem_ext[strip.n_rows + 1].fill_(float('nan'))
                for launch in range(2):                        # the second run re-exchanges into live buffers
                    strip.zero_counters()
                    strip.run()
                got = strip.results()
                bad = {k: int((got[k] != ref[k][r0:r1]).sum()) for k in LAYERS}
                halo_ok = bool(torch.equal(strip.dem_ext.cpu(), torch.from_numpy(t['dem'][m + r0 - 1:m + r1 + 1])))
                total = strip.allreduce_counters().cpu().numpy()[0, :3].astype(np.uint64)
                res[f'{exchange}/{overlap}'] = dict(bad=bad, halo_ok=halo_ok,
                                                    counters_ok=bool(np.array_equal(total, ref['counters'])))
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()
    import json
    with open(os.path.join(out_dir, f'r{rank}.json'), 'w') as f:
        json.dump(res, f)


def test_two_rank_mosaic_with_real_nccl_exchange_matches_the_whole_raster_oracle(pb, tmp_path):
    import json
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs on the box (gpurun --gpus 2)')
    world, h, w = 2, 352, 520            # seam at row 192 (strip starts are multiples of 32), ragged item edges
    mp.spawn(_worker, args=(world, _free_port(), h, w, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        with open(tmp_path / f'r{rank}.json') as f:
            res = json.load(f)
        assert set(res) == {'torch/True', 'torch/False', 'library/True', 'library/False'}
        for key, r in res.items():
            assert r['halo_ok'], (rank, key, 'halo rows differ from the whole DEM')
            assert all(v == 0 for v in r['bad'].values()), (rank, key, r['bad'])
            assert r['counters_ok'], (rank, key, 'all-reduced counters differ from the whole-raster counters')
