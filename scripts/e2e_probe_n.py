"""What limits the end-to-end (host buffer) path at N GPUs: run under torchrun with N ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/e2e_probe_n.py

Every rank moves the byte pattern of one HLS tile per step (256 MB in, 67 MB out), all ranks at once:
  h2d_only / d2h_only / both  - plain copies between pinned buffers and the device (no kernel), mean and best per tile
  pipeline                     - proteus_b200.TilePipeline (the bench's e2e call), mean per tile
Rank 0 prints one JSON object (per-tile milliseconds: max over ranks; GB/s: sum over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import proteus_b200 as pb
    from proteus_b200 import synth
    size = 3660
    dev = torch.device('cuda', local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def pinned(shape, dt):
        return torch.from_numpy(pb.pinned_empty(shape, dt))
    ins = [pinned((size, size), np.int16) for _ in range(6)] + [pinned((size, size), np.uint8) for _ in range(3)]
    dem = pinned((size + 100, size + 100), np.float32)
    outs = [pinned((size, size), np.int16)] + [pinned((size, size), np.uint8) for _ in range(3)]
    for x in ins + outs + [dem]:
        x.zero_()
    dins = [torch.empty_like(x, device=dev) for x in ins]
    ddem = torch.empty_like(dem, device=dev)
    douts = [torch.zeros_like(x, device=dev) for x in outs]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    h2d_bytes = sum(x.numel() * x.element_size() for x in ins) + (size + 2) * (size + 100) * 4
    d2h_bytes = sum(x.numel() * x.element_size() for x in outs)

    def copies(do_in, do_out, n_tiles=6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_tiles):
            if do_in:
                with torch.cuda.stream(s_in):
                    for h, d in zip(ins, dins):
                        d.copy_(h, non_blocking=True)
                    ddem[49:size + 51].copy_(dem[49:size + 51], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    for h, d in zip(outs, douts):
                        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / n_tiles

    res = {'world': world, 'h2d_bytes': h2d_bytes, 'd2h_bytes': d2h_bytes}
    for name, a, b, nbytes in (('h2d_only', True, False, h2d_bytes), ('d2h_only', False, True, d2h_bytes),
                               ('both', True, True, h2d_bytes + d2h_bytes)):
        copies(a, b, 2)
        barrier()
        ts = []
        for _ in range(5):
            ts.append(copies(a, b))
        barrier()
        mean, best = sum(ts) / len(ts), min(ts)
        res[name] = {'ms_per_tile_mean': reduce(mean, dist.ReduceOp.MAX), 'ms_per_tile_best': reduce(best, dist.ReduceOp.MAX),
                     'GBps_all_ranks_mean': reduce(nbytes / mean / 1e6, dist.ReduceOp.SUM)}

    # the product's pipeline
    t = synth.make_tile(rank, size, size)
    pin = dict(bands=[pb.pinned_copy(b) for b in t['bands']], fmask=pb.pinned_copy(t['fmask']), dem=pb.pinned_copy(t['dem']),
               land=pb.pinned_copy(t['land']), ocean=pb.pinned_copy(t['ocean']))
    outb = []
    for _ in range(2):
        o = {n: pb.pinned_empty((size, size), np.uint16 if n == 'DIAG' else np.uint8) for n in pb.GRADED_LAYERS}
        o['counters'] = pb.pinned_empty((12,), np.uint64)
        outb.append(o)
    params = pb.make_params(collapse_wtr_classes=True)
    pipe = pb.TilePipeline()

    def submit(i):
        return pipe.submit(pin['bands'], pin['fmask'], pin['dem'], pin['land'], pin['ocean'], t['sun_azimuth'],
                           t['sun_elevation'], params=params, outputs=pb.GRADED_LAYERS, out=outb[i & 1])
    for i in range(4):
        submit(i)
    pipe.flush()
    barrier()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        for i in range(12):
            submit(i)
        pipe.flush()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3 / 12)
    barrier()
    mean, best = sum(ts) / len(ts), min(ts)
    res['pipeline'] = {'ms_per_tile_mean': reduce(mean, dist.ReduceOp.MAX), 'ms_per_tile_best': reduce(best, dist.ReduceOp.MAX),
                       'ms_per_tile_mean_fastest_rank': reduce(mean, dist.ReduceOp.MIN),
                       'GBps_all_ranks_mean': reduce((h2d_bytes + d2h_bytes) / mean / 1e6, dist.ReduceOp.SUM)}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
