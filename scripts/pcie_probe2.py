"""H2D of one HLS tile in the host pipeline's copy pattern (10 rasters x 5 strips), from pb200 pinned buffers and
from torch pinned buffers, with and without concurrent D2H of the 4 output layers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import proteus_b200 as pb
size = 3660
edges = [0, 1024, 2048, 3072, 3392, 3660]
def make(kind):
    def alloc(shape, dt):
        if kind == 'pb200':
            return torch.from_numpy(pb.pinned_empty(shape, dt))
        return torch.empty(shape, dtype={np.int16: torch.int16, np.uint8: torch.uint8, np.float32: torch.float32, np.uint16: torch.int16}[dt]).pin_memory()
    ins = [alloc((size, size), np.int16) for _ in range(6)] + [alloc((size, size), np.uint8) for _ in range(3)]
    dem = alloc((size + 100, size + 100), np.float32)
    outs = [alloc((size, size), np.uint16)] + [alloc((size, size), np.uint8) for _ in range(3)]
    return ins, dem, outs
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
for kind in ('torch', 'pb200'):
    ins, dem, outs = make(kind)
    for t in ins + [dem] + outs: t.zero_() if t.dtype != torch.float32 else t.fill_(1.0)
    dins = [torch.empty_like(t, device='cuda') for t in ins]; ddem = torch.empty_like(dem, device='cuda')
    douts = [torch.empty_like(t, device='cuda') for t in outs]
    def run(d2h):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        evs = []
        with torch.cuda.stream(s_in):
            for a, b in zip(edges[:-1], edges[1:]):
                for h, d in zip(ins, dins): d[a:b].copy_(h[a:b], non_blocking=True)
                ddem[a + 49:b + 51].copy_(dem[a + 49:b + 51], non_blocking=True)
                e = torch.cuda.Event(); e.record(s_in); evs.append(e)
        if d2h:
            with torch.cuda.stream(s_out):
                for (a, b), e in zip(zip(edges[:-1], edges[1:]), evs):
                    s_out.wait_event(e)
                    for h, d in zip(outs, douts): h[a:b].copy_(d[a:b], non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
    for _ in range(3): run(True)
    print(kind, 'H2D only %.3f ms' % min(run(False) for _ in range(5)), ' H2D + D2H %.3f ms' % min(run(True) for _ in range(5)))
