import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import proteus_b200 as pb
from proteus_b200 import synth
size = 3660
rng = np.random.default_rng(0)
bands = [pb.pinned_copy(rng.integers(1, 5000, (size, size), dtype=np.int16)) for _ in range(6)]
fmask = pb.pinned_copy(rng.integers(0, 255, (size, size), dtype=np.uint8))
land = pb.pinned_copy(rng.integers(0, 255, (size, size), dtype=np.uint8))
ocean = pb.pinned_copy(np.ones((size, size), np.uint8))
dem = pb.pinned_copy(rng.normal(size=(size + 100, size + 100)).astype(np.float32))
out = {n: pb.pinned_empty((size, size), np.uint16 if n == 'DIAG' else np.uint8) for n in pb.GRADED_LAYERS}
out['counters'] = pb.pinned_empty((12,), np.uint64)
params = pb.make_params()
# raw PCIe reference with torch pinned memory
tp = torch.empty(size * size, dtype=torch.int16).pin_memory(); td = torch.empty(size * size, dtype=torch.int16, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): td.copy_(tp, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('torch pinned H2D GB/s', 10 * tp.numel() * 2 / dt / 1e9)
tb = torch.from_numpy(bands[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): td.copy_(tb.view(-1), non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('pb200 pinned H2D via torch GB/s', 10 * tp.numel() * 2 / dt / 1e9, 'is_pinned', tb.is_pinned())
for strip in (3680, 1024, 512, 256, 128):
    for _ in range(2):
        pb.classify_tile(bands, fmask, dem, land, ocean, 150., 45., params=params, outputs=pb.GRADED_LAYERS, out=out, strip_rows=strip)
    t0 = time.perf_counter()
    for _ in range(5):
        pb.classify_tile(bands, fmask, dem, land, ocean, 150., 45., params=params, outputs=pb.GRADED_LAYERS, out=out, strip_rows=strip)
    dt = (time.perf_counter() - t0) / 5
    print(f'strip_rows {strip}: {dt*1e3:.2f} ms/tile  {(256e6+67e6)/dt/1e9:.1f} GB/s total')
