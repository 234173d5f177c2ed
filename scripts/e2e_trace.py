import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import proteus_b200 as pb
from proteus_b200 import synth
size = 3660
t = synth.make_tile(0, size, size)
pin = dict(bands=[pb.pinned_copy(b) for b in t['bands']], fmask=pb.pinned_copy(t['fmask']), dem=pb.pinned_copy(t['dem']),
           land=pb.pinned_copy(t['land']), ocean=pb.pinned_copy(t['ocean']))
out = {n: pb.pinned_empty((size, size), np.uint16 if n == 'DIAG' else np.uint8) for n in pb.GRADED_LAYERS}
out['counters'] = pb.pinned_empty((12,), np.uint64)
params = pb.make_params(collapse_wtr_classes=True)
strip = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def step():
    return pb.classify_tile(pin['bands'], pin['fmask'], pin['dem'], pin['land'], pin['ocean'], t['sun_azimuth'], t['sun_elevation'],
                            params=params, outputs=pb.GRADED_LAYERS, out=out, **({'strip_rows': strip} if strip else {}))
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(6): step()
print(f'python-level: {(time.perf_counter() - t0) / 6 * 1e3:.3f} ms per tile (strip_rows={strip or "default"})', file=sys.stderr)
