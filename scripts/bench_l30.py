"""BASELINE configs[0] on the GPU: L30-style tiles without DEM / LAND / ocean, outputs DIAG + WTR (13 B in + 3 B out =
16 algorithmic bytes per pixel), device resident, CUDA events.  One JSON object on stdout."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import proteus_b200 as pb
from proteus_b200 import synth

n_tiles, size, steps = 16, 3660, 60
tiles = synth.make_device_batch(n_tiles, size, size, device='cuda', seed=1000, n_distinct=4, full_product=False)
plan = pb.Plan(tiles, pb.make_params(collapse_wtr_classes=True), ('DIAG', 'WTR'))
st = torch.cuda.current_stream()
for _ in range(5):
    plan.run(st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(st)
for _ in range(steps):
    plan.run(st)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
px = n_tiles * size * size
peak = 6550.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
gbs = px * 16 / ms / 1e6
print(json.dumps({'workload': 'configs[0]: 16 synthetic L30-style tiles 3660x3660, bands + Fmask -> DIAG + WTR, device resident',
                  'ms_per_launch': round(ms, 4), 'Gpixel_per_s': round(px / ms / 1e6, 1), 'algorithmic_bytes_per_pixel': 16,
                  'achieved_GBps': round(gbs, 1), 'hbm_peak_GBps': peak, 'frac': round(gbs / peak, 3)}))
