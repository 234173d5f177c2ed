import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, proteus_b200 as pb
from proteus_b200 import synth
from proteus_b200.engine import classify_device_cover, Plan
tile = synth.make_device_batch(1, 3660, 3660, device='cuda', seed=7, n_distinct=1)[0]
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print('cover flow        %.3f ms' % t(lambda: classify_device_cover(tile, outputs=pb.GRADED_LAYERS)))
params = pb.make_params(mask_adjacent_to_cloud_mode='ignore', defer_snow=True, collapse_wtr_classes=False)
print('make_params       %.3f ms' % t(lambda: pb.make_params(mask_adjacent_to_cloud_mode='ignore', defer_snow=True, collapse_wtr_classes=False)))
layers = ['DIAG', 'WTR1', 'WTR1_REMAPPED', 'WTR2', 'CLOUD', 'SHAD']
print('Plan create+close %.3f ms' % t(lambda: Plan([tile], params, layers).close()))
plan = Plan([tile], params, layers)
print('plan.run (full)   %.3f ms' % t(lambda: plan.run()))
plan2 = Plan([tile], pb.make_params(), pb.GRADED_LAYERS)
print('plan.run (lean)   %.3f ms' % t(lambda: plan2.run()))
