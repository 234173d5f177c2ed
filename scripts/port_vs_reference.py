"""How much faster per core is the numpy port (oracle/dswx_oracle.py) than the reference's own functions?

Build container only (/root/reference must exist).  Runs the LIVE unmodified reference chain in
generate_dswx_layers' statement order (oracle/make_golden.py:reference_chain) and the port on the same rows of
synthetic tile 0, one process, best of 3 each, checks that the layers agree, and writes
profiles/r2_port_vs_reference.json - the factor bench.py quotes beside the `--impl reference` line (the GPU box has
no /root/reference, so the reference arm there times the port).

    python scripts/port_vs_reference.py [rows]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dswx_oracle as O, make_golden, ref_import      # noqa: E402
from proteus_b200 import synth                                     # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 732
size, m = 3660, 50
t = synth.make_tile(0, size, size)
sub = dict(t)
sub['bands'] = [np.ascontiguousarray(b[:rows]) for b in t['bands']]
for k in ('fmask', 'land', 'ocean'):
    sub[k] = np.ascontiguousarray(t[k][:rows])
sub['dem'] = np.ascontiguousarray(t['dem'][:rows + 2 * m])
ref = ref_import.load()
groups = ref_import.default_runconfig_groups()
th = ref.HlsThresholds()
for k, v in groups['hls_thresholds'].items():
    setattr(th, k, v)


def best(fn, n=3):
    out, times = None, []
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
    return out, min(times)


r_out, r_s = best(lambda: make_golden.reference_chain(ref, sub, groups['processing'], th, 'mask', True))
p_out, p_s = best(lambda: O.reference_chain(sub['bands'], sub['fmask'], sub['dem'], sub['land'], sub['ocean'],
                                            sub['sun_azimuth'], sub['sun_elevation']))
for name in ('DIAG', 'WTR1', 'WTR2', 'CLOUD', 'SHAD', 'WTR', 'BWTR', 'CONF', 'WTR_COLLAPSED'):
    assert np.array_equal(r_out[name], p_out[name]), name
mpx = rows * size / 1e6
res = dict(rows=rows, width=size, mpixel=mpx, reference_s=r_s, port_s=p_s,
           reference_mpixel_per_s_per_core=mpx / r_s, port_mpixel_per_s_per_core=mpx / p_s,
           port_vs_reference_per_core=r_s / p_s, numpy=np.__version__, cpu_count=os.cpu_count(),
           what='live unmodified proteus.dswx_hls functions in the order of dswx_hls.py:5088-5369 vs '
                'oracle.dswx_oracle.reference_chain, same rows of synthetic tile 0, one process, best of 3; '
                'all layers equal')
with open(os.path.join(ROOT, 'profiles', 'r2_port_vs_reference.json'), 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
