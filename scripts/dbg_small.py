import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proteus_b200 as pb
from proteus_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else 'nodem'
t = synth.make_tile(3, 64, 128)
if which == 'nodem':
    r = pb.classify_tile(t['bands'], t['fmask'], None, t['land'], t['ocean'])
elif which == 'dem_notma':
    d = np.ascontiguousarray(t['dem'][:, :-1])     # pitch % 4 != 0 -> generic loader
    r = pb.classify_tile(t['bands'], t['fmask'], d, t['land'], t['ocean'], 150., 45., dem_off=(50, 50))
else:
    r = pb.classify_tile(t['bands'], t['fmask'], t['dem'], t['land'], t['ocean'], 150., 45.)
print(which, 'ok', r['coverage'])
