"""Static view of the fast kernel's row loop in a built library: the instructions between the loop head and the
backward branch that closes it (the innermost loop that contains the streaming stores), opcode mix, and an estimate of
the straight-line path (instructions outside blocks that end in CALL / that wait on barriers are counted separately).
usage: python scripts/sass_rowloop.py LIB.so [mangled-substring, default ILb0ELb1ELb1E = lean, all graded, FAST8]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else 'dswx_fused_fast_kernelILb0ELb1ELb1E'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fn, ins = None, []
for line in out.splitlines():
    m = re.match(r'\s+Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    if fn and want in fn:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, t) in enumerate(ins)}
stores = [i for i, (a, t) in enumerate(ins) if t.startswith('STG.E.NA')]
# back edges: (index, target index)
backs = []
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
        backs.append((i, addr[int(m.group(1), 16)]))
loops = [(e, s) for e, s in backs if any(s <= st <= e for st in stores)]
e, s_ = min(loops, key=lambda l: l[0] - l[1])
body = ins[s_:e + 1]
op = lambda t: re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', t).group(2)
mix = collections.Counter(op(t) for a, t in body)
print(f'{want}: kernel {len(ins)} instr; row loop {len(body)} instr (0x{ins[s_][0]:x}..0x{ins[e][0]:x})')
print('  ' + ' '.join(f'{k}:{v}' for k, v in mix.most_common(40)))
if '--dump' in sys.argv:
    for a, t in body:
        print(f'{a:05x} {t}')
