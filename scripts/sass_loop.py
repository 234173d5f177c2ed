"""Static size of the fast kernel's row loop: instructions between the first streaming band load and the
backward branch that closes the row loop (straight-line count incl. rarely taken blocks placed inside).
usage: python scripts/sass_loop.py LIB.so"""
import re, subprocess, sys, collections
out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
fn = None; ins = []
for line in out.splitlines():
    m = re.match(r'\s+Function : (\S+)', line)
    if m: fn = m.group(1); continue
    if fn and 'dswx_fused_fast_kernelILb0' in fn:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
first = next(i for i, (a, t) in enumerate(ins) if 'LDG.E.NA.64' in t or 'LDG.E.64.NA' in t)
# the row loop's back edge: the last backward branch whose target is at or before the first band load and after the item loop head
back = [i for i, (a, t) in enumerate(ins) if re.search(r'BRA(\.U)?\s', t) and (m := re.search(r'0x([0-9a-f]+)', t)) and int(m.group(1), 16) <= ins[first][0] and i > first]
cands = [(i, int(re.search(r'0x([0-9a-f]+)', ins[i][1]).group(1), 16)) for i in back]
i_end, tgt = max(cands, key=lambda c: c[1])      # innermost loop: the highest target address
start = next(i for i, (a, t) in enumerate(ins) if a >= tgt)
body = ins[start:i_end + 1]
mix = collections.Counter(re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', t).group(2) for a, t in body)
print(f'{sys.argv[1]}: kernel {len(ins)} instr; row loop {len(body)} instr (0x{ins[start][0]:x}..0x{ins[i_end][0]:x})')
print('  ' + ' '.join(f'{k}:{v}' for k, v in mix.most_common(24)))
