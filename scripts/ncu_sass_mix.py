"""Executed opcode mix and stall samples of the profiled kernel from the SASS page of an ncu report.
usage: python scripts/ncu_sass_mix.py REPORT.ncu-rep N_PIXELS [dump.txt]"""
import csv, collections, subprocess, sys, io, re
rep, npx = sys.argv[1], float(sys.argv[2])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
mix, stall = collections.Counter(), collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) != len(hdr): continue
    txt = r[iS].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', txt)
    if not m: continue
    e, n = int(r[iE]), int(r[iN])
    mix[m.group(2)] += e; stall[m.group(2)] += n
    lines.append((e, n, txt))
tot, tots = sum(mix.values()), sum(stall.values())
print(f'total warp-instr {tot}, thread-instr/px {tot * 32 / npx:.1f}; stall samples {tots}')
for op, e in mix.most_common(40):
    print(f'{e * 32 / npx:7.2f} /px  {100 * stall[op] / tots:5.1f}% samples  {op}')
if len(sys.argv) > 3:
    with open(sys.argv[3], 'w') as f:
        for e, n, txt in lines: f.write(f'{e * 32 / npx:7.3f} {n:6d}  {txt}\n')
