"""Summarise an ncu report: headline metrics + executed thread-instructions per pixel per source line.
usage: python scripts/ncu_lines.py REPORT.ncu-rep N_PIXELS [top]"""
import csv, collections, subprocess, sys, io
rep, npx = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed.sum', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'launch__waves_per_multiprocessor']
d = dict(zip(hdr, zip(units, vals)))
for w in want:
    if w in d: print(f'{w:75s} {d[w][1]:>16s} {d[w][0]}')
for h in hdr:
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
        v = float(d[h][1])
        if v > 3: print(f'{h:75s} {v:16.2f} %')
if 'smsp__inst_executed.sum' in d:
    print('thread-instr per pixel (warp-instr*32/px):', float(d['smsp__inst_executed.sum'][1]) * 32 / npx)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = {'file': r[1], 'rows': []}; sections.append(cur); continue
    if r and r[0] == 'Line No': cur['hdr'] = r; continue
    if cur is not None and 'hdr' in cur and len(r) == len(cur['hdr']): cur['rows'].append(r)
for s in sections:
    h = s['hdr']; iL = h.index('Line No'); iE = h.index('Instructions Executed')
    per = collections.OrderedDict()
    for r in s['rows']:
        if not r[iL].strip(): continue
        try: e = int(r[iE])
        except ValueError: continue
        key = (int(r[iL]), r[1][:100].strip())
        per[key] = per.get(key, 0) + e
    tot = sum(per.values()) * 32 / npx
    if tot < 0.5: continue
    print(f'== {s["file"]}  total {tot:.1f} thread-instr/px')
    for (ln, text), e in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
        print(f'{e * 32 / npx:7.2f}  L{ln}: {text}')

# ---- executed instruction mix by opcode (SASS view) ----
import re
mix = collections.Counter()
for s in sections:
    h = s['hdr']; iE = h.index('Instructions Executed'); iA = h.index('Address') if 'Address' in h else None
    for r in s['rows']:
        if r[h.index('Line No')].strip(): continue      # keep only SASS rows (no line number)
        try: e = int(r[iE])
        except ValueError: continue
        txt = r[1].strip() if iA is None else r[h.index('Source', 2)].strip() if False else r[3].strip()
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', txt)
        if m: mix[m.group(2)] += e
    break
tot = sum(mix.values())
print('== executed opcode mix (thread-instr/px), total %.1f' % (tot * 32 / npx))
for op, e in mix.most_common(45):
    print(f'{e * 32 / npx:7.2f}  {op}')
