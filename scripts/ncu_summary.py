"""Summarise an .ncu-rep: key metrics, stall reasons, opcode mix, hot regions."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; warp_ticks = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want: print("%-70s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, sc, ie, at = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed'), h.index('Avg. Threads Executed')
stalls = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[2:]:
    if len(r) <= si: continue
    try: data.append((int(r[si]), int(r[ie]), float(r[at] or 0), r))
    except Exception: pass
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print("SASS instructions", len(data), "executed", toti, ("per warp-tick %.0f" % (toti / warp_ticks)) if warp_ticks else "")
agg = collections.Counter()
for s, i, a, r in data:
    for k, n in stalls: agg[n] += int(r[k] or 0)
print("stalls %:", [(n[6:], round(100 * v / tot, 1)) for n, v in agg.most_common(8)])
byop = collections.defaultdict(lambda: [0, 0])
for s, i, a, r in data:
    t = r[sc].strip().split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    byop[op][0] += s; byop[op][1] += i
print("opcode mix:", " ".join("%s %.1f%%/%.1f%%" % (op, 100 * i / toti, 100 * s / tot) for op, (s, i) in sorted(byop.items(), key=lambda x: -x[1][1])[:14]))
chunk = 300
for c in range(0, len(data), chunk):
    ss = sum(d[0] for d in data[c:c + chunk]); ii = sum(d[1] for d in data[c:c + chunk])
    ath = sum(d[2] * d[1] for d in data[c:c + chunk]) / max(1, ii)
    if ii * 200 > toti or ss * 100 > tot:
        print("  instr %5d-%5d samples %5.1f%% inst %5.1f%% %s avgthr %.1f" % (c, c + chunk, 100 * ss / tot, 100 * ii / toti, ("(%.0f/warp-tick)" % (ii / warp_ticks)) if warp_ticks else "", ath))
