"""Per CUDA source line: stall samples (barrier / long scoreboard / wait) and local-memory instructions.
usage: ncu_lines.py <rep> [warp_ticks]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; wt = float(sys.argv[2]) if len(sys.argv) > 2 else 1370 * 1200
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[2]
si, ie = h.index('# Samples'), h.index('Instructions Executed')
cols = {n: h.index(n) for n in ['stall_barrier', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_no_inst']}
line, txt, cur, tot, toti = {}, {}, None, 0, 0
loc = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[3:]:
    if len(r) != len(h): continue
    if r[0].strip().isdigit():
        cur = int(r[0]); txt[cur] = r[1]
        a = line.setdefault(cur, collections.Counter())
        a['s'] += int(r[si] or 0); a['i'] += int(r[ie] or 0)
        for n, k in cols.items(): a[n] += int(r[k] or 0)
    elif r[2].startswith('0x'):
        try: s, i = int(r[si] or 0), int(r[ie] or 0)
        except ValueError: continue
        tot += s; toti += i
        if 'LDL' in r[3] or 'STL' in r[3]:
            loc[cur][0] += i; loc[cur][1] += 'LDL' in r[3]; loc[cur][2] += 'STL' in r[3]
print("samples", tot, "inst", toti, "per warp-tick %.0f" % (toti / wt))
print("stall totals %:", {n[6:]: round(100 * sum(a[n] for a in line.values()) / tot, 1) for n in cols})
print("--- top lines by samples")
for l, a in sorted(line.items(), key=lambda x: -x[1]['s'])[:30]:
    print("%5d %5.2f%% inst %5.2f%% bar %4.1f lsb %4.1f wait %4.1f ssb %4.1f | %s" % (
        l, 100 * a['s'] / tot, 100 * a['i'] / toti, 100 * a['stall_barrier'] / tot, 100 * a['stall_long_sb'] / tot,
        100 * a['stall_wait'] / tot, 100 * a['stall_short_sb'] / tot, txt[l].strip()[:80]))
print("--- top lines by long scoreboard")
for l, a in sorted(line.items(), key=lambda x: -x[1]['stall_long_sb'])[:25]:
    print("%5d %5.2f%% | %s" % (l, 100 * a['stall_long_sb'] / tot, txt[l].strip()[:90]))
T = sum(a[0] for a in loc.values())
print("--- local memory instructions: %.1f per warp-tick" % (T / wt))
for l, a in sorted(loc.items(), key=lambda x: -x[1][0])[:30]:
    print("%5d %5.1f/warp-tick nLDL %d nSTL %d | %s" % (l, a[0] / wt, a[1], a[2], txt[l].strip()[:80]))
if len(sys.argv) > 3:  # phase table: comma-separated "start:name" boundaries of the kernel body
    bounds = [(int(b.split(':')[0]), b.split(':')[1]) for b in sys.argv[3].split(',')]
    print("--- phases (by kernel-body line; helper functions are listed by their own lines)")
    for (a, n), (b, _) in zip(bounds, bounds[1:] + [(10**9, '')]):
        ss = sum(v['s'] for l, v in line.items() if a <= l < b); ii = sum(v['i'] for l, v in line.items() if a <= l < b)
        bb = sum(v['stall_barrier'] for l, v in line.items() if a <= l < b)
        print("%-34s samples %5.1f%% (barrier %4.1f) inst %5.1f%%" % (n, 100 * ss / tot, 100 * bb / tot, 100 * ii / toti))
