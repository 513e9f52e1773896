"""Turn gpurun_out/{full,launches,bench}_<tag> into the tracked files under profiles/.
usage: ncu_to_profiles.py <tag> <round-label>   e.g. r1p r1"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, label = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rep = os.path.join(G, f"full_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'l1tex__t_sector_hit_rate.pct', 'launch__block_size', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']
metrics = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals) if h in want}
kernel = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "vsr_episode_kernel"
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
out = {
    "command": "python bench.py --steps 2 --warmup 3 --no-cpu-baseline  (plain run exited 0 first) ; "
               "ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1",
    "kernel": kernel, "metrics": metrics, "summary_text": summ.strip().split("\n"),
}
json.dump(out, open(os.path.join(P, f"{label}_ncu_full_summary.json"), "w"), indent=1)
def mb(k):
    m = metrics[k]; f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m["unit"]]
    return float(m["value"]) * f
bench = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().split("\n")[-1])
vs = bench["config"]["robots_per_gpu"] * bench["config"]["voxels_per_robot"] * bench["config"]["ticks"]
json.dump({"dram_bytes_per_launch": mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'),
           "source": f"profiles/{label}_ncu_full_summary.json (ncu --set full on the bench.py command, one launch of vsr_episode_kernel)",
           "algorithmic_bytes_per_launch": vs * bench["roofline"]["algorithmic_bytes_per_voxel_step"]},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)
# launch list: keep the csv rows only
lines = [l for l in open(os.path.join(G, f"launches_{tag}.csv")) if l.startswith('"')]
open(os.path.join(P, f"{label}_launches_bench.csv"), "w").writelines(lines)
json.dump(bench, open(os.path.join(P, f"{label}_bench_line.json"), "w"), indent=1)
tot = {}
for r in csv.DictReader(io.StringIO("".join(lines))):
    tot[r["Kernel Name"]] = tot.get(r["Kernel Name"], 0) + float(r["Metric Value"])
s = sum(tot.values())
print({k[:60]: round(100 * v / s, 2) for k, v in tot.items()})
print("dram bytes/launch", mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'), "time ms", metrics['gpu__time_duration.sum'])
