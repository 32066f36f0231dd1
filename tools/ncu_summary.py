"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (run here, no GPU needed).
usage: python tools/ncu_summary.py <tag> <launches.csv> [<report.ncu-rep> ...]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
out = []
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    name = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("mdb::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.append("## launch list (%s): gpu__time_duration.sum per kernel, cold-cache serialised -- compare shares\n" % sys.argv[2])
out.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.1f | %.1f | %.3f |" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for rep in sys.argv[3:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    Hh, U = r[0], r[1]
    kn = Hh.index("Kernel Name")
    for row in r[2:]:
        out.append("\n## ncu --set full: %s  (%s)\n" % (row[kn].split("(")[0], rep))
        out.append("| metric | unit | value |\n|---|---|---|")
        for i, h in enumerate(Hh):
            if h in KEYS:
                out.append("| %s | %s | %s |" % (h, U[i], row[i]))
open("profiles/%s.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
