"""profiles/rNN_traffic.json from an ncu raw-page CSV of the fused force kernel (tools/ncu_capture.sh) and the build identity
written beside it.  usage: python tools/make_traffic_json.py <raw.csv> <build.json> <out.json> [kick_drift numbers are carried
over from the previous round's file: that kernel did not change]"""
import csv
import json
import sys

raw, build, out = sys.argv[1:4]
r = list(csv.reader(open(raw)))
H, rows = r[0], r[2:]
col = lambda name: H.index(name)
f = lambda row, name: float(row[col(name)].replace(",", ""))
rows = sorted(rows, key=lambda x: f(x, "gpu__time_duration.sum"))
inner = [x for x in rows if f(x, "gpu__time_duration.sum") < 1.15 * f(rows[0], "gpu__time_duration.sum")]
refresh = [x for x in rows if x not in inner]
avg = lambda sel, name: sum(f(x, name) for x in sel) / max(len(sel), 1)
w_in, w_rf = 0.8, 0.2   # one inner-list refresh every ~5 steps
mix = lambda name: w_in * avg(inner, name) + (w_rf * avg(refresh, name) if refresh else w_rf * avg(inner, name))
unit = r[1][col("dram__bytes_read.sum")]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1.0}[unit]
prev = {}
try:
    prev = json.load(open("profiles/r01_traffic.json"))
except Exception:
    pass
doc = {
    "source": "%s (ncu --set full --clock-control none, N=2^24, %d inner-list launches + %d refresh launches)" % (raw, len(inner), len(refresh)),
    "k_force_list_fused": {
        "n_particles": 16777216,
        "dram_bytes_read": mix("dram__bytes_read.sum") * scale,
        "dram_bytes_write": mix("dram__bytes_write.sum") * scale,
        "duration_ms": mix("gpu__time_duration.sum"),
        "fp64_pipe_active_pct": mix("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": mix("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "inner_launch": {"ms": avg(inner, "gpu__time_duration.sum"), "dram_read": avg(inner, "dram__bytes_read.sum") * scale,
                         "dram_write": avg(inner, "dram__bytes_write.sum") * scale},
        "refresh_launch": {"ms": avg(refresh, "gpu__time_duration.sum"), "dram_read": avg(refresh, "dram__bytes_read.sum") * scale,
                           "dram_write": avg(refresh, "dram__bytes_write.sum") * scale} if refresh else None,
        # identity of the kernel build (not of a launch: the resident-CTA count depends on the handle)
        "build": {k: v for k, v in json.load(open(build)).items() if k != "ctas_per_sm"},
        "note": "weighted 0.8 inner / 0.2 refresh like a run",
    },
}
for k in ("k_kick_drift", "k_force_list"):
    if k in prev:
        doc[k] = prev[k]
        doc[k]["carried_over_from"] = "profiles/r01_traffic.json (kernel unchanged)"
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc["k_force_list_fused"], indent=1))
