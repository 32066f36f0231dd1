#!/bin/bash
# usage: tools/ncu_capture.sh <tag> [ENV=VALUE ...]   (run on the GPU box, under gpurun)
# ncu --set full capture of the list-mode pair-force kernel at N = 2^24 (bench.py --eager: ncu cannot profile kernel nodes of
# graphs with conditional nodes; eager mode launches the identical kernels).  The .ncu-rep is reduced ON THE BOX to the raw
# metric table (all captured launches) and the per-line source page of one inner-list launch, then removed: gpurun copies
# back at most 64 MiB.
tag=$1; shift
for kv in "$@"; do export "$kv"; done
B="python bench.py --eager --steps 8 --warmup 3 --melt 300 --no-e2e --no-cpu --no-profile --no-extra"
$B > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_force_list -s 303 -c 6 -o /tmp/${tag} $B > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/${tag}_source.csv 2>/dev/null
# identity of the build the capture belongs to (bench.py refuses ncu numbers of another build)
python - > gpurun_out/${tag}_build.json <<'PY'
import json, sys
sys.path.insert(0, ".")
import mdjl_b200 as md
e = md.Engine(3, 4096, 20.0, 1.5, 0)
print(json.dumps(e.force_kernel_info()))
PY
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}_raw.csv gpurun_out/${tag}_source.csv
rm -f /tmp/${tag}.ncu-rep
