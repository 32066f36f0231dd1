"""Static SASS instruction mix of one kernel of libmdb200.so (runs here, no GPU): what the compiler emitted for the hot
loops before any GPU time is spent.  usage: python tools/sass_mix.py <mangled-name-substring> [lib]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1] if len(sys.argv) > 1 else "k_force_listILi3ENS_6PotPHSELi2ELb0ELb0"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "moleculardynamics.jl_b200", "csrc", "libmdb200.so")
out = subprocess.check_output(["cuobjdump", "-sass", lib]).decode()
name, inside = None, False
ops, full = collections.Counter(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        inside = pat in m.group(1)
        if inside:
            name = m.group(1)
        continue
    if inside:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            full[m.group(1)] += 1
            ops[m.group(1).split(".")[0]] += 1
print("kernel:", name)
print("instructions:", sum(ops.values()))
for k, v in ops.most_common(24):
    print("%6d %s" % (v, k))
print("memory instructions by width:")
for k, v in sorted(full.items()):
    if k.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED")):
        print("%6d %s" % (v, k))
