"""Developer tool: instruction counts per kernel of the built library (cuobjdump -sass), the
evidence that the sweeps use TMA (UTMALDG / UTMASTG), mbarriers (SYNCS) and packed fp32
(FFMA2 / FADD2), and that shared memory is addressed with LDS / STS, not generic LD.E.
Writes profiles/r02_sass_counts.txt; runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "magellanmapper_b200", "libmmb200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_counts.txt")
COLS = ["UTMALDG", "UTMASTG", "FFMA2", "FADD2", "SYNCS", "FFMA", "LDS", "LD.E"]

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
archs = set()
kernel = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kernel = m.group(1)
        counts.setdefault(kernel, collections.Counter())
        continue
    m = re.match(r"arch = (\S+)", line.strip())
    if m:
        archs.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kernel:
        op = m.group(1)
        base = op.split(".")[0]
        counts[kernel][base] += 1
        if op.startswith("LD.E"):
            counts[kernel]["LD.E"] += 1
with open(OUT, "w") as f:
    f.write(f"cuobjdump -sass magellanmapper_b200/libmmb200.so (cubins: {', '.join(sorted(archs))}), "
            "instruction counts per kernel (tools/sass_counts.py).\n")
    f.write("UTMALDG / UTMASTG = TMA tensor loads / stores, SYNCS = mbarrier operations, FFMA2 / FADD2 "
            "= packed fp32, LD.E = generic loads (0 in every kernel that addresses shared memory).\n")
    w = max(len(k) for k in counts) + 2
    f.write("kernel (mangled)".ljust(w) + "".join(c.rjust(9) for c in COLS) + "\n")
    for k in sorted(counts):
        f.write(k.ljust(w) + "".join(str(counts[k][c]).rjust(9) for c in COLS) + "\n")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    f.write("total".ljust(w) + "".join(str(tot[c]).rjust(9) for c in COLS) + "\n")
print(OUT, len(counts), "kernels")
