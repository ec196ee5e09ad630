"""Per-kernel SASS opcode counts of libcnerf.so (evidence that the hot kernels are Blackwell-native: UTCHMMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier).

    python scripts/sass_counts.py > profiles/r2_sass_counts.txt        (cuobjdump only: runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "consistentnerf_b200", "libcnerf.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["total"] += 1
        for o in OPS:
            if op.startswith(o):
                cur[o] += 1
print("# cuobjdump -sass consistentnerf_b200/libcnerf.so (sm_100a): opcode counts per kernel")
print(f"{'kernel':78s} " + " ".join(f"{o:>8s}" for o in ["total"] + OPS))
for name, c in counts.items():
    if c["total"]:
        print(f"{name[:78]:78s} " + " ".join(f"{c[o]:8d}" for o in ["total"] + OPS))
