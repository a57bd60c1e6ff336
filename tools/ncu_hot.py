#!/usr/bin/env python
"""Where a kernel's issue slots and stall samples go: python tools/ncu_hot.py file.ncu-rep  (SASS-level, grouped into
runs of instructions between control-flow / barrier instructions)."""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_ex = sum(int(r[iex] or 0) for r in rows[2:] if len(r) > iex)
tot_s = sum(int(r[ismp] or 0) for r in rows[2:] if len(r) > ismp)
print("instructions executed", tot_ex, "samples", tot_s)
seg, segs = {"ex": 0, "s": 0, "n": 0, "first": None, "ops": {}}, []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    src = r[isrc].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    if seg["first"] is None:
        seg["first"] = r[ia][-5:]
    seg["ex"] += int(r[iex] or 0); seg["s"] += int(r[ismp] or 0); seg["n"] += 1
    seg["ops"][op] = seg["ops"].get(op, 0) + int(r[iex] or 0)
    if op in ("BAR", "BRA", "EXIT", "BSYNC", "SYNCS", "WARPSYNC", "UTMALDG"):
        segs.append(seg)
        seg = {"ex": 0, "s": 0, "n": 0, "first": None, "ops": {}}
segs.append(seg)
for sg in segs:
    if sg["ex"] > 0.01 * tot_ex or sg["s"] > 0.01 * tot_s:
        top = sorted(sg["ops"].items(), key=lambda kv: -kv[1])[:5]
        print("@%s  %3d instr  exec %5.1f%%  samples %5.1f%%  %s" % (sg["first"], sg["n"], 100.0 * sg["ex"] / tot_ex, 100.0 * sg["s"] / max(1, tot_s),
                                                                 " ".join("%s:%.0f%%" % (k, 100.0 * v / max(1, sg["ex"])) for k, v in top)))
