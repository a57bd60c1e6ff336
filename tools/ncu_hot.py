"""Print the hottest SASS instructions (by stall samples) of a kernel from an .ncu-rep, with
the CUDA source line each belongs to.  Usage: python tools/ncu_hot.py REP [kernel_index] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = heads[kidx]; end = heads[kidx + 1] if kidx + 1 < len(heads) else len(rows)
H = rows[start]
si = H.index("# Samples"); src = H.index("Source")
stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[start + 1:end] if len(r) == len(H)]
tot = sum(int(r[si] or 0) for r in body)
print("kernel block", kidx, "instructions", len(body), "samples", tot)
agg = {}
for i, h in stall_cols:
    agg[h] = sum(int(r[i] or 0) for r in body)
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top]:
    st = {h[6:]: int(r[i]) for i, h in stall_cols if r[i] and int(r[i])}
    print("%6s %5.1f%%  %-70s %s" % (r[si], 100.0 * int(r[si] or 0) / max(tot, 1), r[src][:70], st))
