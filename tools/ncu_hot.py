"""Top stall lines of one kernel from an ncu report's source page (run where ncu is installed).
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [min_percent]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [k for k, r in enumerate(rows) if r and r[0] == "Address"]
start = hi[0]
end = hi[1] - 1 if len(hi) > 1 else len(rows)     # first launch only
h = rows[start]
c = h.index("Warp Stall Sampling (All Samples)")
e = h.index("Instructions Executed")
body = [r for r in rows[start + 1:end] if len(r) > c and r[0].startswith("0x")]
tot = sum(float(r[c] or 0) for r in body)
print("instructions", len(body), "samples", tot)
stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and not x.endswith("_not_issued")]
for k, r in enumerate(body):
    f = float(r[c] or 0)
    if f / tot * 100 >= thr:
        why = sorted(((float(r[i] or 0), x) for i, x in stall_cols), reverse=True)[:3]
        why = " ".join(f"{x[6:]}={v:.0f}" for v, x in why if v > 0)
        print(f"{k:5d} {f / tot * 100:5.1f}%  exec {r[e]:>9}  {r[1].strip()[:80]:80s} {why}")
