"""Total warp-stall samples by reason for an .ncu-rep (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:])); h = rows[0]; rows = rows[1:]
cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = {h[i]: sum(int(r[i]) for r in rows) for i in cols}
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"{k:28s} {v:10d} {v/s*100:6.2f}%")
