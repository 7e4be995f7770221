"""Hot SASS of an .ncu-rep source page: lines above a share of executed instructions, in address order."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
h = rows[0]; rows = rows[1:]
ie, sm, src = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
at = h.index("Avg. Threads Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ie]) for r in rows); tots = sum(int(r[sm]) for r in rows)
print(f"total warp-instr {tot:,}  samples {tots:,}  sass lines {len(rows)}")
for k, r in enumerate(rows):
    if int(r[ie]) >= thr * tot:
        st = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if int(r[i])), reverse=True)[:2]
        print(f"{k:5d} {int(r[ie])/tot*100:6.2f}% smp {int(r[sm])/max(tots,1)*100:6.2f}% thr {float(r[at]):5.1f}  {r[src].strip():60s} {st}")
