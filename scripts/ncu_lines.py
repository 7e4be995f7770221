"""Per CUDA-source-line share of executed warp instructions / samples (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
fname, hdr, items = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and "Instructions Executed" in r: hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No") and r[0].isdigit():
        items.append((fname, dict(zip(hdr, r)), r))
ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); at = hdr.index("Avg. Threads Executed")
tot = sum(int(r[ie]) for _, _, r in items); tots = sum(int(r[sm]) for _, _, r in items)
print(f"files lines {len(items)}  warp-instr {tot:,}  samples {tots:,}")
for f, d, r in items:
    if int(r[ie]) >= thr * tot or int(r[sm]) >= thr * tots:
        ti = hdr.index("Thread Instructions Executed")
        thr_avg = int(r[ti]) / max(int(r[ie]), 1)
        print(f"{f:16s}:{r[0]:>4s} inst {int(r[ie])/tot*100:6.2f}%  smp {int(r[sm])/tots*100:6.2f}%  thr {thr_avg:5.1f}  {r[1].strip()[:110]}")
