"""Print the details page of an .ncu-rep as aligned text (optionally filtered by substrings)."""
import csv, subprocess, sys
rep, keep = sys.argv[1], sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
i_sec, i_m, i_u, i_v = h.index('Section Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
for row in r[1:]:
    if not row[i_m]: continue
    if not keep or any(k in row[i_m] for k in keep):
        print(row[i_sec][:30].ljust(30), row[i_m][:45].ljust(45), row[i_u][:12].ljust(12), row[i_v])
