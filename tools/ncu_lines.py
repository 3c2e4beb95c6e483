"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on): executed warp instructions and stall
samples per CUDA source line. Usage: ncu_lines.py <report.ncu-rep> [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = None
agg = collections.OrderedDict()
hdr = None
cur_file = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0]); inst = int(r[7]); samp = int(r[6])
    except (ValueError, IndexError):
        continue
    k = (cur_file, line, r[1].strip()[:110])
    a = agg.setdefault(k, [0, 0]); a[0] += inst; a[1] += samp
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"{fn}\n total warp instructions {tot_i}  samples {tot_s}")
for (f, line, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * i / max(tot_i, 1):5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% smp  {f}:{line:<4d} {src}")
