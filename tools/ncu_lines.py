"""Aggregate the warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[2] != "-":          # per-line rows have '-' as address; per-instruction rows follow
        continue
    d = dict(zip(hdr[4:], r[4:]))
    samples = int(d["# Samples"] or 0)
    if samples == 0:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    out.append((samples, cur_file, int(r[0]), int(d["Instructions Executed"] or 0), stalls, r[1]))
tot = sum(o[0] for o in out)
out.sort(reverse=True)
print("total samples", tot)
for s, f, ln, ie, st, src in out[:top]:
    ss = " ".join("%s=%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %s:%d inst=%d [%s] %s" % (100.0 * s / tot, f, ln, ie, ss, src.strip()[:90]))
