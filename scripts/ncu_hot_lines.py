"""Top CUDA source lines of a kernel by warp-stall samples from an .ncu-rep captured with --import-source on:
   python scripts/ncu_hot_lines.py <report.ncu-rep> <kernel-name> [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
cur, hdr, lines = None, None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("Warp Stall Sampling (All Samples)")
        i_i = hdr.index("Instructions Executed")
        i_t = hdr.index("Thread Instructions Executed")
    elif hdr and len(r) == len(hdr) and r[0] != "":
        lines.append((cur, r[0], r[1], float(r[i_s] or 0), float(r[i_i] or 0), float(r[i_t] or 0)))
tot = sum(l[3] for l in lines) or 1.0
toti = sum(l[4] for l in lines) or 1.0
print(f"{kern}: {int(tot)} samples, {int(toti)} warp instructions")
for f, ln, src, s, ins, tin in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * s / tot:5.1f}% smp {100 * ins / toti:5.1f}% ins {tin / max(ins, 1):5.1f} thr  {f}:{ln}  {src.strip()[:100]}")
