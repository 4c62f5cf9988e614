"""Per-source-line instruction counts and stall samples from an ncu report (compiled with -lineinfo, captured with
--import-source on):  python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
cur, out, hdr = None, [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":      # a source line row (SASS rows carry an address)
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            out.append((cur, int(r[0]), r[1].strip(), int(r[ie]), int(r[ss])))
        except ValueError:
            pass
ti, ts = sum(o[3] for o in out), sum(o[4] for o in out)
print(f"total warp instructions {ti}, stall samples {ts}")
for f, ln, src, n, s in out:
    if n >= ti * min_pct / 100 or s >= ts * min_pct / 100:
        print(f"{n / ti * 100:5.1f}% inst {s / ts * 100:5.1f}% smp  {f}:{ln:<4} {src[:100]}")
