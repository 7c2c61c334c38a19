#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel in an .ncu-rep: python scripts/ncu_lines.py rep kernel-regex [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, text=True).stdout
cur, rows = None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 7 and r[0].isdigit():
        try: rows.append((int(r[6]), cur, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(x[0] for x in rows) or 1
print("total samples", tot)
for s, f, l, src in sorted(rows, reverse=True)[:top]:
    print("%6d %5.1f%% %s:%d  %s" % (s, 100 * s / tot, f, l, src))
