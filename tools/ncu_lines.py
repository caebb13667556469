"""Per-source-line totals (warp instructions executed, stall samples) of one kernel in an .ncu-rep.
usage: python tools/ncu_lines.py REPORT.ncu-rep kernel-regex [topN]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(r for r in rows if r and r[0] == 'Line No')
si, ei, ti = h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
lines = [(int(r[0]), r[1], int(r[si]), int(r[ei]), int(r[ti])) for r in rows if r and r[0].isdigit() and r[si].isdigit()]
tots, tote = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
print(f"total samples {tots}, warp instructions {tote}")
print("by line number:")
for ln, text, s, e, t in sorted(lines):
    if s * 200 >= tots or e * 200 >= tote:
        print(f"  L{ln:4d} samples {100*s/tots:5.1f}%  inst {100*e/tote:5.1f}% ({e:9d}, {t/max(e,1):4.1f} thr)  {text.strip()[:90]}")
