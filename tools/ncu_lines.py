"""Per CUDA source line: warp instructions executed and stall samples of one kernel of an .ncu-rep (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py REPORT.ncu-rep kernel-regex [topN]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
agg = collections.OrderedDict()
fname, hdr = '?', None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ie, ss = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [r[1].strip(), 0, 0])
        try:
            a[1] += int(r[ie]); a[2] += int(r[ss])
        except ValueError:
            pass
tot_i = sum(v[1] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print(f"{kre}: {tot_i / 1e6:.1f} M warp instructions, {tot_s} stall samples")
for (f, ln), (src, i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"  {100 * i / tot_i:5.1f}% instr {100 * s / tot_s:5.1f}% stall  {f}:{ln:<4d} {src[:110]}")
