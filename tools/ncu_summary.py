"""Summarise an .ncu-rep: key raw metrics per kernel + top stall instructions (needs -lineinfo builds).
usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel-regex] [topN]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 18
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index('Kernel Name')
names = []
for r in rows[2:]:
    print('==== ', r[ki][:110])
    names.append(r[ki])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:82s} {r[i]:>16s} {units[i]}")
seen = set()
for nm in names:
    short = nm.split('(')[0].split('::')[-1].split('<')[0].replace('void ', '').strip()
    if short in seen or (kre and kre not in short):
        continue
    seen.add(short)
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + short],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rr) if r and r[0] == 'Address']
    if not hi:
        continue
    h = rr[hi[0]]
    body = [r for r in rr[hi[0] + 1:] if len(r) == len(h) and r[0].startswith('0x')]
    si, ai, ei = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
    tot = sum(int(r[ai]) for r in body) or 1
    print(f'---- top stall instructions of {short} (total samples {tot}, warp-instructions {sum(int(r[ei]) for r in body)})')
    for r in sorted(body, key=lambda r: -int(r[ai]))[:topn]:
        print(f"  {int(r[ai]):7d} {100 * int(r[ai]) / tot:5.1f}%  exec={int(r[ei]):9d}  {r[si].strip()[:100]}")
