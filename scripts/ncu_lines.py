"""Summarise an ncu report: headline metrics + per-source-line instruction / stall-sample shares."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
for h, v in zip(hdr, vals):
    if h in want: print(f'{h:90s} {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur = None; out = []; tot = 0; totsamp = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name') or r[0] == '': continue
    try: inst = int(r[7]); samp = int(r[4])
    except ValueError: continue
    out.append((inst, samp, cur, r[0], r[1].strip()[:100])); tot += inst; totsamp += samp
print('total inst', tot, 'samples', totsamp)
for inst, samp, f, ln, s in sorted(out, reverse=True)[:top]:
    print(f'{inst/tot*100:5.1f}% inst {samp/max(totsamp,1)*100:5.1f}% samp {f}:{ln} {s}')
