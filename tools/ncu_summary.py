"""Summarise an .ncu-rep (raw page + cuda,sass source page) into text: key counters, per-function and per-line hot spots."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
H, U = r[0], r[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'smsp__sass_average_data_bytes_per_sector_mem_local', 'sass__inst_executed_local', 'gpu__dram_throughput.avg.pct',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct', 'smsp__inst_executed_op_branch', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__sass_thread_inst_executed_op_dadd', 'sm__sass_thread_inst_executed_op_dmul', 'sm__sass_thread_inst_executed_op_dfma', 'smsp__sass_thread_inst_executed_op_integer']
for row in r[2:]:
    print("== kernel:", row[H.index('Kernel Name')][:100])
    for h, u, v in zip(H, U, row):
        if any(w in h for w in want) and 'Not Issued' not in h:
            try:
                if float(v.replace(',', '')) == 0 and 'stalled' in h: continue
            except ValueError: pass
            print("  %-90s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0]); text = {}
for x in rows:
    if not x: continue
    if x[0] == 'File Path': cur = x[1].split('/')[-1]; continue
    if x[0] == 'Function Name': continue
    if x[0] == 'Line No': hdr = x; iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or x[0] == '': continue
    key = (cur, int(x[0])); text[key] = x[1]
    try: agg[key][0] += int(x[iI]); agg[key][1] += int(x[iT]); agg[key][2] += int(x[iS])
    except ValueError: pass
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[2] for v in agg.values()) or 1
print("== source lines: total warp-inst %.3e thread-inst %.3e (avg lanes %.2f) samples %d" % (ti, sum(v[1] for v in agg.values()), sum(v[1] for v in agg.values()) / ti, ts))
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top_n]:
    print("  %-16s %5d  inst %5.2f%%  lanes %5.2f  samples %5.2f%%  %s" % (f, l, 100 * v[0] / ti, v[1] / max(1, v[0]), 100 * v[2] / ts, text[(f, l)].strip()[:100]))
