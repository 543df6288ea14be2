"""Summarise an .ncu-rep: key metrics + instruction share / stall samples per source line."""
import csv, os, shlex, subprocess, sys
EXTRA = shlex.split(os.environ.get("NCU_EXTRA", ""))  # e.g. NCU_EXTRA="-k regex:parse_kernel_v4"
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, *EXTRA, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %s %s" % (w, vals[i], units[i]))
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        try:
            if float(vals[i]) > 2:
                print("%-70s %s" % (h, vals[i]))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, *EXTRA, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hd, items = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hd = r; continue
    if hd and len(r) == len(hd) and r[2] == '-':
        try:
            n = int(r[hd.index('Instructions Executed')]); s = int(r[hd.index('# Samples')]); t = int(r[hd.index('Thread Instructions Executed')])
        except ValueError:
            continue
        items.append((n, s, cur, int(r[0]), r[1], t))
tot = sum(i[0] for i in items); tots = sum(i[1] for i in items)
print("total warp-inst %d, samples %d" % (tot, tots))
for n, s, f, ln, sl, t in sorted(items, key=lambda x: -(x[0] / tot + x[1] / tots))[:topn]:
    print("%5.1f%% inst %5.1f%% samp thr %4.1f %s:%d: %s" % (100 * n / tot, 100 * s / tots, t / max(n, 1), f, ln, sl.strip()[:95]))
