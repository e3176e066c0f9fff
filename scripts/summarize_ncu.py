"""Turn .ncu-rep captures and launch-list CSVs in gpurun_out/ into small text summaries under profiles/ (helper)."""
import csv, subprocess, sys, os, io
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']

def raw_summary(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f"# ncu --set full --clock-control none, raw-page excerpt of {os.path.basename(rep)}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n")
            for h in hdr:
                if h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h):
                    i = hdr.index(h)
                    f.write(f"{h} [{units[i]}] = {r[i]}\n")

def launch_summary(csvpath, out):
    rows = [r for r in csv.reader(open(csvpath)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    agg = {}
    order = []
    for r in rows[1:]:
        if r[im] != 'gpu__time_duration.sum': continue
        k = r[ik][:90]
        if k not in agg: agg[k] = [0, 0.0]; order.append(k)
        agg[k][0] += 1; agg[k][1] += float(r[iv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list of {os.path.basename(csvpath)} (cold-cache, serialised: compare shares)\n")
        f.write(f"# total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
        for k in sorted(order, key=lambda k: -agg[k][1]):
            f.write(f"{agg[k][1]/1e3:12.1f} us  {100*agg[k][1]/tot:6.2f}%  x{agg[k][0]:<4d} {k}\n")

if __name__ == '__main__':
    kind, src, dst = sys.argv[1:4]
    (raw_summary if kind == 'raw' else launch_summary)(src, dst)
