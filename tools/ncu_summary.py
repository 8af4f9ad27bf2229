"""Summarise an .ncu-rep (read on the CPU box): key raw metrics, stall reasons, opcode mix.
python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep [--sass]"""
import collections, csv, io, re, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'local_load_bytes', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-70s %s %s' % (k, r[i], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
                st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        print('  stalls/issue:', ', '.join('%s=%.2f' % (n, v) for v, n in sorted(st, reverse=True)[:9]))
    if '--sass' in sys.argv:
        rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
        h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
        ix = {n: j for j, n in enumerate(rows[h])}
        op, samp, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[h + 1:]:
            if not r or not r[0].startswith('0x'):
                if r and r[0] == 'Kernel Name':
                    break
                continue
            n, s = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
            o = m.group(2) if m else '?'
            o = o if o.startswith('MUFU') else o.split('.')[0]
            op[o] += n; samp[o] += s; tot += n
        print('total warp instructions', tot)
        for o, n in op.most_common(28):
            print('  %-12s %12d %.3f samples %d' % (o, n, n / tot, samp[o]))


main()
