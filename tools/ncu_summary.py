"""Summarise an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py gpurun_out/x.ncu-rep [more metrics...]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d['Kernel Name'][:90])
        for k in KEYS + extra:
            if k in d:
                print(f'  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}')
        st = []
        for k, v in d.items():
            if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
                try:
                    st.append((float(v.replace(',', '')), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        print('  stalls per issue:', ', '.join(f'{k}={v:.2f}' for v, k in sorted(st, reverse=True)[:8]))


if __name__ == '__main__':
    main()
