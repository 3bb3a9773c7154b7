"""Prints the handful of ncu metrics the kernel notes in DESIGN.md quote, from a .ncu-rep:
python benchmarks/ncu_key.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'smsp__inst_executed_op_shared_ld.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for k, v in zip(hdr, r):
        if k in WANT or 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(v or 0) > 0.3:
            print('  %-95s %s %s' % (k, v, units[hdr.index(k)]))
