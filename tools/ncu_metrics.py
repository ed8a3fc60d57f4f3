"""Prints the headline metrics of every launch in an .ncu-rep (raw page): duration, tensor-pipe %, DRAM bytes, L2, occupancy."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__cycles_active.avg', 'gpc__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform.sum',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('===', r[hdr.index('Kernel Name')][:100])
    for h, u, v in zip(hdr, units, r):
        if h in WANT:
            print(f'  {h:88s} {u:10s} {v}')
