#!/bin/bash
# usage: tools/ncu_summary.sh report.ncu-rep > profiles/rNN_ncu_<kernel>.txt   -- the headline numbers of an `ncu --set full` capture
REP=$1
echo "# $(basename $REP): ncu --set full --clock-control none (cold-cache, serialised; see tools/gpu/*.sh for the command)"
ncu -i $REP --page details 2>/dev/null | grep -E "^  [a-zA-Z_:0-9<>(), ]+\(|Duration|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc|Issue Slots Busy|No Eligible|Shared Memory Config|Dynamic Shared|Block Limit|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Compute \(SM\) Throughput|Memory Throughput|DRAM Throughput|Grid Size|Block Size|Waves Per SM|excessive|bank conflicts|Local" | sed 's/  */ /g'
echo "# raw counters"
ncu -i $REP --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:100], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for w in want:
        if w in hdr: print(f'  {w:95s} {rows[1][hdr.index(w)]:8s} {r[hdr.index(w)]}')
"
