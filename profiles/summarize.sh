#!/bin/bash
# Summarise an ncu report into profiles/<name>.txt:  profiles/summarize.sh gpurun_out/prof.ncu-rep r01_step_fast_late
set -e
REP=$1; NAME=$2; OUT=profiles/$NAME.txt
{
  echo "# ncu summary of $REP ($(date -u +%F)) -- ncu --set full --clock-control none --import-source on"
  ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__block_size','launch__grid_size','launch__shared_mem_per_block_dynamic','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warps_issue_stalled','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','derived__memory_l1_wavefronts_shared_excessive']
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:60])
    for i,h in enumerate(hdr):
        if any(h.startswith(w) for w in want) and not h.endswith('per_second') and '.pct_of_peak_sustained_elapsed' not in h[40:]: print(f'  {h} [{rows[1][i]}] = {r[i]}')
"
  ncu -i $REP --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/_src.csv
  echo; echo "## warp-instructions by kernel phase"; python profiles/ncu_phases.py /tmp/_src.csv ${3:-4096}
  echo; echo "## top source lines"; python profiles/ncu_lines.py /tmp/_src.csv 25
} > $OUT
echo wrote $OUT
