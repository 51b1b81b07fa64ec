#!/bin/bash
# usage: tools/summarize_ncu.sh gpurun_out/prof_push_l8.ncu-rep profiles/r01_push_l8 <warps*substeps>
REP=$1; OUT=$2; WS=${3:-306176}
ncu -i $REP --page details 2>/dev/null | grep -vE "^\s+(OPT|INF)|^\s{10}" > $OUT.details.txt
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for i,c in enumerate(h):
    if any(k in c for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','launch__registers','launch__block_size','launch__grid_size','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','thread_inst_executed_per_inst','sm__warps_active','smsp__pcsamp_warps_issue_stalled','launch__shared_mem')) and not c.endswith('_not_issued'):
        print(c, rows[1][i], rows[2][i] if len(rows)>2 else '')
" > $OUT.raw.txt
ncu -i $REP --page source --csv --print-source cuda,sass > /tmp/_src_cs.csv 2>/dev/null
python tools/ncu_regions.py /tmp/_src_cs.csv $WS > $OUT.regions.txt
echo "wrote $OUT.{details,raw,regions}.txt"
