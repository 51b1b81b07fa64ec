#!/bin/bash
# usage: tools/summarize_ncu_wpe.sh gpurun_out/prof_X.ncu-rep profiles/r02_name <env-substeps in the captured launch>
# writes profiles/r02_name.{details,raw,regions,lines}.txt (the source tree must be the one the capture was taken on)
R=$1; O=$2; U=${3:-1217000}
ncu -i $R --page details 2>/dev/null | grep -vE "^\s+(OPT|INF)|^\s{10}" > $O.details.txt
ncu -i $R --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for i,c in enumerate(h):
    if any(k in c for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','launch__registers','launch__block_size','launch__grid_size','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','thread_inst_executed_per_inst','sm__warps_active','smsp__pcsamp_warps_issue_stalled','launch__shared_mem','sm__icc_request_hit_rate','op_local')) and not c.endswith('_not_issued'):
        print(c, rows[1][i], rows[2][i] if len(rows)>2 else '')
" > $O.raw.txt
bash tools/ncu_summary.sh $R $U 60 > $O.regions.txt 2>&1
python tools/ncu_lines.py /tmp/_src.csv $U 70 > $O.lines.txt
echo "wrote $O.{details,raw,regions,lines}.txt"
