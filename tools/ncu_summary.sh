#!/bin/bash
# usage: tools/ncu_summary.sh gpurun_out/prof_X.ncu-rep <env-substeps in the launch>   -> prints headline metrics, stall shares, regions
REP=$1; UNIT=${2:-1223000}
ncu -i $REP --page details 2>/dev/null | grep -vE "^\s+(OPT|INF)|^\s{10}" | grep -E "Duration|Registers|Executed Ipc|Issue Slots Busy|Warp Cycles Per Issued|Achieved Occupancy|Achieved Active|L1/TEX Hit|Eligible Warps|Active Warps Per Sch|Dynamic Shared|Branch Eff|Avg. Active Threads"
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
vals={}
for i,c in enumerate(h):
    if 'smsp__pcsamp_warps_issue_stalled' in c and not c.endswith('_not_issued'):
        try: vals[c.replace('smsp__pcsamp_warps_issue_stalled_','')]=float(rows[2][i])
        except: pass
    if c in ('smsp__inst_executed.sum','sm__icc_request_hit_rate.pct','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum','dram__bytes_read.sum','dram__bytes_write.sum'):
        print(c, rows[2][i], rows[1][i])
tot=sum(vals.values())
print(' '.join('%s %.3f'%(k, v/tot) for k,v in sorted(vals.items(), key=lambda kv:-kv[1])[:9]))
"
ncu -i $REP --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
python tools/ncu_regions_wpe.py /tmp/_src.csv $UNIT | head -${3:-32}
