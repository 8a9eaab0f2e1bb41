#!/bin/bash
# usage: ncu_summ.sh OUT_PREFIX KERNEL_REGEX SKIP COUNT -- command ...
# One `ncu --set full` capture, reduced ON THE BOX to text (details page + the raw metrics the
# roofline needs + the hottest source lines); the .ncu-rep itself is removed (gpurun copies back
# at most 64 MiB).
out=$1; rx=$2; skip=$3; cnt=$4; shift 5
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $out "$@" > $out.log 2>&1
tail -2 $out.log
ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
ncu -i $out.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
keys=('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','dram__throughput.avg.pct_of_peak_sustained_elapsed')
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('kernel:', d.get('Kernel Name'))
    for k in keys: print('  ',k,'=',d.get(k), rows[1][hdr.index(k)] if k in hdr else '')
    tot=sum(float(v.replace(',','')) for h,v in d.items() if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and v)
    for h,v in sorted(((h,float(v.replace(',',''))) for h,v in d.items() if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and v), key=lambda t:-t[1])[:6]:
        print('   stall %-45s %5.1f%%'%(h.replace('smsp__pcsamp_warps_issue_stalled_',''), 100*v/max(tot,1)))
" > $out.raw.txt
rm -f $out.ncu-rep
