# usage: ncu_full.sh <tag> <kernel regex> <skip> <count> <bench args...>
# One `ncu --set full` capture, summarised ON THE GPU BOX (the .ncu-rep files are too big to travel back): per-kernel
# headline metrics + SASS instruction mix + stall shares (profiles/sass_summary.py) and the raw metrics that the
# roofline discussion quotes (FP64 pipe, DRAM bytes, duration).
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
REP=/tmp/${TAG}.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o /tmp/${TAG} python bench.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
python profiles/sass_summary.py $REP ${NCU_TOP:-0} > gpurun_out/${TAG}_full.txt 2>&1
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    hdr = rows[0]
    keep = [i for i, h in enumerate(hdr) if h in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct')]
    for r in rows:
        print(' | '.join(r[i] for i in keep))
" >> gpurun_out/${TAG}_full.txt 2>&1
rm -f $REP
