#!/bin/bash
# profiles/refresh_traffic.sh -- run ON THE GPU BOX (gpurun): one ncu capture of ONE step's alignment launches per workload, plus the seed-lookup
# kernel; after each capture profiles/extract_traffic.py adds DRAM bytes and warp instructions per step, stamped with the hashes of the kernel
# sources and of the kernels' SASS (profiles/kernel_stamp.py), to profiles/traffic.json (copied to gpurun_out/traffic.json as it grows: copy that
# file back to profiles/).  bench.py quotes `roofline.traffic` only from a capture whose stamp matches the kernels it runs.
#
# MODE=full    ncu --set full (39 passes per launch, ~2.5 min per workload: the captures behind profiles/r02_ncu_*_details.csv)
# MODE=counters (default) only the counters the line quotes: dram__bytes_read.sum, dram__bytes_write.sum, smsp__inst_executed.sum,
#              gpu__time_duration.sum -- the same hardware counters `--set full` reads them from, in one or two passes (~1 min per workload)
# usage: refresh_traffic.sh [workload ...]      workloads: single paired lookup ag_d20 ne_d20 (default: all, in that order)
set -u
mkdir -p gpurun_out
MODE=${MODE:-counters}
B="python bench.py --only-headline --no-cpu-baseline --no-sam-phase --steps 2 --warmup 3"
if [ "$MODE" = full ]; then NCU="ncu --set full --clock-control none"; else
  NCU="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none"; fi
R=/tmp/snapgpu_prof; mkdir -p $R
W="$*"; [ -z "$W" ] && W="single paired lookup ag_d20 ne_d20"
cp profiles/traffic.json gpurun_out/traffic_before.json 2>/dev/null
for w in $W; do
  case $w in
    single) SEL="-k regex:sg_align_kernel -s 8 -c 2"; ARGS="--no-seed-phase";;
    paired) SEL="-k regex:sg_align_paired_kernel -s 16 -c 4"; ARGS="--no-seed-phase --workload paired";;
    ag_d20) SEL="-k regex:sg_align_kernel -s 8 -c 2"; ARGS="--no-seed-phase --workload ag_d20";;
    ne_d20) SEL="-k regex:sg_align_kernel -s 4 -c 1"; ARGS="--no-seed-phase --workload ne_d20";;
    lookup) SEL="-k regex:sg_lookup -s 5 -c 1"; ARGS="";;
    *) echo "unknown workload $w"; continue;;
  esac
  $NCU $SEL -f -o $R/prof_$w $B $ARGS > gpurun_out/prof_$w.log 2>&1
  python profiles/extract_traffic.py ${w}_1048576reads_3000mbp $R/prof_$w.ncu-rep "$NCU $SEL $B $ARGS" >> gpurun_out/extract_traffic.log 2>&1
  cp profiles/traffic.json gpurun_out/traffic.json
  ncu -i $R/prof_$w.ncu-rep --page raw --csv > gpurun_out/prof_${w}_raw.csv 2>/dev/null
  echo "$w done $(date +%T)"
done
