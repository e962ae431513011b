#!/bin/bash
# profiles/refresh_traffic.sh -- run ON THE GPU BOX (gpurun): one `ncu --set full` capture of ONE step's alignment launches per workload,
# plus the seed-lookup kernel, then profiles/extract_traffic.py writes DRAM bytes and warp instructions per step, stamped with the hash
# of the kernel sources, into gpurun_out/traffic.json (copy it to profiles/traffic.json).  bench.py quotes `roofline.traffic` only from
# a capture whose csrc_sha16 equals the sources it runs.
set -u
mkdir -p gpurun_out
B="python bench.py --only-headline --no-cpu-baseline --no-sam-phase --steps 2 --warmup 3"
NCU="ncu --set full --clock-control none"
R=/tmp/snapgpu_prof; mkdir -p $R
$NCU -k regex:sg_align_kernel -s 8 -c 2 -f -o $R/prof_single $B --no-seed-phase > gpurun_out/prof_single.log 2>&1
$NCU -k regex:sg_align_paired_kernel -s 16 -c 4 -f -o $R/prof_paired $B --no-seed-phase --workload paired > gpurun_out/prof_paired.log 2>&1
$NCU -k regex:sg_align_kernel -s 8 -c 2 -f -o $R/prof_ag_d20 $B --no-seed-phase --workload ag_d20 > gpurun_out/prof_ag_d20.log 2>&1
$NCU -k regex:sg_align_kernel -s 4 -c 1 -f -o $R/prof_ne_d20 $B --no-seed-phase --workload ne_d20 > gpurun_out/prof_ne_d20.log 2>&1
$NCU -k regex:sg_lookup -s 5 -c 1 -f -o $R/prof_lookup $B > gpurun_out/prof_lookup.log 2>&1
cp profiles/traffic.json gpurun_out/traffic_before.json 2>/dev/null
python profiles/extract_traffic.py \
  single_1048576reads_3000mbp $R/prof_single.ncu-rep "ncu --set full -k regex:sg_align_kernel -s 8 -c 2 $B --no-seed-phase" \
  paired_1048576reads_3000mbp $R/prof_paired.ncu-rep "ncu --set full -k regex:sg_align_paired_kernel -s 16 -c 4 $B --workload paired" \
  ag_d20_1048576reads_3000mbp $R/prof_ag_d20.ncu-rep "ncu --set full -k regex:sg_align_kernel -s 8 -c 2 $B --workload ag_d20" \
  ne_d20_1048576reads_3000mbp $R/prof_ne_d20.ncu-rep "ncu --set full -k regex:sg_align_kernel -s 4 -c 1 $B --workload ne_d20" \
  lookup_1048576reads_3000mbp $R/prof_lookup.ncu-rep "ncu --set full -k regex:sg_lookup -s 5 -c 1 $B" > gpurun_out/extract_traffic.log 2>&1
cp profiles/traffic.json gpurun_out/traffic.json
for r in single paired ag_d20 ne_d20 lookup; do
  ncu -i $R/prof_$r.ncu-rep --page details --csv > gpurun_out/prof_${r}_details.csv 2>/dev/null
  ncu -i $R/prof_$r.ncu-rep --page raw --csv > gpurun_out/prof_${r}_raw.csv 2>/dev/null
done
ls -la $R
