set -x
mkdir -p gpurun_out
for lib in "" /root/repo/gpurun_ab_discard.so; do
export SNAPGPU_LIB=$lib
tag=$( [ -z "$lib" ] && echo base || echo discard )
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sg_align_kernel -s 4 -c 1 --csv --log-file gpurun_out/traffic_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-seed-phase > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sg_align_paired_kernel -s 8 -c 1 --csv --log-file gpurun_out/trafficp_$tag.csv python bench.py --workload paired --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
for f in ('gpurun_out/traffic_$tag.csv','gpurun_out/trafficp_$tag.csv'):
    rows=list(csv.reader(open(f))); hdr=[r for r in rows if 'Metric Name' in r][0]
    mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
    print('$tag', f, {r[mi]: r[vi] for r in rows if len(r)==len(hdr) and r[mi] != 'Metric Name'})
PY
timeout 600 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_dp.json 2> gpurun_out/bench_dp.err; python -c "import json;d=json.load(open('gpurun_out/bench_dp.json'));print('DISCP','$tag',d['value'],d['e2e']['value'],d['ms_per_step'])"
done
export SNAPGPU_LIB=/root/repo/gpurun_ab_discard.so
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
