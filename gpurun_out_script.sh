set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for mb in 4 3; do
SNAPGPU_BLOCKS_PER_SM=$mb timeout 900 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_v6_mb$mb.json 2> gpurun_out/bench_v6_mb$mb.err; echo "bench mb=$mb rc=$?"
tail -2 gpurun_out/bench_v6_mb$mb.err; python -c "import json;d=json.load(open('gpurun_out/bench_v6_mb$mb.json'));print('V6 MB',$mb,d['value'],d['e2e']['value'],d['ms_per_step'])"
SNAPGPU_PAIRED_BLOCKS_PER_SM=$mb timeout 900 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_p3_mb$mb.json 2> gpurun_out/bench_p3_mb$mb.err; echo "bench paired mb=$mb rc=$?"
tail -2 gpurun_out/bench_p3_mb$mb.err; python -c "import json;d=json.load(open('gpurun_out/bench_p3_mb$mb.json'));print('P3 MB',$mb,d['value'],d['e2e']['value'],d['ms_per_step'])"
done
