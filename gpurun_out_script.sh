set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for mb in 3 4 2; do
SNAPGPU_PAIRED_BLOCKS_PER_SM=$mb timeout 900 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_p1_mb$mb.json 2> gpurun_out/bench_p1_mb$mb.err; echo "bench paired mb=$mb rc=$?"
tail -3 gpurun_out/bench_p1_mb$mb.err; python -c "import json;d=json.load(open('gpurun_out/bench_p1_mb$mb.json'));print('P1 MB',$mb,d['value'],d['e2e']['value'],d['ms_per_step'],d['per_read'])"
done
timeout 900 python bench.py --workload paired --steps 3 --warmup 3 --cpu-sample-reads 400000 > gpurun_out/bench_p1_full.json 2> gpurun_out/bench_p1_full.err; echo "bench paired full rc=$?"; tail -3 gpurun_out/bench_p1_full.err; cat gpurun_out/bench_p1_full.json
timeout 600 python bench.py --no-cpu-baseline --no-seed-phase --steps 3 --warmup 3 > gpurun_out/bench_s_check.json 2> gpurun_out/bench_s_check.err; echo "single rc=$?"; cat gpurun_out/bench_s_check.json | head -c 600
