set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_v10.err; python -c "import json;d=json.load(open('gpurun_out/bench_v10.json'));print('V10',d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 900 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_p8.json 2> gpurun_out/bench_p8.err; echo "bench paired rc=$?"
tail -2 gpurun_out/bench_p8.err; python -c "import json;d=json.load(open('gpurun_out/bench_p8.json'));print('P8',d['value'],d['e2e']['value'],d['ms_per_step'])"
