set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for rep in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; python -c "import json;d=json.load(open('gpurun_out/bench_b4.json'));print('B4S',d['value'],d['e2e']['value'],d['ms_per_step'])"
done
timeout 600 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_b4p.json 2> gpurun_out/bench_b4p.err; python -c "import json;d=json.load(open('gpurun_out/bench_b4p.json'));print('B4P',d['value'],d['e2e']['value'],d['ms_per_step'])"
