set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_pin.json 2> gpurun_out/bench_pin.err; tail -2 gpurun_out/bench_pin.err; python -c "import json;d=json.load(open('gpurun_out/bench_pin.json'));print('PINS',d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 600 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_pinp.json 2> gpurun_out/bench_pinp.err; python -c "import json;d=json.load(open('gpurun_out/bench_pinp.json'));print('PINP',d['value'],d['e2e']['value'],d['ms_per_step'])"
