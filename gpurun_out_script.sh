set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_v4.err; python -c "import json;d=json.load(open('gpurun_out/bench_v4.json'));print('V4',d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sg_align_kernel -s 1 -c 1 -o gpurun_out/prof_align_v4 python bench.py --genome-mbp 240 --steps 1 --warmup 1 --batch-reads 131072 --no-cpu-baseline --no-seed-phase > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
