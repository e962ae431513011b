set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sg_align_kernel -s 1 -c 1 -o gpurun_out/prof_align_v2 python bench.py --genome-mbp 240 --steps 1 --warmup 1 --batch-reads 131072 --no-cpu-baseline --no-seed-phase > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
