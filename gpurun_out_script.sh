set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload paired --steps 3 --warmup 3 --cpu-sample-reads 400000 > gpurun_out/bench_p2_full.json 2> gpurun_out/bench_p2_full.err; echo "bench paired full rc=$?"; tail -3 gpurun_out/bench_p2_full.err; cat gpurun_out/bench_p2_full.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sg_align_paired_kernel -s 2 -c 1 -o gpurun_out/prof_paired_v1 python bench.py --workload paired --genome-mbp 240 --steps 1 --warmup 1 --batch-reads 131072 --no-cpu-baseline > gpurun_out/ncu_paired.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_paired.log
