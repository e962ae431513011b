set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for mb in 4 3; do
SNAPGPU_BLOCKS_PER_SM=$mb timeout 900 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_v5_mb$mb.json 2> gpurun_out/bench_v5_mb$mb.err; echo "bench mb=$mb rc=$?"
tail -2 gpurun_out/bench_v5_mb$mb.err; python -c "import json;d=json.load(open('gpurun_out/bench_v5_mb$mb.json'));print('V5 MB',$mb,d['value'],d['e2e']['value'],d['ms_per_step'])"
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sg_align_kernel -s 1 -c 1 -o gpurun_out/prof_align_v5 python bench.py --genome-mbp 240 --steps 1 --warmup 1 --batch-reads 131072 --no-cpu-baseline --no-seed-phase > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
