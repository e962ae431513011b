set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for tag in v3packed v3nopacked; do
case $tag in v3packed) export SNAPGPU_LIB= ;; v3nopacked) export SNAPGPU_LIB=/root/repo/gpurun_ab_v2nopacked.so ;; esac
timeout 900 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_v9_$tag.json 2> gpurun_out/bench_v9_$tag.err; echo "bench $tag rc=$?"
tail -2 gpurun_out/bench_v9_$tag.err; python -c "import json;d=json.load(open('gpurun_out/bench_v9_$tag.json'));print('V9','$tag',d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 900 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_p7_$tag.json 2> gpurun_out/bench_p7_$tag.err; echo "bench paired $tag rc=$?"
tail -2 gpurun_out/bench_p7_$tag.err; python -c "import json;d=json.load(open('gpurun_out/bench_p7_$tag.json'));print('P7','$tag',d['value'],d['e2e']['value'],d['ms_per_step'])"
done
export SNAPGPU_LIB=
SNAPGPU_TEST_WORKERS=8192 SNAPGPU_TEST_REPEAT=5 AGB_WARPS=4736 AGB_BANDED=0 AGB_W=29 python tests/agbench.py 2>&1 | tail -2
