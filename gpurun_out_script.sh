set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for lib in "" /root/repo/gpurun_ab_nopacked.so; do
export SNAPGPU_LIB=$lib
tag=$( [ -z "$lib" ] && echo packed || echo nopacked )
timeout 900 python bench.py --no-cpu-baseline --no-seed-phase --steps 4 --warmup 3 > gpurun_out/bench_v7_$tag.json 2> gpurun_out/bench_v7_$tag.err; echo "bench $tag rc=$?"
tail -2 gpurun_out/bench_v7_$tag.err; python -c "import json;d=json.load(open('gpurun_out/bench_v7_$tag.json'));print('V7','$tag',d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 900 python bench.py --workload paired --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_p4_$tag.json 2> gpurun_out/bench_p4_$tag.err; echo "bench paired $tag rc=$?"
tail -2 gpurun_out/bench_p4_$tag.err; python -c "import json;d=json.load(open('gpurun_out/bench_p4_$tag.json'));print('P4','$tag',d['value'],d['e2e']['value'],d['ms_per_step'])"
done
