set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r01_final.err
timeout 900 python bench.py --workload paired > gpurun_out/bench_r01_final_paired.json 2> gpurun_out/bench_r01_final_paired.err; echo "bench paired rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_final_reference.json 2> gpurun_out/bench_r01_final_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --impl reference --workload paired --steps 3 --warmup 1 > gpurun_out/bench_r01_final_reference_paired.json 2> gpurun_out/bench_r01_final_reference_paired.err; echo "ref paired rc=$?"
python - <<'PY'
import json
for f in ('bench_r01_final','bench_r01_final_paired','bench_r01_final_reference','bench_r01_final_reference_paired'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('cpu_baseline',{}).get('value'), d.get('roofline',{}).get('frac'), d.get('clocks'))
PY
