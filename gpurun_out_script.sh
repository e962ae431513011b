set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sg_fastq|DeviceScan" -c 40 --csv --log-file gpurun_out/launches_fastq.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 --genome-mbp 240 > gpurun_out/ncu_fastq.log 2>&1; echo rc=$?
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_fastq.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows:
    if len(r)==len(hdr) and r[ii].isdigit():
        cur.setdefault((int(r[ii]), r[ki][:50]), {})[r[mi]]=r[vi]
for k in sorted(cur)[-16:]: print(k, cur[k])
PY
