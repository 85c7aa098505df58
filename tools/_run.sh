cd /root/repo
timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -2
SS_GEMM_2CTA=0 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file /tmp/g.csv python tools/gemm_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/g.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
out=[]
for r in rows[start:]:
    if len(r)>vi: out.append('%s g%s %.1fus'%(r[ki].split('kernel')[1].split('(')[0], r[gi].split(',')[0][1:], float(r[vi].replace(',',''))/1e3))
print(' | '.join(out[1::2]))
PY
SS_BENCH_NO_BATCH=1 SS_BENCH_NO_FALLBACK=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['stage_ms_per_step'], d['roofline_encoder']['frac'])"
timeout 100 python tools/batch_bench.py large-v3 32 2 1 2>/dev/null | cut -c100-420
timeout 100 python -m pytest tests/test_gpu_transcribe.py -x -q -m gpu -k "cap_then" 2>&1 | tail -2
