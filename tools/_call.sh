cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s7_tests.txt
for cfg in "A=1" "SS_ENC_PDL=0" "SS_ENC_GRAPH=0" "SS_ENC_GRAPH=0 SS_ENC_PDL=0"; do
  echo "== $cfg" | tee -a gpurun_out/s7_enc.txt
  env $cfg SS_BENCH_NO_BATCH=1 SS_BENCH_NO_FALLBACK=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['roofline_encoder']['frac'], d['gpu_launches'])" | tee -a gpurun_out/s7_enc.txt
done
timeout 100 python tools/batch_bench.py large-v3 32 2 1 2>&1 | tee -a gpurun_out/s7_batch.json | cut -c1-400
