cd "$(dirname "$0")/.."
run() { env "$@" SS_BENCH_NO_BATCH=1 SS_BENCH_NO_FALLBACK=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['roofline_encoder']['frac'], d['roofline']['launch_ms'], d['clocks'])"; }
echo "== split 2" | tee -a gpurun_out/s8_enc.txt; run A=1 | tee -a gpurun_out/s8_enc.txt
echo "== split 4" | tee -a gpurun_out/s8_enc.txt; run SS_ATTN_SPLIT=4 | tee -a gpurun_out/s8_enc.txt
SS_ATTN_SPLIT=4 timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_transcribe.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s8_tests_split4.txt
cp speaksense_b200/lib/libspeaksense_whisper.so /tmp/m.so; cp speaksense_b200/lib/variants/att_lsum.so speaksense_b200/lib/libspeaksense_whisper.so
echo "== lsum f32, split 2" | tee -a gpurun_out/s8_enc.txt; run A=1 | tee -a gpurun_out/s8_enc.txt
echo "== lsum f32, split 4" | tee -a gpurun_out/s8_enc.txt; run SS_ATTN_SPLIT=4 | tee -a gpurun_out/s8_enc.txt
SS_ATTN_SPLIT=4 timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s8_tests_lsum.txt
cp /tmp/m.so speaksense_b200/lib/libspeaksense_whisper.so
SS_ATTN_SPLIT=4 timeout 100 python tools/batch_bench.py large-v3 32 2 1 2>&1 | tee -a gpurun_out/s8_batch.json | cut -c1-400
