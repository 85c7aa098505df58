#!/bin/bash
# One gpurun call that re-establishes the measured numbers of the repo on a fresh 1 x B200 box (about 8 minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh r2z'
# Outputs land in gpurun_out/<tag>_*; copy what is to be kept into profiles/ (tools/ncu_summary.py for the decode capture).
cd "$(dirname "$0")/.."
TAG=${1:-check}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "gpu tests rc=$? $(tail -1 $OUT/${TAG}_tests.log)"
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench_reference.json
timeout 100 python bench.py --impl reference --shape tiny.en --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_tiny_en.json 2>> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench_reference_tiny_en.json
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; tail -1 $OUT/${TAG}_ncu_bench.log | cut -c1-200
# the decode kernel, --set full with source (16 decode steps in one launch)
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_mega -c 1 -o $OUT/${TAG}_mega -f \
    python tools/profile_clip.py large-v3 steps > $OUT/${TAG}_ncu_mega.log 2>&1; tail -2 $OUT/${TAG}_ncu_mega.log
# the batched step: per-kernel durations and DRAM bytes (large-v3 shapes, 2 layers, 32 sequences)
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bd_ -c 300 --csv \
    --log-file $OUT/${TAG}_batch_launches.csv python tools/batch_bench.py large-v3-l2 32 0 1 > $OUT/${TAG}_ncu_batch.log 2>&1; tail -1 $OUT/${TAG}_ncu_batch.log | cut -c1-200
# the critical path of one decode step (trace build of the decode kernel, if tools/build_variant.sh t_trace was run)
if [ -f speaksense_b200/lib/variants/t_trace.so ]; then
  cp speaksense_b200/lib/libspeaksense_whisper.so /tmp/main.so; cp speaksense_b200/lib/variants/t_trace.so speaksense_b200/lib/libspeaksense_whisper.so
  timeout 200 python tools/mega_trace.py large-v3 $OUT/${TAG}_mega_trace.txt > /dev/null 2>&1; head -14 $OUT/${TAG}_mega_trace.txt
  cp /tmp/main.so speaksense_b200/lib/libspeaksense_whisper.so
fi
# BASELINE configs[2] and small batches; concurrent streams on one GPU
for B in 4 8 16 32; do timeout 100 python tools/batch_bench.py large-v3 $B 2 1 2>> $OUT/${TAG}_bench.err | tee -a $OUT/${TAG}_batch.json | cut -c1-330; done
timeout 150 python tools/stream_bench.py large-v3 120 0 8 1 > $OUT/${TAG}_stream8_batched.json 2>> $OUT/${TAG}_bench.err; head -1 $OUT/${TAG}_stream8_batched.json | cut -c1-400
timeout 150 python tools/stream_bench.py large-v3 60 5 > $OUT/${TAG}_stream_beam5.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_stream_beam5.json | cut -c1-400
