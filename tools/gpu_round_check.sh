#!/bin/bash
# One gpurun call that re-establishes every measured number of the repo on a fresh B200 box (about 5 minutes):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_round_check.sh r2a'
# Outputs land in gpurun_out/<tag>_*; copy what is to be kept into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-check}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "gpu tests rc=$? $(tail -1 $OUT/${TAG}_tests.log)"
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json
# BASELINE configs[2]: 32 clips, clip-by-clip against the batched decoder, and the batched decoder without PDL
timeout 120 python tools/batch_bench.py large-v3 32 2 > $OUT/${TAG}_batch_bench.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_batch_bench.json
SS_BATCH_PDL=0 timeout 60 python tools/batch_bench.py large-v3 32 2 1 > $OUT/${TAG}_batch_bench_nopdl.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_batch_bench_nopdl.json
# opt-in batched encoder pass (unverified at the end of round 1): parity, then its effect on configs[2]
SS_TEST_BATCH_ENCODER=1 timeout 120 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k encoder_pass > $OUT/${TAG}_tests_batch_encoder.log 2>&1; echo "batched encoder rc=$? $(tail -1 $OUT/${TAG}_tests_batch_encoder.log)"
SS_BATCH_ENCODER=1 timeout 60 python tools/batch_bench.py large-v3 32 2 1 > $OUT/${TAG}_batch_bench_encoder.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_batch_bench_encoder.json
# opt-in beam search on the batched step (unverified at the end of round 1): parity, then one stream with beam 5 both ways
SS_TEST_BATCH_BEAM=1 timeout 120 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k beam_on_batched > $OUT/${TAG}_tests_batch_beam.log 2>&1; echo "batched beam rc=$? $(tail -1 $OUT/${TAG}_tests_batch_beam.log)"
timeout 120 python tools/stream_bench.py large-v3 60 5 > $OUT/${TAG}_stream_beam5.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_stream_beam5.json
SS_BATCH_BEAM=1 timeout 120 python tools/stream_bench.py large-v3 60 5 > $OUT/${TAG}_stream_beam5_batched.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_stream_beam5_batched.json
# smaller batches (where the step is launch / latency bound)
for B in 4 8 16; do timeout 60 python tools/batch_bench.py large-v3 $B 2 1 2>> $OUT/${TAG}_bench.err | tee -a $OUT/${TAG}_batch_small.json; done
# concurrent gRPC streams on one GPU: taking turns against the micro-batching front end
timeout 120 python tools/stream_bench.py large-v3 120 0 8 0 > $OUT/${TAG}_stream8_turns.json 2>> $OUT/${TAG}_bench.err; head -1 $OUT/${TAG}_stream8_turns.json
timeout 120 python tools/stream_bench.py large-v3 120 0 8 1 > $OUT/${TAG}_stream8_batched.json 2>> $OUT/${TAG}_bench.err; head -1 $OUT/${TAG}_stream8_batched.json
# launch list of the batched step (large-v3 shapes, 2 layers): per-kernel durations, serialised by ncu
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bd_ -c 500 --csv --log-file $OUT/${TAG}_batch_launches.csv \
    python tools/batch_bench.py large-v3-l2 32 0 1 > $OUT/${TAG}_ncu_batch.log 2>&1; tail -1 $OUT/${TAG}_ncu_batch.log
