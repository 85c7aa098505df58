"""Streaming caller benchmark (BASELINE config 5 shape on ONE GPU): one 5-minute synthetic stream, PCM16LE -> base64 ->
32 KiB messages, through AsrStreamSession (gRPC handler semantics: 5 s chunks, 0.5 s overlap, denoise + transcribe per
chunk on one state).  Prints one JSON line.   python tools/stream_bench.py [shape] [seconds] [beam_size] [grpc_streams] [batching: 0 | 1] [n_gpus]
With grpc_streams > 0 the same messages go through the gRPC server (proto/asr.proto, speaksense_b200/grpc_server.py) as that
many concurrent client streams against ONE GPU (each stream = its own ss_state; the decode kernels of different streams
take turns on the device).  batching = 1 puts the micro-batching front end (speaksense_b200/batching.py) between the
sessions and the engine: chunks of different streams that are ready together share one batched decoder step per token.
n_gpus > 1 (BASELINE configs[4]): ONE process owns that many GPUs (ss_engine_open_multi: one weight replica per device, in-process NCCL
broadcast); every stream's state is pinned to the least loaded replica, i.e. one stream per GPU for 8 streams on 8 GPUs."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import WhisperAsr, stream, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 300
beam = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n_grpc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
batching = len(sys.argv) > 5 and sys.argv[5] == "1"
n_gpus = int(sys.argv[6]) if len(sys.argv) > 6 else 1
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
eng = WhisperAsr(path, devices=list(range(n_gpus))) if n_gpus > 1 else WhisperAsr(path)
clips = [synth.synth_audio(seed=5000 + i) for i in range((seconds + 29) // 30)]
pcm = np.concatenate(clips)[:seconds * 16000]
msgs = stream.encode_messages(pcm)
for rep in range(2):      # first pass warms up (allocations, first launches)
    ses = stream.AsrStreamSession(eng)
    if beam > 1:
        ses.params.beam_size = beam
    n_resp = 0
    t0 = time.perf_counter()
    for m, e in msgs:
        n_resp += len(ses.feed(m, e, "bench"))
    dt = time.perf_counter() - t0
    n_chunks = ses.n_chunks
    ses.close()
if n_grpc > 0:
    import threading
    from speaksense_b200 import grpc_server

    from speaksense_b200 import BatchingEngine
    front = BatchingEngine(eng, max_batch=32, linger_s=0.001) if batching else eng

    def factory():
        ses = stream.AsrStreamSession(front)
        if beam > 1:
            ses.params.beam_size = beam
        return ses
    server = grpc_server.serve(eng, "127.0.0.1:0", max_workers=2 * n_grpc, session_factory=factory)
    addr = "127.0.0.1:%d" % server.bound_port
    counts = [0] * n_grpc

    def run(i):
        counts[i] = sum(1 for _ in grpc_server.transcribe_stream(addr, msgs, "bench-%d" % i))
    th = [threading.Thread(target=run, args=(i,)) for i in range(n_grpc)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt_g = time.perf_counter() - t0
    server.stop(0)
    extra = {}
    if batching:
        extra = {"batches": front.n_batches, "requests": front.n_requests, "largest_batch": front.max_seen}
        front.close()
    print(json.dumps({"workload": "%d concurrent gRPC streams (asr.Asr/Transcribe over localhost) of %d s each on %d GPU(s) owned by one process, ggml-%s synthetic, "
                                  "beam_size=%d, micro-batching front end %s" % (n_grpc, seconds, n_gpus, shape, beam, "on" if batching else "off"),
                      "aggregate_stream_rtf": n_grpc * seconds / dt_g, "wall_s": dt_g, "responses": counts, "n_gpus": n_gpus, **extra}))
print(json.dumps({"workload": "one %d s stream, ggml-%s synthetic, 32 KiB base64 PCM16 messages, gRPC handler semantics "
                              "(5 s chunks, 4.5 s advance, denoise + transcribe per chunk), beam_size=%d" % (seconds, shape, beam),
                  "stream_rtf": seconds / dt, "wall_s": dt, "chunks": n_chunks, "ms_per_chunk": dt / max(n_chunks, 1) * 1e3,
                  "responses": n_resp, "messages": len(msgs)}))
eng.close()
