"""Streaming caller benchmark (BASELINE config 5 shape on ONE GPU): one 5-minute synthetic stream, PCM16LE -> base64 ->
32 KiB messages, through AsrStreamSession (gRPC handler semantics: 5 s chunks, 0.5 s overlap, denoise + transcribe per
chunk on one state).  Prints one JSON line.   python tools/stream_bench.py [shape] [seconds] [beam_size]"""
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
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
eng = WhisperAsr(path)
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
print(json.dumps({"workload": "one %d s stream, ggml-%s synthetic, 32 KiB base64 PCM16 messages, gRPC handler semantics "
                              "(5 s chunks, 4.5 s advance, denoise + transcribe per chunk), beam_size=%d" % (seconds, shape, beam),
                  "stream_rtf": seconds / dt, "wall_s": dt, "chunks": n_chunks, "ms_per_chunk": dt / max(n_chunks, 1) * 1e3,
                  "responses": n_resp, "messages": len(msgs)}))
eng.close()
