"""Batch benchmark (BASELINE.json configs[2]: batch = 32 x 30 s clips on one B200): `ss_transcribe_batch` over B synthetic
clips, clip-by-clip decode (default) against the batched decoder (SS_BATCH_DECODE=1, csrc/decoder_batch.cu), host
buffers in, results checked equal between the two.  Prints one JSON line.

    python tools/batch_bench.py [shape] [batch] [reps] [modes: "01" | "0" | "1"]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import AsrParams, WhisperAsr, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
modes = sys.argv[4] if len(sys.argv) > 4 else "01"
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
clips = [synth.synth_audio(seed=1234 + i) for i in range(batch)]
eng = WhisperAsr(path)
states = [eng.create_state() for _ in clips]
p = AsrParams(language="en", stream_mode=True)
out = {"workload": "ggml-%s (synthetic peaked), %d x 30 s clips, one GPU, ss_transcribe_batch, host buffers" % (shape, batch)}
results = {}
for mode in modes:
    os.environ["SS_BATCH_DECODE"] = mode
    best = None
    for rep in range(reps + 1):      # first pass warms up
        t0 = time.perf_counter()
        res = eng.transcribe_batch(states, clips, p)
        dt = time.perf_counter() - t0
        if (rep or reps == 0) and (best is None or dt < best):
            best = dt
    results[mode] = [(r.full_text, [(s.start, s.end) for s in r.segments]) for r in res]
    st = [s.stats() for s in states]
    key = "batched" if mode == "1" else "clip_by_clip"
    out[key] = {"rtf": 30.0 * batch / best, "wall_s": best, "tokens": sum(len(s.result_tokens()[0]) for s in states),
                "launches": sum(x["n_launches"] for x in st), "decode_ms": sum(x["decode_ms"] for x in st),
                # batched: the clips' encoders run concurrently on their own streams, so their per-clip device times overlap;
                # what the encode phase (log-mel + encoders + cross-KV of all clips) costs is the wall time outside the decode
                "encoder_ms_sum_of_overlapping_clips": sum(x["encoder_ms"] for x in st),
                "mel_encode_host_ms": best * 1e3 - sum(x["decode_ms"] for x in st)}
if len(results) == 2:
    out["results_equal"] = results["0"] == results["1"]
print(json.dumps(out))
eng.close()
