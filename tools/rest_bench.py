"""REST task-processor benchmark: one synthetic 16 kHz PCM16 WAV through TranscribeProcessor.process_audio
(4096-sample chunks -> StreamAudioProcessor frames, denoised on the GPU in batched launches -> 30 s buffers -> stream-mode
transcribe on one state).  Prints one JSON line.   python tools/rest_bench.py [shape] [seconds]"""
import json
import os
import sys
import tempfile
import time
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import WhisperAsr, rest, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 300
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
pcm = np.concatenate([synth.synth_audio(seed=7000 + i) for i in range((seconds + 29) // 30)])[:seconds * 16000]
wav = os.path.join(tempfile.gettempdir(), "ss_rest_bench.wav")
with wave.open(wav, "wb") as w:
    w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
    w.writeframes((np.clip(pcm, -1, 1) * 32767).astype("<i2").tobytes())
eng = WhisperAsr(path)
proc = rest.TranscribeProcessor(eng)
for rep in range(2):      # first pass warms up
    t0 = time.perf_counter()
    r = proc.process_audio(wav, language="en")
    dt = time.perf_counter() - t0
print(json.dumps({"workload": "one %d s 16 kHz PCM16 WAV, ggml-%s synthetic, REST processor semantics (2048-sample frames, VAD gain, "
                              "per-frame denoise, 30 s buffers, stream-mode transcribe)" % (seconds, shape),
                  "rest_rtf": seconds / dt, "wall_s": dt, "transcribe_calls": r.n_calls, "segments": len(r.segments),
                  "text_bytes": len(r.text.encode("utf-8"))}))
eng.close()
