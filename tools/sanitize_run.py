"""Workload for compute-sanitizer (tools/sanitize.sh): every kernel family of the library once, on the smallest model shapes -
log-mel, the encoder (tcgen05 GEMMs, fused attention, LayerNorm, graph replay on the second window), the persistent decode kernel
(greedy, two windows with a context prompt), the batched step (3 clips of different lengths, beam search, a temperature-ladder
rung with drawn tokens) and the audio denoiser.  Prints a digest of the results so that a sanitized run can be compared with a
plain one."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import AsrParams, WhisperAsr, audio, synth  # noqa: E402
import numpy as np  # noqa: E402

mdir = os.environ.get("SS_MODEL_DIR", "/tmp/ss_models")
out = []


def digest(res):
    return hashlib.md5(repr([(s.start, s.end, s.text) for s in res.segments]).encode()).hexdigest()[:10]


peaked = os.path.join(mdir, "ggml-tiny.en-peaked-s0.bin")
synth.ensure_model(peaked, shape="tiny.en", family="peaked", seed=0)
eng = WhisperAsr(peaked)
st = eng.create_state()
out.append(("greedy 45 s, two windows, context", digest(eng.transcribe_with_state(st, synth.synth_audio(45 * 16000, seed=11), AsrParams(stream_mode=False)))))
out.append(("greedy 30 s, stream mode", digest(eng.transcribe_with_state(st, synth.synth_audio(seed=1234), AsrParams(stream_mode=True)))))
out.append(("beam 5", digest(eng.transcribe_with_state(st, synth.synth_audio(seed=7), AsrParams(stream_mode=True, beam_size=5)))))
states = [eng.create_state() for _ in range(3)]
clips = [synth.synth_audio(seed=21), synth.synth_audio(12 * 16000, seed=5), synth.synth_audio(40 * 16000, seed=9)]
os.environ["SS_BATCH_MIN"] = "2"
res = eng.transcribe_batch(states, clips, AsrParams(stream_mode=False))
out.append(("batch of 3 (30 s, 12 s, 40 s)", " ".join(digest(r) for r in res)))
den = audio.denoise_audio(eng, st, synth.synth_audio(5 * 16000, seed=3))
out.append(("denoise 5 s", hashlib.md5(np.asarray(den[0] if isinstance(den, tuple) else den, dtype=np.float32).round(3).tobytes()).hexdigest()[:10]))
for s in states:
    s.close()
st.close()
eng.close()
soft = os.path.join(mdir, "ggml-tiny.en-soft10-s0.bin")
try:
    synth.ensure_model(soft, shape="tiny.en", family="soft10", seed=0)
    eng = WhisperAsr(soft)
    st = eng.create_state()
    r = eng.transcribe_with_state(st, synth.synth_audio(seed=1234), AsrParams(stream_mode=True))
    out.append(("soft model (temperature ladder, drawn tokens), fallbacks %d" % st.stats().get("n_fallbacks", -1), digest(r)))
    st.close()
    eng.close()
except Exception as e:  # noqa: BLE001
    out.append(("soft model", "skipped: %r" % (e,)))
for k, v in out:
    print("%-60s %s" % (k, v))
