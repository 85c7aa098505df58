"""Decode-kernel timing / cycle breakdown: python tools/mega_prof.py [shape] [steps]
Prints ms per decode step; the per-phase / per-stage cycle counters additionally need a profiling build of the library
(`SS_MEGA_PROFILE=1 python -m speaksense_b200.build --force`; the counters cost ~13 % and are compiled out by default).
SS_NO_PROF=1: plain timing (tools/ab.sh)."""
import os, sys
if not os.environ.get("SS_NO_PROF"):
    os.environ["SS_MEGA_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import AsrParams, WhisperAsr, synth  # noqa: E402
shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
eng = WhisperAsr(path)
st = eng.create_state()
eng.upload_pcm(st, synth.synth_audio(seed=1234))
res = eng.transcribe_resident(st, AsrParams(language=None if shape.endswith(".en") else "en", stream_mode=True))
import hashlib  # noqa: E402
print("result-hash", hashlib.md5(repr([(s.start, s.end, s.text) for s in res.segments]).encode()).hexdigest()[:12], len(res.segments), flush=True)
for i in range(3):
    print("ms/step", eng.bench_decode_steps(st, steps, 0), flush=True)
