"""One warm transcription of a 30 s clip inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  Usage: python tools/profile_clip.py [shape] [mode]
mode: clip (whole transcribe) | steps (16 decode-step graph replays only)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import AsrParams, WhisperAsr, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
mode = sys.argv[2] if len(sys.argv) > 2 else "clip"
path = os.path.join(os.environ.get("SS_MODEL_DIR", "/tmp/ss_models"), "ggml-%s-peaked-s0.bin" % shape)
synth.ensure_model(path, shape=shape, family="peaked", seed=0)
eng = WhisperAsr(path)
st = eng.create_state()
pcm = synth.synth_audio(seed=1234)
params = AsrParams(language=None if shape.endswith(".en") else "en", stream_mode=True)
eng.upload_pcm(st, pcm)
eng.transcribe_resident(st, params)
eng.bench_decode_steps(st, 8, 0)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaProfilerStart()
if mode == "clip":
    eng.transcribe_resident(st, params)
else:
    eng.bench_decode_steps(st, 16, 0)
rt.cudaProfilerStop()
print("profiled", mode, st.stats())
