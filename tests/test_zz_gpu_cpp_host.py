"""GPU: the C++ mirror of the trait (include/speaksense_asr.hpp) driven by examples/asr_host.cpp gives what the Python mirror
gives on the same model and clips - single call (create_state + transcribe_with_state) and transcribe_batch.
(Named to sort last: it compiles a program, the rest of the suite does not depend on it.)"""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _full_texts(stdout: str):
    return [ln[len("full_text: "):] for ln in stdout.split("\n") if ln.startswith("full_text: ")]


def test_cpp_host_matches_python_mirror(tmp_path, micro_v3_peaked):
    from speaksense_b200 import AsrParams, WhisperAsr, build, synth
    if not shutil.which("g++"):
        pytest.skip("no g++")
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(build.LIB)
    exe = tmp_path / "asr_host"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", inc, os.path.join(ROOT, "examples", "asr_host.cpp"), "-o", str(exe),
                           "-L", libdir, "-lspeaksense_whisper", "-Wl,-rpath," + libdir])
    clips = [synth.synth_audio(seed=700 + i) for i in range(3)]
    files = []
    for i, c in enumerate(clips):
        f = tmp_path / ("clip%d.f32" % i)
        np.ascontiguousarray(c, dtype="<f4").tofile(f)
        files.append(str(f))
    eng = WhisperAsr(micro_v3_peaked)
    p = AsrParams(language="zh", stream_mode=True, min_segment_length=5)      # what examples/asr_host.cpp sets (asr.rs:154-157)
    ref = [eng.transcribe(c, p) for c in clips]
    eng.close()
    one = subprocess.run([str(exe), micro_v3_peaked, files[0]], capture_output=True, text=True, timeout=120)
    assert one.returncode == 0, one.stderr
    assert _full_texts(one.stdout) == [ref[0].full_text]
    assert one.stdout.count("speaker ") == len(ref[0].segments)
    many = subprocess.run([str(exe), micro_v3_peaked] + files, capture_output=True, text=True, timeout=120)
    assert many.returncode == 0, many.stderr
    assert _full_texts(many.stdout) == [r.full_text for r in ref]
    bad = subprocess.run([str(exe), str(tmp_path / "missing.bin"), files[0]], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "failed to open whisper model: " in bad.stderr
