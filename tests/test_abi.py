"""The C-ABI library loads on a machine without a GPU, exports every symbol include/speaksense_whisper.h
declares, fails loudly (no CPU fallback) on compute entry points, and its host-only helpers agree with
the Python restatement of the reference's Rust text rules."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from speaksense_b200 import build, _native
    build.build()
    return _native.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "speaksense_whisper.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(L):
    from speaksense_b200 import _native
    names = _declared()
    assert len(names) >= 40
    bound = {n for n, _, _ in _native.SYMBOLS}
    for n in names:
        assert hasattr(L, n), "library lacks %s" % n
        assert n in bound, "python binding lacks %s" % n
    assert L.ss_abi_version() == 1
    assert b"sm_100a" in L.ss_build_info()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(L, tiny_en_peaked):
    import numpy as np
    from speaksense_b200 import NativeError, WhisperAsr
    h = C.c_void_p()
    assert L.ss_engine_open(tiny_en_peaked.encode(), 0, C.byref(h)) == -3          # SS_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.ss_last_error()
    with pytest.raises(NativeError):
        WhisperAsr(tiny_en_peaked)
    a = np.zeros((4, 4), np.float16)
    d = np.zeros((4, 4), np.float32)
    assert L.ss_debug_gemm(0, a.ctypes.data, a.ctypes.data, d.ctypes.data, 4, 4, 4, 0) == -3


def test_text_rules_match_rust_restatement(L):
    from tests.rust_post import PROMOTIONAL_TEXT, add_punctuation, is_promotional_text
    buf = C.create_string_buffer(4096)
    samples = ["你好吗", "这是什么", "啊", "真好", "天气不错", "句号。", "逗号，", "感叹！", "问？", "", "plain ascii", "为何怎么呢",
               "请订阅", "打赏支持明镜与点点栏目", "x" * 500] + PROMOTIONAL_TEXT
    for s in samples:
        n = L.ss_add_punctuation(s.encode(), buf, len(buf))
        assert n >= 0 and buf.value.decode() == add_punctuation(s)
        assert bool(L.ss_is_promotional_text(s.encode())) == is_promotional_text(s)
    assert L.ss_add_punctuation("长".encode() * 2000, buf, 16) < 0               # buffer too small is an error, not a truncation
    for b, ok in [(b"abc", 1), ("中文".encode(), 1), (b"\xe4\xb8", 0), (b"\xff", 0), (b"\xc0\x80", 0), (b"\xed\xa0\x80", 0), (b"", 1)]:
        assert L.ss_is_valid_utf8(b, len(b)) == ok


def test_model_probe(L, tiny_en_peaked, micro_v3_random, tmp_path):
    hp = (C.c_int * 11)()
    ab, h, eot, beg, nv = C.c_int64(), C.c_uint64(), C.c_int(), C.c_int(), C.c_int()
    assert L.ss_model_probe(tiny_en_peaked.encode(), C.byref(hp), C.byref(ab), C.byref(h), C.byref(eot), C.byref(beg), C.byref(nv)) == 0
    assert list(hp) == [51864, 1500, 384, 6, 4, 448, 384, 6, 4, 80, 1]
    assert (eot.value, beg.value, nv.value) == (50256, 50363, 50256)
    assert ab.value > os.path.getsize(tiny_en_peaked) * 0.99
    h1 = h.value
    assert L.ss_model_probe(tiny_en_peaked.encode(), None, None, C.byref(h), None, None, None) == 0 and h.value == h1   # deterministic packing
    assert L.ss_model_probe(micro_v3_random.encode(), C.byref(hp), None, None, C.byref(eot), C.byref(beg), None) == 0
    assert (hp[0], eot.value, beg.value) == (51866, 50257, 50365)                  # large-v3 token layout
    # malformed files fail with SS_ERR_IO, like WhisperContext::new_with_params -> Err (whisper.rs:23-24)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"nope" + b"\0" * 100)
    assert L.ss_model_probe(str(bad).encode(), None, None, None, None, None, None) == -2
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(open(tiny_en_peaked, "rb").read(3_000_000))
    assert L.ss_model_probe(str(trunc).encode(), None, None, None, None, None, None) == -2
    assert L.ss_model_probe(b"/nonexistent/model.bin", None, None, None, None, None, None) == -2


def test_header_is_plain_c_and_links(tmp_path, L):
    """the drop-in boundary is a C ABI: include/speaksense_whisper.h compiles as C99 (and C++), and a C program that only
    includes it links against the library and gets the documented no-device error without a GPU"""
    import shutil
    import subprocess
    from speaksense_b200 import build
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "speaksense_whisper.h"\n'
                   'int main(void) { ss_params p; ss_denoise_config c; ss_engine *e = 0; int rc;\n'
                   '  ss_params_default(&p); ss_denoise_config_default(&c);\n'
                   '  rc = ss_engine_open("/nonexistent/model.bin", 0, &e);\n'
                   '  printf("%d %d %d %s\\n", ss_abi_version(), c.frame_size, rc, ss_last_error());\n'
                   '  return rc == 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)])
    exe = tmp_path / "t"
    libdir = os.path.dirname(build.LIB)
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, "-o", str(exe), str(src), "-L", libdir, "-lspeaksense_whisper",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    ver, fs, rc, msg = out.stdout.strip().split(" ", 3)
    assert int(ver) == 1 and int(fs) == 2048 and int(rc) < 0 and len(msg) > 0


def test_loader_survives_corrupt_files(L, micro_v3_random, tmp_path):
    """truncations at every structural boundary and corrupted tensor headers come back as SS_ERR_MODEL (-2), never a crash"""
    import struct
    blob = open(micro_v3_random, "rb").read()
    probe = lambda b: (open(tmp_path / "m.bin", "wb").write(b), L.ss_model_probe(str(tmp_path / "m.bin").encode(), None, None, None, None, None, None))[1]  # noqa: E731
    assert probe(blob) == 0
    # header | mel filters | vocabulary | first tensor header | inside tensor data | last byte
    meta = 4 + 44 + 8 + 128 * 201 * 4
    cuts = [0, 3, 4, 20, 47, 52, 56, meta - 1, meta + 2, meta + 1000, meta + 400000, len(blob) // 2, len(blob) - 1]
    for c in cuts:
        assert probe(blob[:c]) == -2, c
    # find the first tensor header (after the vocabulary) and corrupt its fields one at a time
    o = meta
    n_tok, = struct.unpack_from("<i", blob, o); o += 4
    for _ in range(n_tok):
        ln, = struct.unpack_from("<I", blob, o); o += 4 + ln
    for field, value in ((0, 9), (0, -1), (1, -5), (1, 100000), (2, 99), (2, -3)):      # n_dims, name length, type
        b = bytearray(blob)
        struct.pack_into("<i", b, o + 4 * field, value)
        assert probe(bytes(b)) == -2, (field, value)
    b = bytearray(blob)
    struct.pack_into("<i", b, o + 12, 0x7fffffff)            # absurd first dimension
    assert probe(bytes(b)) == -2
    assert b"" != L.ss_last_error()


def test_text_rules_literal_cases(L):
    """expected values written down by hand from whisper.rs:9-14 (promo list), :41-43 (contains any), :175-201 (punctuation) -
    independent of tests/rust_post.py"""
    buf = C.create_string_buffer(1024)

    def punct(s):
        assert L.ss_add_punctuation(s.encode(), buf, len(buf)) >= 0
        return buf.value.decode()
    assert punct("你好吗") == "你好吗？"            # 吗 -> question
    assert punct("这是什么好东西") == "这是什么好东西？"  # question wins over exclamation (checked first)
    assert punct("太棒了") == "太棒了！"            # 太 -> exclamation
    assert punct("今天开会") == "今天开会 "          # neither -> a single space
    assert punct("hello") == "hello "
    assert punct("") == " "
    for tail in "。！？，":
        assert punct("已有标点" + tail) == "已有标点" + tail      # already punctuated: unchanged
    assert punct("英文标点.") == "英文标点. "       # ASCII '.' is not in the ends_with list
    assert L.ss_is_promotional_text("欢迎订阅我们的频道".encode()) == 1        # contains 订阅
    assert L.ss_is_promotional_text("感谢收看 請按讚、訂閱、分享!".encode()) == 1
    assert L.ss_is_promotional_text("订 阅".encode()) == 0                     # substring match, not fuzzy
    assert L.ss_is_promotional_text("今天天气不错".encode()) == 0


def test_cpp_trait_mirror_compiles_and_reports_errors(tmp_path):
    """include/speaksense_asr.hpp (C++ mirror of mod.rs:9-73 / whisper.rs:16-129 above the C ABI) builds warning-free and the
    example host program runs its host-only part: text rules through the library, and WhisperAsr::new failing with the
    reference's message prefix (whisper.rs:24) - here because the model file does not exist / there is no GPU."""
    import shutil
    import subprocess
    from speaksense_b200 import build
    if not shutil.which("g++"):
        pytest.skip("no g++")
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(build.LIB)
    exe = tmp_path / "asr_host"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, os.path.join(ROOT, "examples", "asr_host.cpp"),
                           "-o", str(exe), "-L", libdir, "-lspeaksense_whisper", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout.splitlines()
    assert out[0] == "promo 1 0"
    assert out[1] == "punct [hello ]"
    assert out[2].startswith("error -") and "failed to open whisper model: " in out[2]


def test_engine_beam_candidate_assignment_hand_derived_cases(L):
    """The function the engine's beam decoders run (beam_pick, csrc/engine.cc) against the hand-derived cases of tests/beam_cases.py -
    on the CPU, through the C ABI, independently of the oracle."""
    import ctypes as C
    import numpy as np
    from tests.beam_cases import CASES
    for name, cands, live, i, want in CASES:
        n = len(cands)
        max_len = max(len(c[0]) for c in cands)
        ids = np.zeros((n, max_len), np.int32)
        for k, c in enumerate(cands):
            ids[k, :len(c[0])] = c[0]
        lens = np.asarray([len(c[0]) for c in cands], np.int32)
        sums = np.asarray([c[1] for c in cands], np.float64)
        dec = np.asarray([c[2] for c in cands], np.int32)
        lv = np.asarray([1 if x else 0 for x in live], np.int32)
        out = np.full(len(live), -2, np.int32)
        rc = L.ss_debug_beam_assign(ids.ctypes.data, lens.ctypes.data, max_len, sums.ctypes.data, dec.ctypes.data, n, lv.ctypes.data, len(live), i,
                                    out.ctypes.data)
        assert rc == 0, name
        assert out.tolist() == want, name
    bad = np.zeros(1, np.int32)
    assert L.ss_debug_beam_assign(None, None, 0, None, None, 1, bad.ctypes.data, 1, 0, bad.ctypes.data) < 0      # null candidate arrays


def test_engine_logits_filter_matches_hf_timestamp_processor(L, micro_v3_random):
    """The ENGINE's host-side whisper_process_logits (process_logits_host, csrc/engine.cc, through ss_debug_process_logits on the CPU)
    against HuggingFace's WhisperTimeStampLogitsProcessor (tests/golden/logits_filter.npz, tools/make_golden_logits_filter.py) - the
    comparison tests/test_oracle_golden.py makes for the oracle, with the same two documented exceptions (W: tokens only whisper.cpp
    suppresses in this function; D: the open timestamp OpenAI forbids repeating), made for the product's own function."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "logits_filter.npz"))
    nv, beg, eot = int(g["n_vocab"]), int(g["beg"]), int(g["eot"])
    sot = eot + 1
    W = np.zeros(nv, bool)
    W[sot:sot + 101] = True                      # sot + the language table
    W[beg - 6:beg - 1] = True                    # translate, transcribe, solm, prev, nosp
    path = micro_v3_random.encode()
    for name in [str(x) for x in g["names"]]:
        ids = np.asarray(g[name + "_ids"], np.int32)
        rng = np.random.default_rng(int(g[name + "_seed"]))
        raw = rng.standard_normal(nv).astype(np.float32) * 2.0
        raw[beg:] += np.float32(g[name + "_boost"])
        has_ts, seek_delta = 0, 0
        for t in ids.tolist():                   # the decoder state whisper_full keeps beside the history
            if t > beg:
                has_ts, seek_delta = 1, 2 * (t - beg)
        out = np.empty(nv, np.float32)
        rc = L.ss_debug_process_logits(path, ids.ctypes.data if ids.size else None, int(ids.size), has_ts, seek_delta, raw.ctypes.data, 0.0,
                                       out.ctypes.data)
        assert rc == 0, name
        got = np.isneginf(out)
        want = np.unpackbits(g[name + "_masked"])[:nv].astype(bool)
        D = np.zeros(nv, bool)
        il = ids.tolist()
        ts = [t for t in il if t >= beg]
        if ts and not (il[-1] >= beg and (len(il) >= 2 and il[-2] < beg)):
            D[ts[-1]] = True
        cmp = ~(W | D)
        assert np.array_equal(got[cmp], want[cmp]), name
        assert got[W].all(), name
        assert np.array_equal(out[~got], raw[~got]), name      # what is not masked passes through unchanged (temperature 0)


def test_engine_sequence_score_hand_derived_cases(L):
    """sequence_score of the engine (whisper_sequence_score + the entropy gate of the temperature ladder) on the hand-derived cases
    tests/test_oracle_rules.py holds the oracle to."""
    import math
    import numpy as np
    import pytest

    def score(ids, plogs, result_len, length_penalty=-1.0):
        a = np.asarray(ids, np.int32); p = np.asarray(plogs, np.float32)
        import ctypes as C
        out = (C.c_double * 4)()
        assert L.ss_debug_sequence_score(a.ctypes.data, p.ctypes.data, int(a.size), result_len, length_penalty, out) == 0
        return dict(sum_logprobs=out[0], avg_logprobs=out[1], entropy=out[2], score=out[3])

    r = score([7] * 40, [-0.5] * 40, 40)                                     # one token repeated: entropy 0 (the repetition detector)
    assert r["entropy"] == pytest.approx(0.0, abs=1e-12) and r["avg_logprobs"] == pytest.approx(-0.5, rel=1e-6)
    assert r["sum_logprobs"] == pytest.approx(-20.0, rel=1e-6)
    assert score(list(range(100, 132)), [-1.0] * 32, 32)["entropy"] == pytest.approx(math.log(32.0), rel=1e-12)
    ids = [5] * 8 + list(range(100, 132))                                    # only the LAST 32 tokens of the first result_len count
    assert score(ids, [-0.25] * len(ids), len(ids))["entropy"] == pytest.approx(math.log(32.0), rel=1e-12)
    r = score([9] * 16 + list(range(200, 216)), [-0.1] * 32, 32)             # (1/2) ln 2 + (1/2) ln 32 = 3 ln 2 < 2.4
    assert r["entropy"] == pytest.approx(3.0 * math.log(2.0), rel=1e-12)
    r = score([1, 2, 3, 4, 5], [-0.1, -0.2, -0.3, -5.0, -5.0], 3)             # statistics over the first result_len tokens only
    assert r["sum_logprobs"] == pytest.approx(-0.6, rel=1e-6) and r["avg_logprobs"] == pytest.approx(-0.2, rel=1e-6)
    assert r["score"] == pytest.approx(-0.2, rel=1e-6) and r["entropy"] == pytest.approx(math.log(3.0), rel=1e-12)
    # length_penalty > 0: score = sum / ((5 + len) / 6) ^ penalty  (whisper_sequence_score)
    r = score([1, 2, 3, 4], [-1.0] * 4, 4, length_penalty=1.0)
    assert r["score"] == pytest.approx(-4.0 / ((5.0 + 4.0) / 6.0), rel=1e-6)
