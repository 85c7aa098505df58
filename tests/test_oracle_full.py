"""whisper_full semantics of the oracle on scripted ("peaked") models: prompt, timestamp grammar,
segmentation, stream mode, short / ragged inputs, temperature fallback."""
import numpy as np
import pytest

from tests.rust_post import add_punctuation, is_promotional_text, post_process


@pytest.fixture(scope="module")
def tiny(oracle_mod, tiny_en_peaked):
    m = oracle_mod.OracleModel(tiny_en_peaked)
    yield m
    m.close()


def test_scripted_transcript(tiny, audio30):
    from speaksense_b200 import synth
    st = tiny.new_state()
    r = st.full(audio30, keep_logits=True)
    targets = synth.scripted_targets(synth.SHAPES["tiny.en"], 0)
    assert r["tokens"] == [int(t) for t in targets[:len(r["tokens"])]]       # greedy follows the script
    assert r["n_fallbacks"] == 0 and r["n_windows"] == 1 and r["n_decoded"] == len(r["tokens"])
    beg = tiny.token_id("beg")
    assert r["tokens"][0] == beg and r["tokens"][-1] == beg + 1480            # stops when the window is covered
    assert [(s["t0"], s["t1"]) for s in r["segments"]] == [(0, 592), (592, 1184), (1184, 1776), (1776, 2368), (2368, 2960)]
    text = b"".join(tiny.token_bytes(t) for t in r["tokens"][1:25])
    assert r["segments"][0]["text"] == text
    logits = st.kept_logits()
    assert logits.shape == (len(r["tokens"]), tiny.hparams["n_vocab"])
    assert (np.sort(logits, 1)[:, -1] - np.sort(logits, 1)[:, -2]).min() > 5      # wide top-1 margin by construction
    st.close()


def test_stream_mode_keeps_last_segment(tiny, audio30):
    st = tiny.new_state()
    r = st.full(audio30, stream_mode=True)
    out = post_process(r["segments"], True)
    assert len(out["segments"]) == 1 and out["segments"][0][2:] == (2368.0, 2960.0)
    assert out["full_text"] == out["segments"][0][0]
    full = post_process(r["segments"], False)
    assert len(full["segments"]) == 5 and full["full_text"] == "".join(s[0] for s in full["segments"])
    st.close()


@pytest.mark.parametrize("n", [0, 1, 15999, 16000 + 159, 32000])
def test_short_audio_yields_nothing_or_little(tiny, n):
    from speaksense_b200 import synth
    st = tiny.new_state()
    r = st.full(synth.synth_audio(max(n, 1), seed=3)[:n])
    if n < 16400:
        assert r["segments"] == [] and r["n_windows"] == 0            # <= ~1 s (100 mel frames): whisper returns without decoding
    else:
        assert r["n_windows"] == 1
    st.close()


def test_two_windows_carry_context(tiny):
    """45 s of audio: second window is prompted with [prev] + past tokens unless no_context (stream mode)."""
    from speaksense_b200 import synth
    pcm = synth.synth_audio(45 * 16000, seed=11)
    st = tiny.new_state()
    a = st.full(pcm, stream_mode=False)
    b = st.full(pcm, stream_mode=True)
    assert a["n_windows"] == 2 and b["n_windows"] == 2
    assert a["segments"][0]["t0"] == 0 and a["segments"][-1]["t1"] > 2960
    # no_context=true: the second window sees the same prompt as the first -> decodes the same script again
    n1 = 130
    assert b["tokens"][:n1] == a["tokens"][:n1]
    st.close()


def test_fallback_ladder_on_flat_logits(oracle_mod, micro_v3_random):
    """Random weights: t=0 fails the logprob gate, the sampled decoders (best_of=5) run at t>0."""
    from speaksense_b200 import synth
    m = oracle_mod.OracleModel(micro_v3_random)
    st = m.new_state()
    r = st.full(synth.synth_audio(3 * 16000, seed=5), language="zh", max_tokens=0)
    assert r["n_fallbacks"] >= 1
    assert r["n_decoded"] > len(r["tokens"])
    st.close()
    m.close()


def test_unknown_language(oracle_mod, micro_v3_random, audio30):
    m = oracle_mod.OracleModel(micro_v3_random)
    st = m.new_state()
    with pytest.raises(oracle_mod.OracleError):
        st.full(audio30[:32000], language="xx")
    st.close()
    m.close()


def test_rust_text_rules():
    assert add_punctuation("你好吗") == "你好吗？"
    assert add_punctuation("太好了") == "太好了！"
    assert add_punctuation("今天下雨") == "今天下雨 "
    assert add_punctuation("已经有了。") == "已经有了。"
    assert is_promotional_text("欢迎订阅我的频道") and not is_promotional_text("hello")


def test_sampler_matches_real_libstdcxx(oracle_mod, micro_v3_random, tmp_path):
    """whisper_sample_token at temperature > 0 draws from std::discrete_distribution<> driven by std::mt19937
    (SURVEY App. A.5).  The oracle restates both in C; here they are held to the real libstdc++ classes."""
    import ctypes as C
    import subprocess
    src = tmp_path / "sampler.cpp"
    src.write_text('''
#include <random>
#include <vector>
#include <cstdio>
int main(int argc, char **argv) {
    unsigned seed = 0; int n = 0, count = 0;
    if (scanf("%u %d %d", &seed, &n, &count) != 3) return 1;
    std::vector<float> p(n);
    for (int i = 0; i < n; i++) if (scanf("%f", &p[i]) != 1) return 1;
    std::mt19937 rng(seed);
    std::discrete_distribution<> dist(p.begin(), p.end());
    for (int i = 0; i < count; i++) printf("%d\\n", dist(rng));
    return 0;
}
''')
    exe = tmp_path / "sampler"
    subprocess.check_call(["g++", "-O1", "-o", str(exe), str(src)])
    m = oracle_mod.OracleModel(micro_v3_random)
    st = m.new_state()
    L = oracle_mod.lib()
    L.wo_probe_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(5)
    for seed, n in ((0, 7), (1, 1000), (12345, 51866)):
        p = rng.random(n).astype(np.float32) ** 8          # peaky, unnormalised (discrete_distribution normalises)
        p[rng.integers(0, n, size=max(1, n // 10))] = 0.0
        count = 200
        out = np.zeros(count, np.int32)
        L.wo_probe_sample(st.h, seed, p.ctypes.data, n, count, out.ctypes.data)
        text = "%d %d %d\n" % (seed, n, count) + " ".join("%.9g" % v for v in p) + "\n"
        ref = subprocess.run([str(exe)], input=text, capture_output=True, text=True, check=True).stdout.split()
        assert out.tolist() == [int(x) for x in ref], (seed, n)
    st.close(); m.close()
