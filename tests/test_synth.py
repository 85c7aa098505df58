"""Synthetic ggml writer / reader and audio generator (test + bench tooling)."""
import numpy as np

from speaksense_b200 import synth


def test_roundtrip_and_determinism(tmp_path):
    hp = synth.HParams(51865, 1500, 128, 2, 1, 448, 128, 2, 1, 80)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    ma = synth.write_model(a, family="peaked", seed=5, hparams=hp)
    synth.write_model(b, family="peaked", seed=5, hparams=hp)
    assert open(a, "rb").read() == open(b, "rb").read()
    m = synth.read_model(a)
    assert m["hparams"] == hp and m["filters"].shape == (80, 201) and len(m["vocab"]) == 50257
    t = m["tensors"]
    assert t["decoder.token_embedding.weight"].dtype == np.float16 and t["decoder.token_embedding.weight"].shape == (51865, 128)
    assert t["encoder.conv1.weight"].shape == (128, 80, 3) and t["encoder.conv1.bias"].dtype == np.float32
    assert "encoder.blocks.0.attn.key.bias" not in t and "decoder.blocks.0.cross_attn.key.bias" not in t
    assert ma["n_tensors"] == len(t)


def test_special_tokens_table():
    # SURVEY.md Appendix A.5
    en, v2, v3 = synth.special_tokens(51864), synth.special_tokens(51865), synth.special_tokens(51866)
    assert (en["eot"], en["sot"], en["transcribe"], en["beg"]) == (50256, 50257, 50358, 50363)
    assert (v2["eot"], v2["sot"], v2["transcribe"], v2["beg"]) == (50257, 50258, 50359, 50364)
    assert (v3["eot"], v3["sot"], v3["translate"], v3["transcribe"], v3["beg"]) == (50257, 50258, 50359, 50360, 50365)


def test_mel_filters_match_transformers():
    from transformers.audio_utils import mel_filter_bank
    for n in (80, 128):
        ref = mel_filter_bank(201, n, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T
        assert np.abs(synth.mel_filters(n) - ref).max() < 1e-6


def test_audio_is_seeded_and_bounded():
    a, b, c = synth.synth_audio(16000, 1), synth.synth_audio(16000, 1), synth.synth_audio(16000, 2)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a).max() <= 1.0 and a.std() > 0.05


def test_calculate_checksum_is_rust_default_hasher():
    """whisper.rs:225-234 (row a13): DefaultHasher = SipHash-1-3, key (0, 0), over the f32 bit patterns.  The generic
    SipHash routine is pinned by the reference vector of the SipHash paper (2-4 rounds, key 00..0f, message 00..0e)."""
    import numpy as np
    from speaksense_b200.asr import calculate_checksum, siphash
    k0, k1 = int.from_bytes(bytes(range(8)), "little"), int.from_bytes(bytes(range(8, 16)), "little")
    assert siphash(bytes(range(15)), k0, k1, 2, 4) == 0xa129ca6149be45e5
    assert siphash(b"", k0, k1, 2, 4) == 0x726fdb47dd0e0e31
    x = np.array([0.0, 1.0, -1.0, 0.5], np.float32)
    assert calculate_checksum(x) == siphash(x.tobytes(), 0, 0, 1, 3)
    assert calculate_checksum(x) != calculate_checksum(x[::-1].copy())
    assert calculate_checksum(np.zeros(0, np.float32)) == siphash(b"", 0, 0, 1, 3)
