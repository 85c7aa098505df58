"""Pin the CPU oracle against an independent implementation and write tests/golden/*.npz.

The reference keeps no golden vectors for the transcribe path ("parity
unpinned") and its arithmetic (whisper.cpp via crates.io) is not in /root/reference.  The only
independent implementation of Whisper available in this container is HuggingFace `transformers`
(v5.5, torch CPU fp32).  This script loads the SAME synthetic ggml tensors into
WhisperForConditionalGeneration, runs feature extraction / encoder / teacher-forced decoder, and
stores small slices of the HF outputs as fixtures.  tests/test_oracle_golden.py then holds the C
oracle to those fixtures (tolerances account for whisper.cpp's f16-rounded matmul inputs, which HF
fp32 does not have).  Expected, documented deltas: STFT end padding (reflect vs
zeros: last frame only), fp32 vs f16-rounded activations.

Run:  python tools/make_golden.py            (needs transformers + torch; CPU only)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200 import synth  # noqa: E402

CASES = [  # (fixture name, shape, family, model seed, audio seed, language)
    ("tiny_en_peaked", "tiny.en", "peaked", 0, 1234, None),
    ("micro_v3_random", "micro-v3", "random", 3, 1235, "zh"),
]
N_STEPS = 12


def hf_model(model):
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    hp, T = model["hparams"], model["tensors"]
    cfg = WhisperConfig(
        vocab_size=hp.n_vocab, num_mel_bins=hp.n_mels, d_model=hp.n_audio_state,
        encoder_layers=hp.n_audio_layer, encoder_attention_heads=hp.n_audio_head,
        decoder_layers=hp.n_text_layer, decoder_attention_heads=hp.n_text_head,
        encoder_ffn_dim=4 * hp.n_audio_state, decoder_ffn_dim=4 * hp.n_text_state,
        max_source_positions=hp.n_audio_ctx, max_target_positions=hp.n_text_ctx,
        activation_function="gelu_pytorch_tanh", dropout=0.0, attention_dropout=0.0,
        activation_dropout=0.0, scale_embedding=False, pad_token_id=0, bos_token_id=0, eos_token_id=0,
        decoder_start_token_id=0, suppress_tokens=None, begin_suppress_tokens=None)
    hf = WhisperForConditionalGeneration(cfg).eval()
    sd = {}

    def put(k, name):
        sd[k] = torch.from_numpy(T[name].astype(np.float32))

    put("model.encoder.conv1.weight", "encoder.conv1.weight")
    sd["model.encoder.conv1.bias"] = torch.from_numpy(T["encoder.conv1.bias"].astype(np.float32).reshape(-1))
    put("model.encoder.conv2.weight", "encoder.conv2.weight")
    sd["model.encoder.conv2.bias"] = torch.from_numpy(T["encoder.conv2.bias"].astype(np.float32).reshape(-1))
    put("model.encoder.embed_positions.weight", "encoder.positional_embedding")
    put("model.encoder.layer_norm.weight", "encoder.ln_post.weight")
    put("model.encoder.layer_norm.bias", "encoder.ln_post.bias")
    put("model.decoder.embed_positions.weight", "decoder.positional_embedding")
    put("model.decoder.embed_tokens.weight", "decoder.token_embedding.weight")
    put("proj_out.weight", "decoder.token_embedding.weight")
    put("model.decoder.layer_norm.weight", "decoder.ln.weight")
    put("model.decoder.layer_norm.bias", "decoder.ln.bias")

    def attn(dst, src):
        for a, b in (("q_proj", "query"), ("k_proj", "key"), ("v_proj", "value"), ("out_proj", "out")):
            put(f"{dst}.{a}.weight", f"{src}.{b}.weight")
            if b != "key":
                put(f"{dst}.{a}.bias", f"{src}.{b}.bias")

    def ln(dst, src):
        put(f"{dst}.weight", f"{src}.weight")
        put(f"{dst}.bias", f"{src}.bias")

    for i in range(hp.n_audio_layer):
        d, s = f"model.encoder.layers.{i}", f"encoder.blocks.{i}"
        attn(d + ".self_attn", s + ".attn"); ln(d + ".self_attn_layer_norm", s + ".attn_ln")
        ln(d + ".final_layer_norm", s + ".mlp_ln")
        for a, b in (("fc1", "mlp.0"), ("fc2", "mlp.2")):
            put(f"{d}.{a}.weight", f"{s}.{b}.weight"); put(f"{d}.{a}.bias", f"{s}.{b}.bias")
    for i in range(hp.n_text_layer):
        d, s = f"model.decoder.layers.{i}", f"decoder.blocks.{i}"
        attn(d + ".self_attn", s + ".attn"); ln(d + ".self_attn_layer_norm", s + ".attn_ln")
        attn(d + ".encoder_attn", s + ".cross_attn"); ln(d + ".encoder_attn_layer_norm", s + ".cross_attn_ln")
        ln(d + ".final_layer_norm", s + ".mlp_ln")
        for a, b in (("fc1", "mlp.0"), ("fc2", "mlp.2")):
            put(f"{d}.{a}.weight", f"{s}.{b}.weight"); put(f"{d}.{a}.bias", f"{s}.{b}.bias")
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "k_proj.bias" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


def hf_log_mel(pcm, filters):
    """transformers.WhisperFeatureExtractor with the model file's own filterbank."""
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=filters.shape[0], sampling_rate=16000, hop_length=160,
                                 chunk_length=30, n_fft=400)
    fe.mel_filters = filters.T.astype(np.float64)      # [201, n_mels]
    out = fe(pcm, sampling_rate=16000, return_tensors="np")
    return out["input_features"][0]                    # [n_mels, 3000]


def prompt_tokens(hp, lang):
    st = synth.special_tokens(hp.n_vocab)
    if not st["multilingual"]:
        return [st["sot"]]
    langs = ["en", "zh", "de", "es", "ru", "ko", "fr", "ja"]
    return [st["sot"], st["sot"] + 1 + langs.index(lang or "en"), st["transcribe"]]


def main():
    import torch
    torch.set_grad_enabled(False)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, shape, family, mseed, aseed, lang in CASES:
        path = "/tmp/ss_golden_%s.bin" % name
        meta = synth.write_model(path, shape, family, mseed)
        model = synth.read_model(path)
        hp = model["hparams"]
        pcm = synth.synth_audio(seed=aseed)
        mel_hf = hf_log_mel(pcm, model["filters"])
        hf = hf_model(model)
        feats = torch.from_numpy(mel_hf[None].astype(np.float32))
        enc = hf.model.encoder(feats).last_hidden_state            # [1, 1500, d]
        if meta["targets"] is not None:
            p0 = len(prompt_tokens(hp, lang))
            forced = prompt_tokens(hp, lang) + [int(t) for t in meta["targets"][p0 - 1:p0 - 1 + N_STEPS]]
        else:
            rng = np.random.default_rng(99)
            forced = prompt_tokens(hp, lang) + [int(t) for t in rng.integers(256, 50000, size=N_STEPS)]
        ids = torch.tensor([forced])
        dec = hf.model.decoder(input_ids=ids, encoder_hidden_states=enc).last_hidden_state
        logits = hf.proj_out(dec)[0].numpy()                         # [n_tok, n_vocab]
        top = np.argsort(-logits, axis=1)[:, :8]
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            shape=shape, family=family, model_seed=mseed, audio_seed=aseed, language=lang or "",
            mel_slice=mel_hf[::4, ::50].astype(np.float32),          # frames 0,50,..: [n_mels/4, 60]
            mel_mean=np.float64(mel_hf[:, :2999].mean()), mel_max=np.float64(mel_hf.max()),
            enc_slice=enc[0, ::25, ::16].numpy().astype(np.float32), # [60, d/16]
            enc_abs_mean=np.float64(enc.abs().mean()),
            forced=np.array(forced, np.int32),
            logits_top_idx=top.astype(np.int32),
            logits_top_val=np.take_along_axis(logits, top, 1).astype(np.float32),
            logits_slice=logits[:, ::997].astype(np.float32),
            logits_lse=np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1))
        print(name, "mel", mel_hf.shape, "enc", tuple(enc.shape), "logits", logits.shape,
              "top1", top[:, 0][:6], "abs mean", float(enc.abs().mean()))
        os.remove(path)


if __name__ == "__main__":
    main()
