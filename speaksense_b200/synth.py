"""Synthetic ggml model files and synthetic audio.

There are no real Whisper weights on this box or on the GPU box (no network), so every
parity test and every bench line runs on a *synthetic* model written in the legacy ggml
``.bin`` container that the reference loads (``src/asr/whisper.rs:23`` ->
``WhisperContext::new_with_params``; file layout: SURVEY.md Appendix A.1).

Two weight families, both seeded:

* ``random``  - every tensor i.i.d. normal.  Logits are nearly flat, so this family is
  used for *teacher-forced* logits parity only.
* ``peaked``  - same, except the decoder positional embedding points at the token
  embedding of a scripted target token for every position and the residual-branch output
  projections are scaled down (GPT-2 style).  Greedy decoding then emits a scripted,
  grammar-valid transcript (timestamp pairs + ~24 text tokens per segment) with
  avg_logprob ~ 0 and high entropy, so the temperature-0 pass succeeds and the decode
  loop, the timestamp grammar and the segmenter are exercised end to end.  The layer
  stack still perturbs the logits by O(1), so encoder / cross-attention bugs remain visible
  in teacher-forced comparisons.

The product never imports this module; it is used by tests, bench.py and smoke().
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

GGML_MAGIC = 0x67676D6C
SAMPLE_RATE = 16000


@dataclass(frozen=True)
class HParams:
    n_vocab: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int
    n_mels: int
    ftype: int = 1

    def as_list(self):
        return [self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head,
                self.n_audio_layer, self.n_text_ctx, self.n_text_state, self.n_text_head,
                self.n_text_layer, self.n_mels, self.ftype]


SHAPES = {
    # public Whisper configs
    "tiny.en": HParams(51864, 1500, 384, 6, 4, 448, 384, 6, 4, 80),
    "tiny": HParams(51865, 1500, 384, 6, 4, 448, 384, 6, 4, 80),
    "base": HParams(51865, 1500, 512, 8, 6, 448, 512, 8, 6, 80),
    "small": HParams(51865, 1500, 768, 12, 12, 448, 768, 12, 12, 80),
    "medium": HParams(51865, 1500, 1024, 16, 24, 448, 1024, 16, 24, 80),
    "large-v3": HParams(51866, 1500, 1280, 20, 32, 448, 1280, 20, 32, 128),
    # large-v3 vocabulary / mel count at toy width: CPU-suite sized
    "micro-v3": HParams(51866, 1500, 256, 4, 2, 448, 256, 4, 2, 128),
    # large-v3 width with 2+2 layers: exercises every large-v3 kernel shape in seconds on CPU
    "large-v3-l2": HParams(51866, 1500, 1280, 20, 2, 448, 1280, 20, 2, 128),
    # large-v3-turbo (script/download-ggml-model.sh:49): the large-v3 encoder with a 4-layer decoder
    "large-v3-turbo": HParams(51866, 1500, 1280, 20, 32, 448, 1280, 20, 4, 128),
    # turbo-like asymmetry (more encoder than decoder layers) at toy width
    "micro-turbo": HParams(51866, 1500, 256, 4, 3, 448, 256, 4, 1, 128),
}


def special_tokens(n_vocab: int) -> dict:
    """Token ids derived from n_vocab exactly as SURVEY.md Appendix A.5 describes."""
    t = dict(eot=50256, sot=50257, translate=50357, transcribe=50358, solm=50359,
             prev=50360, nosp=50361, no_timestamps=50362, beg=50363)
    multilingual = n_vocab >= 51865
    if multilingual:
        t["eot"] += 1
        t["sot"] += 1
        dt = (n_vocab - 51765 - 1) - 98
        for k in ("translate", "transcribe", "solm", "prev", "nosp", "no_timestamps", "beg"):
            t[k] += dt
    t["multilingual"] = multilingual
    return t


# --------------------------------------------------------------------------------------
# mel filterbank (slaney scale, slaney norm) - what the real ggml files embed
# --------------------------------------------------------------------------------------
def mel_filters(n_mels: int, n_fft: int = 400, sr: int = SAMPLE_RATE) -> np.ndarray:
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        mel = 3.0 * f / 200.0
        log_reg = f >= 1000.0
        mel = np.where(log_reg, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4)), mel)
        return mel

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = 200.0 * m / 3.0
        log_reg = m >= 15.0
        f = np.where(log_reg, 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0)), f)
        return f

    n_bins = 1 + n_fft // 2
    fft_freqs = np.linspace(0, sr / 2, n_bins)
    mel_pts = np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2)
    hz_pts = mel_to_hz(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)  # [n_mels, n_bins]


# --------------------------------------------------------------------------------------
# vocabulary
# --------------------------------------------------------------------------------------
_CJK = ("的一是在不了有和人这中大为上个国我以要他时来用们生到作地于出就分对成会可主发年动同工也能下过子说产种面而方后多定行学法所"
        "民得经十三之进着等部度家电力里如水化高自二理起小物现实加量都两体制机当使点从业本去把性好应开它合还因由其些然前外天政四日那社义事平形相全表间"
        "样与关各重新线内数正心反你明看原又么利比或但质气第向道命此变条只没结解问意建月公无系军很情者最立代想已通并提直题党程展五果料象员革位入常文总次品式活设及管特件长求老头基资边流路级少图山统接知较将组见计别她手角期根论运农指几九区强放决西被干做必战先回则任取据处队南给色光门即保治北造百规热领七海口东导器压志世金增争济阶油思术极交受联什认六共权收证改清己美再采转更单风切打白教速花带安场身车例真务具万每目至达走积示议声报斗完类八离华名确才科张信马节话米整空元况今集温传土许步群广石记需段研界拉林律叫且究观越织装影算低持音众书布复容儿须际商非验连断深难近矿千周委素技备半办青省列习响约支般史感劳便团往酸历市克何除消构府称太准精值号率族维划选标写存候毛亲快效斯院查江型眼王按格养易置派层片始却专状育厂京识适属圆包火住调满县局照参红细引听该铁价严"
        "吗呢啊哇太订阅点赞打赏请")


def _gpt2_byte_order():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    rest = [b for b in range(256) if b not in bs]
    return bs + rest  # id -> byte value; id 220 == 0x20 (" ")


def make_vocab(n_vocab_file: int, seed: int = 7):
    """Deterministic pseudo vocabulary: ids 0..255 are single bytes in GPT-2 order (so id 220
    is " " as whisper.cpp's suppress_blank expects and ids >= 94 are partial UTF-8 like the
    real byte-level BPE), the rest alternate ASCII words and CJK characters."""
    order = _gpt2_byte_order()
    rng = np.random.default_rng(seed)
    toks = []
    letters = "abcdefghijklmnopqrstuvwxyz"
    for i in range(n_vocab_file):
        if i < 256:
            toks.append(bytes([order[i]]))
        elif i % 3 == 0:
            toks.append(_CJK[(i * 7919) % len(_CJK)].encode("utf-8"))
        else:
            n = 2 + int(rng.integers(0, 6))
            w = "".join(letters[int(c)] for c in rng.integers(0, 26, size=n))
            toks.append(((" " if i % 2 else "") + w).encode("utf-8"))
    return toks


# --------------------------------------------------------------------------------------
# scripted transcript for the ``peaked`` family
# --------------------------------------------------------------------------------------
def scripted_targets(hp: HParams, seed: int, seg_tokens: int = 24, seg_ticks: int = 296):
    """target[p] = token the decoder should emit after consuming position p.

    Aligned to the no-context prompt ([sot] for *.en, [sot, lang, transcribe] otherwise):
    ``<|0.00|> text*seg <|t1|><|t1|> text*seg <|t2|><|t2|> ...`` with t_k = k*seg_ticks*0.02 s.
    Text ids are >= 256 (valid UTF-8 in make_vocab) and all distinct inside a window so the
    entropy gate (>= 2.4 over the last 32 tokens) passes."""
    st = special_tokens(hp.n_vocab)
    rng = np.random.default_rng(seed + 101)
    p0 = (3 if st["multilingual"] else 1) - 1
    text_pool = rng.permutation(np.arange(256, st["eot"] - 1))
    tgt = rng.integers(256, st["eot"] - 1, size=hp.n_text_ctx)  # filler for p < p0
    p, k, ti = p0, 0, 0
    tgt[p] = st["beg"] + 0
    p += 1
    while p < hp.n_text_ctx:
        for _ in range(seg_tokens):
            if p >= hp.n_text_ctx:
                break
            tgt[p] = text_pool[ti % len(text_pool)]
            ti += 1
            p += 1
        k += 1
        ts = st["beg"] + min(k * seg_ticks, 1500)
        for _ in range(2):
            if p < hp.n_text_ctx:
                tgt[p] = ts
                p += 1
    return tgt.astype(np.int64)


# --------------------------------------------------------------------------------------
# file writer
# --------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, path):
        self.f = open(path, "wb", buffering=1 << 22)
        self.n_tensors = 0

    def header(self, hp: HParams, filters: np.ndarray, vocab):
        f = self.f
        f.write(struct.pack("<I", GGML_MAGIC))
        f.write(struct.pack("<11i", *hp.as_list()))
        f.write(struct.pack("<2i", filters.shape[0], filters.shape[1]))
        f.write(np.ascontiguousarray(filters, dtype="<f4").tobytes())
        f.write(struct.pack("<i", len(vocab)))
        for t in vocab:
            f.write(struct.pack("<I", len(t)))
            f.write(t)

    def tensor(self, name: str, arr: np.ndarray, f16: bool):
        """numpy shape is row-major [.., ne1, ne0]; ggml stores dims reversed (ne0 fastest)."""
        f = self.f
        nb = name.encode()
        arr = np.ascontiguousarray(arr)
        f.write(struct.pack("<3i", arr.ndim, len(nb), 1 if f16 else 0))
        for d in reversed(arr.shape):
            f.write(struct.pack("<i", d))
        f.write(nb)
        f.write(arr.astype("<f2" if f16 else "<f4", copy=False).tobytes())
        self.n_tensors += 1

    def close(self):
        self.f.close()


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def write_model(path: str, shape: str = "tiny.en", family: str = "peaked", seed: int = 0,
                seg_tokens: int = 24, seg_ticks: int = 296, hparams: HParams | None = None) -> dict:
    """Write a synthetic legacy-ggml Whisper file.  Returns a small dict of metadata
    (hparams, special tokens, scripted targets) for the tests."""
    hp = hparams or SHAPES[shape]
    # "soft<a>" = the peaked script with a target logit of only <a> (peaked: >= 30): with a ~ 10 the greedy tokens at temperature 0
    # have log-probabilities below logprob_thold (-1.0), so whisper_full walks the temperature ladder (best_of sampled decoders at
    # t > 0) - on a distribution that is still peaked enough for token-for-token comparison of the draws
    peak = None
    if family.startswith("soft"):
        peak = float(family[4:]); family = "peaked"
    assert family in ("random", "peaked")
    rng = np.random.default_rng(seed)
    st = special_tokens(hp.n_vocab)
    de, dd = hp.n_audio_state, hp.n_text_state
    n_vocab_file = 50257 if st["multilingual"] else 50256
    vocab = make_vocab(n_vocab_file)

    def nrm(shape_, std):
        return rng.standard_normal(shape_, dtype=np.float32) * np.float32(std)

    tmp = path + ".tmp%d" % os.getpid()
    w = _Writer(tmp)
    w.header(hp, mel_filters(hp.n_mels), vocab)

    def ln(prefix, d):
        w.tensor(prefix + ".weight", 1.0 + nrm((d,), 0.05), False)
        w.tensor(prefix + ".bias", nrm((d,), 0.05), False)

    def attn(prefix, d, out_std):
        w.tensor(prefix + ".query.weight", nrm((d, d), 0.02), True)
        w.tensor(prefix + ".query.bias", nrm((d,), 0.02), False)
        w.tensor(prefix + ".key.weight", nrm((d, d), 0.02), True)
        w.tensor(prefix + ".value.weight", nrm((d, d), 0.02), True)
        w.tensor(prefix + ".value.bias", nrm((d,), 0.02), False)
        w.tensor(prefix + ".out.weight", nrm((d, d), out_std), True)
        w.tensor(prefix + ".out.bias", nrm((d,), out_std), False)

    def mlp(prefix, d, out_std):
        w.tensor(prefix + ".0.weight", nrm((4 * d, d), 0.02), True)
        w.tensor(prefix + ".0.bias", nrm((4 * d,), 0.02), False)
        w.tensor(prefix + ".2.weight", nrm((d, 4 * d), out_std), True)
        w.tensor(prefix + ".2.bias", nrm((d,), out_std), False)

    # ---- encoder
    enc_out = 0.02 / np.sqrt(2.0 * hp.n_audio_layer)
    w.tensor("encoder.positional_embedding", sinusoids(hp.n_audio_ctx, de), False)
    w.tensor("encoder.conv1.weight", nrm((de, hp.n_mels, 3), 0.05), True)
    w.tensor("encoder.conv1.bias", nrm((de, 1), 0.02), False)
    w.tensor("encoder.conv2.weight", nrm((de, de, 3), 0.02), True)
    w.tensor("encoder.conv2.bias", nrm((de, 1), 0.02), False)
    for i in range(hp.n_audio_layer):
        b = "encoder.blocks.%d" % i
        ln(b + ".attn_ln", de)
        attn(b + ".attn", de, enc_out)
        ln(b + ".mlp_ln", de)
        mlp(b + ".mlp", de, enc_out)
    ln("encoder.ln_post", de)

    # ---- decoder
    dec_out = 0.02 / np.sqrt(3.0 * hp.n_text_layer) if family == "peaked" else 0.02
    # peaked: target logit = |e| * sqrt(d) * sqrt(d) = emb_std * d must clear logsumexp over the
    # vocabulary (~ln(51866) + var/2) by a wide margin at temperature 0 for every model width
    emb = nrm((hp.n_vocab, dd), (peak / dd if peak is not None else max(0.02, 30.0 / dd)) if family == "peaked" else 0.02)
    targets = None
    if family == "peaked":
        targets = scripted_targets(hp, seed, seg_tokens, seg_ticks)
        e16 = emb.astype(np.float16).astype(np.float32)[targets]
        e16 /= np.linalg.norm(e16, axis=1, keepdims=True)
        pos = (e16 * np.sqrt(dd)).astype(np.float32)
    else:
        pos = nrm((hp.n_text_ctx, dd), 0.02)
    w.tensor("decoder.positional_embedding", pos, False)
    w.tensor("decoder.token_embedding.weight", emb, True)
    del emb
    for i in range(hp.n_text_layer):
        b = "decoder.blocks.%d" % i
        ln(b + ".attn_ln", dd)
        attn(b + ".attn", dd, dec_out)
        ln(b + ".cross_attn_ln", dd)
        attn(b + ".cross_attn", dd, dec_out)
        ln(b + ".mlp_ln", dd)
        mlp(b + ".mlp", dd, dec_out)
    ln("decoder.ln", dd)
    w.close()
    os.replace(tmp, path)
    return {"hparams": hp, "special": st, "targets": targets, "n_tensors": w.n_tensors,
            "vocab": vocab, "path": path}


def read_model(path: str) -> dict:
    """Parse a legacy ggml Whisper file back into numpy (tests / golden tooling).  Returns
    {"hparams": HParams, "filters": [n_mel, 201] f32, "vocab": [bytes], "tensors": {name: ndarray}}
    with tensors in numpy (row-major, reversed ggml dims) order and their stored dtype."""
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    magic, = struct.unpack_from("<I", buf, o); o += 4
    assert magic == GGML_MAGIC, hex(magic)
    hp = HParams(*struct.unpack_from("<11i", buf, o)); o += 44
    n_mel, n_fft = struct.unpack_from("<2i", buf, o); o += 8
    filters = np.frombuffer(buf, "<f4", n_mel * n_fft, o).reshape(n_mel, n_fft).copy(); o += 4 * n_mel * n_fft
    n_tok, = struct.unpack_from("<i", buf, o); o += 4
    vocab = []
    for _ in range(n_tok):
        ln, = struct.unpack_from("<I", buf, o); o += 4
        vocab.append(buf[o:o + ln]); o += ln
    tensors = {}
    while o < len(buf):
        n_dims, nlen, tt = struct.unpack_from("<3i", buf, o); o += 12
        ne = struct.unpack_from("<%di" % n_dims, buf, o); o += 4 * n_dims
        name = buf[o:o + nlen].decode(); o += nlen
        cnt = int(np.prod(ne))
        dt = "<f2" if tt == 1 else "<f4"
        tensors[name] = np.frombuffer(buf, dt, cnt, o).reshape(tuple(reversed(ne))).copy()
        o += cnt * (2 if tt == 1 else 4)
    return {"hparams": hp, "filters": filters, "vocab": vocab, "tensors": tensors}


# --------------------------------------------------------------------------------------
# ggml block-quantised tensor types (QK = 32), as written by whisper.cpp's `quantize` tool
# (the reference downloads such files: script/download-ggml-model.sh:28-51, e.g. large-v3-q5_0).
# --------------------------------------------------------------------------------------
QTYPES = {   # name: (GGML_TYPE in the tensor header, GGML_FTYPE in hparams.ftype, bytes per block of 32)
    "q4_0": (2, 2, 18), "q4_1": (3, 3, 20), "q5_0": (6, 8, 22), "q5_1": (7, 9, 24), "q8_0": (8, 7, 34)}
GGML_QNT_VERSION_FACTOR = 1000      # hparams.ftype = ftype + 2 * 1000 for quantisation format version 2
QUANT_SKIP = ("encoder.conv1.bias", "encoder.conv2.bias", "encoder.positional_embedding", "decoder.positional_embedding")


def _round_away(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def quantize_blocks(x: np.ndarray, qtype: str) -> bytes:
    """ggml quantize_row_<qtype>_reference over x (f32, size % 32 == 0) -> raw block bytes"""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 32)
    nb = x.shape[0]
    if qtype == "q8_0":
        d = np.abs(x).max(axis=1) / np.float32(127.0)
        inv = np.where(d > 0, np.float32(1.0) / np.where(d > 0, d, 1), 0).astype(np.float32)
        q = _round_away(x * inv[:, None]).astype(np.int8)
        out = np.zeros((nb, 34), np.uint8)
        out[:, :2] = d.astype("<f2").view(np.uint8).reshape(nb, 2)
        out[:, 2:] = q.view(np.uint8)
        return out.tobytes()
    sym = qtype in ("q4_0", "q5_0")
    levels = 16 if qtype.startswith("q4") else 32
    if sym:
        idx = np.abs(x).argmax(axis=1)
        mx = x[np.arange(nb), idx]
        d = mx / np.float32(-(levels // 2))
        mn = None
    else:
        mn, mx = x.min(axis=1), x.max(axis=1)
        d = (mx - mn) / np.float32(levels - 1)
    inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), 0).astype(np.float32)
    if sym:
        q = np.minimum(levels - 1, (x * inv[:, None] + np.float32(levels // 2 + 0.5)).astype(np.int32).astype(np.int8)).astype(np.uint8)
    else:
        q = ((x - mn[:, None]) * inv[:, None] + np.float32(0.5)).astype(np.int32).astype(np.uint8)
        q = np.minimum(q, levels - 1).astype(np.uint8)
    lo, hi = q[:, :16], q[:, 16:]
    qs = ((lo & 0x0F) | ((hi & 0x0F) << 4)).astype(np.uint8)
    head = [d.astype("<f2").view(np.uint8).reshape(nb, 2)]
    if not sym:
        head.append(mn.astype("<f2").view(np.uint8).reshape(nb, 2))
    if levels == 32:
        qh = np.zeros(nb, np.uint32)
        for j in range(16):
            qh |= ((lo[:, j].astype(np.uint32) & 0x10) >> 4) << j
            qh |= ((hi[:, j].astype(np.uint32) & 0x10) >> 4) << (j + 16)
        head.append(qh.astype("<u4").view(np.uint8).reshape(nb, 4))
    return np.concatenate(head + [qs], axis=1).tobytes()


def dequantize_blocks(raw: bytes, qtype: str, count: int) -> np.ndarray:
    """ggml dequantize_row_<qtype> -> f32 (the specification our loaders follow)"""
    bs = QTYPES[qtype][2]
    b = np.frombuffer(raw, np.uint8, (count // 32) * bs).reshape(-1, bs)
    d = b[:, 0:2].copy().view("<f2").astype(np.float32)[:, 0]
    o = 2
    m = None
    if qtype in ("q4_1", "q5_1"):
        m = b[:, 2:4].copy().view("<f2").astype(np.float32)[:, 0]; o = 4
    if qtype == "q8_0":
        q = b[:, 2:].copy().view(np.int8).astype(np.float32)
        return (q * d[:, None]).reshape(-1)
    qh = None
    if qtype.startswith("q5"):
        qh = b[:, o:o + 4].copy().view("<u4")[:, 0]; o += 4
    qs = b[:, o:o + 16]
    lo, hi = (qs & 0x0F).astype(np.int32), (qs >> 4).astype(np.int32)
    if qh is not None:
        j = np.arange(16)
        lo |= (((qh[:, None] >> j[None, :]) & 1) << 4).astype(np.int32)
        hi |= (((qh[:, None] >> (j[None, :] + 16)) & 1) << 4).astype(np.int32)
    q = np.concatenate([lo, hi], axis=1).astype(np.float32)
    if m is None:
        return ((q - np.float32(8 if qtype == "q4_0" else 16)) * d[:, None]).reshape(-1)
    return (q * d[:, None] + m[:, None]).reshape(-1)


def quantize_model(src: str, dst: str, qtype: str, dequantized_f16_twin: str | None = None) -> None:
    """Re-write the f16 synthetic file `src` the way whisper.cpp's quantize tool does: every 2-D tensor (except the
    skip list) becomes `qtype` blocks, everything else is copied.  With `dequantized_f16_twin` also writes an f16 file
    whose 2-D tensors hold dequantize(quantize(w)) rounded to f16: an engine that dequantises at load must give
    bit-identical results on both files."""
    ttype, ftype, _ = QTYPES[qtype]
    m = read_model(src)
    hp: HParams = m["hparams"]
    outs = [(dst, True)] + ([(dequantized_f16_twin, False)] if dequantized_f16_twin else [])
    for path, quantised in outs:
        tmp = path + ".tmp%d" % os.getpid()
        w = _Writer(tmp)
        vals = hp.as_list()
        if quantised:
            vals[-1] = ftype + 2 * GGML_QNT_VERSION_FACTOR
        hq = HParams(*vals)
        w.header(hq, m["filters"], m["vocab"])
        for name, arr in m["tensors"].items():
            is_f16 = arr.dtype == np.float16
            if arr.ndim == 2 and name not in QUANT_SKIP and arr.shape[1] % 32 == 0 and not (arr.shape[0] == 1 or arr.shape[1] == 1):
                raw = quantize_blocks(arr.astype(np.float32), qtype)
                if quantised:
                    nb = name.encode()
                    w.f.write(struct.pack("<3i", 2, len(nb), ttype))
                    w.f.write(struct.pack("<2i", arr.shape[1], arr.shape[0]))
                    w.f.write(nb)
                    w.f.write(raw)
                    w.n_tensors += 1
                else:
                    w.tensor(name, dequantize_blocks(raw, qtype, arr.size).reshape(arr.shape), True)
            else:
                w.tensor(name, arr, is_f16)
        w.close()
        os.replace(tmp, path)


def write_twin_from_raw(qpath: str, twin_path: str, dequantize) -> None:
    """Copy the (possibly block-quantised) file `qpath` to `twin_path` with every quantised tensor replaced by the f16 rounding
    of dequantize(raw_block_bytes, numpy_shape) -> f32 array; header (hparams incl. ftype), filters, vocabulary and all other
    tensors are copied byte for byte.  tests/test_quantized.py passes gguf.quants.dequantize here."""
    buf = open(qpath, "rb").read()
    o = 4 + 44
    n_mel, n_fft = struct.unpack_from("<2i", buf, o); o += 8 + 4 * n_mel * n_fft
    n_tok, = struct.unpack_from("<i", buf, o); o += 4
    for _ in range(n_tok):
        ln, = struct.unpack_from("<I", buf, o); o += 4 + ln
    by_type = {t: bs for t, _, bs in QTYPES.values()}
    tmp = twin_path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(buf[:o])
        while o < len(buf):
            n_dims, nlen, tt = struct.unpack_from("<3i", buf, o)
            ne = struct.unpack_from("<%di" % n_dims, buf, o + 12)
            name = buf[o + 12 + 4 * n_dims:o + 12 + 4 * n_dims + nlen]
            head = 12 + 4 * n_dims + nlen
            cnt = int(np.prod(ne))
            if tt in by_type:
                nbytes = cnt // 32 * by_type[tt]
                shape = tuple(reversed(ne))
                vals = np.asarray(dequantize(buf[o + head:o + head + nbytes], shape), np.float32).reshape(shape)
                f.write(struct.pack("<3i", n_dims, nlen, 1)); f.write(struct.pack("<%di" % n_dims, *ne)); f.write(name)
                f.write(vals.astype("<f2").tobytes())
            else:
                nbytes = cnt * (2 if tt == 1 else 4)
                f.write(buf[o:o + head + nbytes])
            o += head + nbytes
    os.replace(tmp, twin_path)


def ensure_model(path: str, **kw) -> str:
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        write_model(path, **kw)
    return path


# --------------------------------------------------------------------------------------
# audio
# --------------------------------------------------------------------------------------
def synth_audio(n_samples: int = 30 * SAMPLE_RATE, seed: int = 1234) -> np.ndarray:
    """Sum of 5 chirps (100-4000 Hz, random phase) x 4 Hz envelope x 0.3 + N(0, 0.01^2),
    clipped to [-1, 1]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    dur = max(n_samples / SAMPLE_RATE, 1e-3)
    x = np.zeros(n_samples, dtype=np.float64)
    for _ in range(5):
        f0, f1 = np.sort(rng.uniform(100.0, 4000.0, size=2))
        ph = rng.uniform(0, 2 * np.pi)
        x += np.sin(ph + 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / dur))
    env = 0.5 * (1.0 + np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
    x = 0.3 * (x / 5.0) * env * 3.0 + rng.normal(0.0, 0.01, size=n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--shape", default="tiny.en", choices=sorted(SHAPES))
    ap.add_argument("--family", default="peaked", help="random | peaked | soft<target logit>, e.g. soft10")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    m = write_model(a.path, a.shape, a.family, a.seed)
    print(a.path, m["n_tensors"], "tensors", os.path.getsize(a.path) / 1e6, "MB")
