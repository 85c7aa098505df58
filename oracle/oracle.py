"""ctypes binding of the CPU oracle (oracle/whisper_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (speaksense_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libwhisper_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("whisper_oracle.c", "whisper_oracle.h", "audio_oracle.c", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return _LIB_PATH


class WoHParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_text_ctx",
        "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype")]


class WoParams(C.Structure):
    _fields_ = [("language", C.c_char_p), ("tdrz_enable", C.c_int), ("no_context", C.c_int),
                ("single_segment", C.c_int), ("best_of", C.c_int), ("beam_size", C.c_int),
                ("temperature", C.c_float), ("temperature_inc", C.c_float), ("entropy_thold", C.c_float),
                ("logprob_thold", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
                ("suppress_blank", C.c_int), ("n_max_text_ctx", C.c_int), ("max_tokens", C.c_int),
                ("n_threads", C.c_int), ("keep_logits", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = _LIB_PATH
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.wo_load.restype = C.c_void_p
    L.wo_load.argtypes = [C.c_char_p]
    L.wo_free.argtypes = [C.c_void_p]
    L.wo_get_hparams.restype = C.POINTER(WoHParams)
    L.wo_get_hparams.argtypes = [C.c_void_p]
    L.wo_last_error.restype = C.c_char_p
    L.wo_token_id.argtypes = [C.c_void_p, C.c_char_p]
    L.wo_token_bytes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
    L.wo_set_threads.argtypes = [C.c_int]
    L.wo_default_params.argtypes = [C.POINTER(WoParams)]
    L.wo_log_mel.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                             C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.wo_free_buf.argtypes = [C.c_void_p]
    L.wo_state_new.restype = C.c_void_p
    L.wo_state_new.argtypes = [C.c_void_p]
    L.wo_state_free.argtypes = [C.c_void_p]
    L.wo_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.wo_encoder_out.restype = C.POINTER(C.c_float)
    L.wo_encoder_out.argtypes = [C.c_void_p]
    L.wo_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.wo_full.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(WoParams)]
    for f in ("wo_n_segments", "wo_n_result_tokens", "wo_n_fallbacks", "wo_n_decoded", "wo_n_windows",
              "wo_n_kept_logits"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.wo_segment_text.restype = C.c_char_p
    L.wo_segment_text.argtypes = [C.c_void_p, C.c_int]
    L.wo_segment_t0.restype = C.c_int64
    L.wo_segment_t0.argtypes = [C.c_void_p, C.c_int]
    L.wo_segment_t1.restype = C.c_int64
    L.wo_segment_t1.argtypes = [C.c_void_p, C.c_int]
    L.wo_segment_speaker_turn_next.argtypes = [C.c_void_p, C.c_int]
    L.wo_result_token.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.wo_min_sample_margin.restype = C.c_double
    L.wo_min_sample_margin.argtypes = [C.c_void_p]
    L.wo_n_draws.restype = C.c_long
    L.wo_n_draws.argtypes = [C.c_void_p]
    L.wo_kept_logits.restype = C.POINTER(C.c_float)
    L.wo_kept_logits.argtypes = [C.c_void_p, C.c_int]
    _lib = L
    return L


def beam_assign(cands, live, i):
    """One beam-search step's candidate assignment (test probe of beam_assign in whisper_oracle.c).  cands: list of
    (token ids, sum_logprobs_all, decoder index) in the order the decoders produced them; live[j]: decoder j still running;
    i: index of the token being sampled.  Returns, per decoder, the index into `cands` of the candidate it continues with (-1)."""
    L = lib()
    n = len(cands)
    max_len = max([len(c[0]) for c in cands] + [1])
    ids = np.zeros((max(n, 1), max_len), np.int32)
    for k, c in enumerate(cands):
        ids[k, :len(c[0])] = c[0]
    lens = np.asarray([len(c[0]) for c in cands] + ([] if n else [0]), np.int32)
    sums = np.asarray([c[1] for c in cands] + ([] if n else [0.0]), np.float64)
    dec = np.asarray([c[2] for c in cands] + ([] if n else [0]), np.int32)
    lv = np.asarray([1 if x else 0 for x in live], np.int32)
    out = np.full(len(live), -2, np.int32)
    L.wo_probe_beam_assign.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.wo_probe_beam_assign.restype = C.c_int
    rc = L.wo_probe_beam_assign(ids.ctypes.data, lens.ctypes.data, max_len, sums.ctypes.data, dec.ctypes.data, n, lv.ctypes.data, len(live), i,
                                out.ctypes.data)
    if rc:
        raise RuntimeError("wo_probe_beam_assign rc=%d" % rc)
    return out.tolist()


class OracleError(RuntimeError):
    pass


class OracleModel:
    def __init__(self, path: str):
        self.L = lib()
        self.h = self.L.wo_load(path.encode())
        if not self.h:
            raise OracleError(self.L.wo_last_error().decode())
        hp = self.L.wo_get_hparams(self.h).contents
        self.hparams = {n: getattr(hp, n) for n, _ in WoHParams._fields_}

    def close(self):
        if self.h:
            self.L.wo_free(self.h)
            self.h = None

    def token_id(self, name: str) -> int:
        return self.L.wo_token_id(self.h, name.encode())

    def token_bytes(self, i: int) -> bytes:
        p = C.c_char_p()
        n = self.L.wo_token_bytes(self.h, i, C.byref(p))
        return C.string_at(p, n) if n >= 0 else b""

    def log_mel(self, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        out = C.c_void_p()
        n_len, n_org = C.c_int(), C.c_int()
        self.L.wo_log_mel(self.h, pcm.ctypes.data, pcm.size, C.byref(out), C.byref(n_len), C.byref(n_org))
        n_mels = self.hparams["n_mels"]
        arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(n_mels, n_len.value)).copy()
        self.L.wo_free_buf(out)
        return arr, n_len.value, n_org.value

    def new_state(self) -> "OracleState":
        return OracleState(self)


class OracleState:
    def __init__(self, model: OracleModel):
        self.m = model
        self.L = model.L
        self.h = self.L.wo_state_new(model.h)

    def close(self):
        if self.h:
            self.L.wo_state_free(self.h)
            self.h = None

    def encode(self, mel: np.ndarray, seek: int = 0) -> np.ndarray:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        self.L.wo_encode(self.h, mel.ctypes.data, mel.shape[1], seek)
        hp = self.m.hparams
        p = self.L.wo_encoder_out(self.h)
        return np.ctypeslib.as_array(p, shape=(hp["n_audio_ctx"], hp["n_audio_state"])).copy()

    def decode(self, tokens, n_past: int, seq: int = 0) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.m.hparams["n_vocab"], dtype=np.float32)
        rc = self.L.wo_decode(self.h, seq, t.ctypes.data, t.size, n_past, out.ctypes.data)
        if rc:
            raise OracleError(self.L.wo_last_error().decode())
        return out

    def full(self, pcm: np.ndarray, language=None, stream_mode=False, speaker_diarization=False,
             keep_logits=False, n_threads=16, beam_size=0, **over):
        """== reference transcribe_with_state up to (not including) the Rust post-processing:
        build_params (whisper.rs:131-173) + overrides (:60-71) + state.full (:75)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        P = WoParams()
        self.L.wo_default_params(C.byref(P))
        self._lang = language.encode() if language else None
        P.language = self._lang
        P.tdrz_enable = int(speaker_diarization)
        if stream_mode:
            P.single_segment = 0
            P.no_context = 1
        P.keep_logits = int(keep_logits)
        P.n_threads = n_threads
        P.beam_size = beam_size
        for k, v in over.items():
            setattr(P, k, v)
        rc = self.L.wo_full(self.h, pcm.ctypes.data, pcm.size, C.byref(P))
        if rc:
            raise OracleError("wo_full rc=%d: %s" % (rc, self.L.wo_last_error().decode()))
        segs = []
        for i in range(self.L.wo_n_segments(self.h)):
            segs.append(dict(text=self.L.wo_segment_text(self.h, i), t0=self.L.wo_segment_t0(self.h, i),
                             t1=self.L.wo_segment_t1(self.h, i),
                             speaker_turn_next=bool(self.L.wo_segment_speaker_turn_next(self.h, i))))
        toks, plogs = [], []
        p, pl = C.c_float(), C.c_float()
        for i in range(self.L.wo_n_result_tokens(self.h)):
            toks.append(self.L.wo_result_token(self.h, i, C.byref(p), C.byref(pl)))
            plogs.append(pl.value)
        return dict(segments=segs, tokens=toks, plogs=plogs, n_fallbacks=self.L.wo_n_fallbacks(self.h),
                    n_decoded=self.L.wo_n_decoded(self.h), n_windows=self.L.wo_n_windows(self.h),
                    n_draws=self.L.wo_n_draws(self.h), min_sample_margin=self.L.wo_min_sample_margin(self.h))

    def process_logits(self, ids, raw, has_ts=False, seek_delta=0, temperature=0.0, **over) -> np.ndarray:
        """whisper_process_logits probe: filtered logits (-inf = masked) for the history `ids` of sampled tokens"""
        P = WoParams()
        self.L.wo_default_params(C.byref(P))
        for k, v in over.items():
            setattr(P, k, v)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        out = np.empty_like(raw)
        self.L.wo_probe_process_logits.argtypes = [C.c_void_p, C.POINTER(WoParams), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_float, C.c_void_p]
        rc = self.L.wo_probe_process_logits(self.h, C.byref(P), ids.ctypes.data, ids.size, int(has_ts), int(seek_delta),
                                            raw.ctypes.data, temperature, out.ctypes.data)
        if rc:
            raise OracleError("wo_probe_process_logits rc=%d" % rc)
        return out

    def bookkeeping(self, ids, seek=0, seek_end=3000, n_max=220, **over) -> dict:
        """whisper_full's per-token decoder bookkeeping replayed over `ids` (test probe)"""
        P = WoParams()
        self.L.wo_default_params(C.byref(P))
        for k, v in over.items():
            setattr(P, k, v)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = (C.c_int * 6)()
        self.L.wo_probe_bookkeeping.argtypes = [C.c_void_p, C.POINTER(WoParams), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.L.wo_probe_bookkeeping(self.h, C.byref(P), ids.ctypes.data, ids.size, seek, seek_end, n_max, out)
        return dict(failed=bool(out[0]), completed=bool(out[1]), result_len=out[2], seek_delta=out[3], has_ts=bool(out[4]), steps=out[5])

    def score(self, ids, plogs, result_len, **over) -> dict:
        """whisper_sequence_score + the entropy statistic of the fallback gate (test probe)"""
        P = WoParams()
        self.L.wo_default_params(C.byref(P))
        for k, v in over.items():
            setattr(P, k, v)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        plogs = np.ascontiguousarray(plogs, dtype=np.float32)
        out = (C.c_double * 4)()
        self.L.wo_probe_score.argtypes = [C.c_void_p, C.POINTER(WoParams), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self.L.wo_probe_score(self.h, C.byref(P), ids.ctypes.data, plogs.ctypes.data, ids.size, result_len, out)
        return dict(sum_logprobs=out[0], avg_logprobs=out[1], entropy=out[2], score=out[3])

    def kept_logits(self) -> np.ndarray:
        n = self.L.wo_n_kept_logits(self.h)
        nv = self.m.hparams["n_vocab"]
        if n == 0:
            return np.zeros((0, nv), np.float32)
        p = self.L.wo_kept_logits(self.h, 0)
        return np.ctypeslib.as_array(p, shape=(n, nv)).copy()


# ---- audio pre-processing oracle (oracle/audio_oracle.c: /root/reference/src/audio/mod.rs restated) ----
def _audio_lib():
    L = lib()
    if not getattr(L, "_audio_ready", False):
        L.ao_analyze_noise.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_float)]
        L.ao_denoise_audio.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ao_estimate_noise_floor.restype = C.c_float
        L.ao_estimate_noise_floor.argtypes = [C.c_void_p, C.c_size_t]
        L.ao_normalize_audio.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.ao_process_frame.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L._audio_ready = True
    return L


def analyze_noise(samples, frame_size: int = 2048):
    """-> (noise type 0 stationary / 1 non-stationary / 2 mixed, normalised spectral variance)"""
    x = np.ascontiguousarray(samples, dtype=np.float32)
    nv = C.c_float()
    t = _audio_lib().ao_analyze_noise(x.ctypes.data, x.size, frame_size, C.byref(nv))
    return t, float(nv.value)


def denoise_audio(samples, frame_size: int = 2048, overlap: float = 0.75, strength: float = 0.2):
    """reference denoise_audio (audio/mod.rs:507-528) -> (f32 samples, noise type)"""
    x = np.ascontiguousarray(samples, dtype=np.float32)
    out = np.zeros_like(x)
    t = _audio_lib().ao_denoise_audio(x.ctypes.data, x.size, frame_size, overlap, strength, out.ctypes.data)
    if t < 0:
        raise ValueError("denoise_audio: fewer samples than one frame (the reference panics here)")
    return out, t


class StreamAudioProcessor:
    """reference StreamAudioProcessor (audio/mod.rs:80-155): per-chunk peak normalisation, 2048-sample frames,
    VAD gain, denoise, noise gate; returns the list of processed frames of each call."""

    def __init__(self, frame_size=2048, overlap=0.75, strength=0.2, noise_gate=0.003, enable_noise_reduction=True):
        self.cfg = (frame_size, overlap, strength, noise_gate, enable_noise_reduction)
        self.buffer = np.zeros(0, np.float32)
        self.state = np.zeros(2, np.float32)        # noise_floor, prev_energy

    def _frame(self, frame):
        fs, ov, st, ng, nr = self.cfg
        out = np.zeros(fs, np.float32)
        f = np.ascontiguousarray(frame, np.float32)
        _audio_lib().ao_process_frame(f.ctypes.data, fs, ov, st, ng, int(nr), self.state.ctypes.data, out.ctypes.data)
        return out

    def process_chunk(self, chunk):
        x = np.ascontiguousarray(chunk, np.float32)
        norm = np.zeros_like(x)
        _audio_lib().ao_normalize_audio(x.ctypes.data, x.size, norm.ctypes.data)
        self.buffer = np.concatenate([self.buffer, norm])
        fs = self.cfg[0]
        frames = []
        while self.buffer.size >= fs:
            frames.append(self._frame(self.buffer[:fs]))
            self.buffer = self.buffer[fs:]
        return frames

    def finish(self):
        fs = self.cfg[0]
        if self.buffer.size == 0:
            return []
        frame = np.zeros(fs, np.float32)
        frame[:self.buffer.size] = self.buffer
        self.buffer = np.zeros(0, np.float32)
        return [self._frame(frame)]
