"""ctypes binding of libspeaksense_whisper.so (include/speaksense_whisper.h).

The product path is the CUDA library only: if the extension is missing or no sm_100 device is
present, calls raise - there is no CPU / eager fallback (and nothing here imports oracle/)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SS_LIB_PATH: another build of the same library (tools/asan_host.sh: the host side compiled with AddressSanitizer / UBSan)
LIB_PATH = os.environ.get("SS_LIB_PATH") or os.path.join(_HERE, "lib", "libspeaksense_whisper.so")


class SsParams(C.Structure):
    _fields_ = [("language", C.c_char_p), ("speaker_diarization", C.c_int), ("stream_mode", C.c_int),
                ("min_segment_length", C.c_int), ("beam_size", C.c_int), ("debug_keep_logits", C.c_int)]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("speaksense_whisper error %d: %s" % (code, msg))
        self.code = code


# (name, restype, argtypes) == every symbol include/speaksense_whisper.h declares
_P = C.c_void_p
SYMBOLS = [
    ("ss_params_default", None, [C.POINTER(SsParams)]),
    ("ss_abi_version", C.c_int, []),
    ("ss_last_error", C.c_char_p, []),
    ("ss_build_info", C.c_char_p, []),
    ("ss_engine_open", C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    ("ss_nccl_unique_id", C.c_int, [C.c_char_p]),
    ("ss_engine_open_dist", C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_P)]),
    ("ss_engine_open_multi", C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    ("ss_engine_n_devices", C.c_int, [_P]),
    ("ss_engine_device", C.c_int, [_P, C.c_int]),
    ("ss_engine_n_states", C.c_int, [_P, C.c_int]),
    ("ss_engine_arena_fnv1a", C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64)]),
    ("ss_engine_close", None, [_P]),
    ("ss_engine_info", C.c_int, [_P] + [C.POINTER(C.c_int)] * 5 + [C.POINTER(C.c_int64)]),
    ("ss_state_new", C.c_int, [_P, C.POINTER(_P)]),
    ("ss_state_new_on", C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    ("ss_state_device", C.c_int, [_P]),
    ("ss_state_status", C.c_int, [_P]),
    ("ss_state_error", C.c_char_p, [_P]),
    ("ss_state_free", None, [_P]),
    ("ss_transcribe", C.c_int, [_P, _P, _P, C.c_size_t, C.POINTER(SsParams)]),
    ("ss_upload_pcm", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("ss_transcribe_resident", C.c_int, [_P, _P, C.POINTER(SsParams)]),
    ("ss_denoise_config_default", None, [_P]),
    ("ss_denoise_audio", C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    ("ss_denoise_frames", C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    ("ss_bench_decode_steps", C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    ("ss_transcribe_batch", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_size_t), C.c_int, C.POINTER(SsParams)]),
    ("ss_n_segments_raw", C.c_int, [_P]),
    ("ss_segment_text_raw", C.c_char_p, [_P, C.c_int]),
    ("ss_segment_t0_raw", C.c_int64, [_P, C.c_int]),
    ("ss_segment_t1_raw", C.c_int64, [_P, C.c_int]),
    ("ss_segment_speaker_turn_next_raw", C.c_int, [_P, C.c_int]),
    ("ss_n_segments", C.c_int, [_P]),
    ("ss_segment_text", C.c_char_p, [_P, C.c_int]),
    ("ss_segment_start", C.c_double, [_P, C.c_int]),
    ("ss_segment_end", C.c_double, [_P, C.c_int]),
    ("ss_segment_speaker_id", C.c_int, [_P, C.c_int]),
    ("ss_full_text", C.c_char_p, [_P]),
    ("ss_n_result_tokens", C.c_int, [_P]),
    ("ss_result_token", C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    ("ss_n_fallbacks", C.c_int, [_P]),
    ("ss_n_decoded", C.c_int, [_P]),
    ("ss_n_windows", C.c_int, [_P]),
    ("ss_n_kernel_launches", C.c_int, [_P]),
    ("ss_stage_ms", C.c_int, [_P] + [C.POINTER(C.c_float)] * 3),
    ("ss_debug_logits", C.POINTER(C.c_float), [_P, C.c_int, C.POINTER(C.c_int)]),
    ("ss_is_promotional_text", C.c_int, [C.c_char_p]),
    ("ss_add_punctuation", C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    ("ss_is_valid_utf8", C.c_int, [C.c_char_p, C.c_size_t]),
    ("ss_model_probe", C.c_int, [C.c_char_p, C.POINTER(C.c_int * 11), C.POINTER(C.c_int64), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("ss_log_mel", C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("ss_encode", C.c_int, [_P, _P, C.c_int, _P, C.c_size_t]),
    ("ss_decode", C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    ("ss_debug_gemm", C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("ss_debug_process_logits", C.c_int, [C.c_char_p, _P, C.c_int, C.c_int, C.c_int, _P, C.c_float, _P]),
    ("ss_debug_sequence_score", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P]),
    ("ss_debug_beam_assign", C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
]

_lib = None


def lib():
    """Load the CUDA extension; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(-3, "CUDA extension %s is missing: run `python -m speaksense_b200.build` "
                                  "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise NativeError(rc, lib().ss_last_error().decode("utf-8", "replace"))
