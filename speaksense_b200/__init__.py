"""speaksense_b200 - B200-native Whisper transcribe path behind SpeakSense's AsrEngine boundary.

Package contents: csrc/ (sm_100a CUDA kernels + C ABI), build.py (nvcc driver), _native.py (ctypes
binding), asr.py (host mirror of /root/reference/src/asr), synth.py (synthetic ggml models / audio
for tests and benchmarks), audio.py (host mirror of /root/reference/src/audio denoise, GPU-backed), stream.py / grpc_server.py / rest.py (the callers either
side of the path), batching.py (front end that merges concurrent calls into ss_transcribe_batch)."""
from .asr import AsrEngine, AsrParams, TranscribeResult, TranscribeSegment, WhisperAsr, WhisperState  # noqa: F401
from ._native import NativeError  # noqa: F401
from .audio import DenoiseConfig, StreamAudioProcessor, denoise_audio  # noqa: F401
from .batching import BatchingEngine  # noqa: F401

__all__ = ["AsrEngine", "AsrParams", "TranscribeResult", "TranscribeSegment", "WhisperAsr", "WhisperState", "NativeError",
           "DenoiseConfig", "StreamAudioProcessor", "denoise_audio", "BatchingEngine"]
