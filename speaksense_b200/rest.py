"""Host mirror of the reference's REST task processor around the transcribe path
(/root/reference/src/schedule/processors/transcribe.rs:62-167 + src/audio/mod.rs:157-233):
WAV -> 4096-sample chunks (/32768) -> mono -> StreamAudioProcessor (2048-sample frames) -> 30 s buffering ->
transcribe_with_state(stream_mode = true) per buffer on ONE state -> concatenated text + all segments.

Kept as in the reference: the buffer is cut when it reaches >= 16000 * 30 samples, i.e. after 235 frames = 481280 samples
(transcribe.rs:104-109); every call keeps only its last segment (stream mode, whisper.rs:102-111); segment times are the
raw 10 ms ticks (transcribe.rs:160-165).  Not restated: ffmpeg transcoding (ensure_wav_format) and the rubato resampler
(input must be 16 kHz PCM16 WAV; SURVEY Appendix B.6 documents that the reference's resampling path errors for most
inputs anyway); the reference forwards frames through one tokio task each, so their order is not guaranteed
(Appendix B.7) - here frames keep file order."""
from __future__ import annotations

import wave
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from ._native import NativeError
from .asr import AsrParams, WhisperAsr
from .audio import DenoiseConfig, StreamAudioProcessor

CHUNK_SIZE = 4096                   # audio/mod.rs:186
BUFFER_SIZE = 16000 * 30            # processors/transcribe.rs:104
VALID_LANGUAGES = ("zh", "en", "ja")      # processors/transcribe.rs:200-204


@dataclass
class RestSegment:                  # TranscribeSegment of the task result (transcribe.rs:160-165)
    text: str
    speaker_id: Optional[int]
    start_time: float
    end_time: float


@dataclass
class RestTranscribeResult:
    text: str
    segments: List[RestSegment] = field(default_factory=list)
    n_calls: int = 0
    n_failed: int = 0


def read_wav_pcm16(path: str):
    """-> (interleaved f32 samples / 32768, channels, sample rate); 16-bit integer PCM only (audio/mod.rs:355-389)"""
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("Unsupported bits per sample: expected 16 bits")
        ch, sr = w.getnchannels(), w.getframerate()
        raw = w.readframes(w.getnframes())
    return np.frombuffer(raw, "<i2").astype(np.float32) / np.float32(32768.0), ch, sr


def convert_to_mono(samples: np.ndarray, num_channels: int) -> np.ndarray:      # audio/mod.rs:391-398
    if num_channels == 1:
        return samples
    n = (samples.size + num_channels - 1) // num_channels
    pad = np.zeros(n * num_channels, np.float32)
    pad[:samples.size] = samples
    return (pad.reshape(n, num_channels).sum(axis=1, dtype=np.float32) / np.float32(num_channels)).astype(np.float32)


class TranscribeProcessor:
    """TranscribeProcessor::process_audio (processors/transcribe.rs:62-167)"""

    def __init__(self, engine: WhisperAsr):
        self.asr = engine

    @staticmethod
    def validate_params(language: Optional[str]) -> None:      # transcribe.rs:196-208
        if language is not None and language not in VALID_LANGUAGES:
            raise ValueError("Unsupported language: %s" % language)

    def process_audio(self, wav_path: str, language: Optional[str] = None, speaker_diarization: bool = False,
                      config: Optional[DenoiseConfig] = None) -> RestTranscribeResult:
        self.validate_params(language)
        samples, ch, sr = read_wav_pcm16(wav_path)
        if sr != 16000:
            raise ValueError("resampling is not restated: the input must be 16 kHz (got %d)" % sr)
        params = AsrParams(language=language, speaker_diarization=speaker_diarization, stream_mode=True)      # :66-69
        state = self.asr.create_state()                                                                      # :100
        frames: List[np.ndarray] = []
        proc = StreamAudioProcessor(self.asr, state, config or DenoiseConfig(), frames.append)
        for o in range(0, samples.size, CHUNK_SIZE):                       # audio/mod.rs:190-221
            proc.process_chunk(convert_to_mono(samples[o:o + CHUNK_SIZE], ch))
        proc.finish()
        res = RestTranscribeResult(text="")
        buf: List[np.ndarray] = []
        n_buf = 0

        def flush():
            nonlocal buf, n_buf
            audio, buf, n_buf = np.concatenate(buf), [], 0
            try:
                r = self.asr.transcribe_with_state(state, audio, params)      # :110 / :128
            except NativeError:                                               # :117-119: log, drop this buffer, go on
                res.n_failed += 1
                return
            res.text += r.full_text
            res.segments += [RestSegment(s.text, s.speaker_id, s.start, s.end) for s in r.segments]
            res.n_calls += 1

        for f in frames:                                                    # transcribe.rs:106-123
            buf.append(f); n_buf += f.size
            if n_buf >= BUFFER_SIZE:
                flush()
        if n_buf:                                                           # :125-140
            flush()
        state.close()
        return res


# ---- callback JSON of a finished task (src/schedule/callback/mod.rs:28-33,80-97; src/schedule/types.rs:87-95,118-138) ----
def callback_payload(task_id: str, result: RestTranscribeResult) -> dict:
    """what HttpCallback::on_complete POSTs: CallbackPayload{task_id, status: Completed, data: TaskResult::Transcribe(..)}
    (TaskResult is #[serde(tag = "type", content = "result")]; speaker_id is Option<usize>)"""
    return {"task_id": task_id, "status": "Completed",
            "data": {"type": "Transcribe",
                     "result": {"text": result.text,
                                "segments": [{"text": s.text, "speaker_id": s.speaker_id, "start_time": s.start_time, "end_time": s.end_time}
                                             for s in result.segments]}}}


def callback_error_payload(task_id: str, error: str) -> dict:
    """HttpCallback::on_error: status = TaskStatus::Failed(msg) (an externally tagged newtype variant), data = the message"""
    return {"task_id": task_id, "status": {"Failed": error}, "data": error}
