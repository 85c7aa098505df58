"""Host mirror of the reference's gRPC streaming caller of the transcribe path
(/root/reference/src/grpc/handlers/asr.rs; SURVEY.md §8 row f2, minus the tonic transport): the per-stream
state machine that turns base64 PCM16 messages into 5 s chunks, denoises and transcribes every chunk on one
`ss_state`, de-duplicates the text and rebases segment times.  Same names and quirks as the reference
(SURVEY Appendix B.2/B.3):

  * CHUNK_SIZE = SAMPLE_RATE * 10 is applied to a BYTE buffer => 5 s of audio per chunk (asr.rs:14-18,187);
    the buffer advances by CHUNK_SIZE - OVERLAP_SIZE bytes = 4.5 s, yet block times assume 5.0 s (:40,231);
  * samples are scaled by 1/32767 (:192);
  * TranscribeSegment.start/end are 10 ms ticks but are treated as seconds: ((block*5 + t) * 1000) as i64 (:40-43);
  * language "zh", stream_mode, min_segment_length 5 are hard-coded (:154-157);
  * the tail (end == 1) is NOT denoised and runs through AsrEngine::transcribe, i.e. on a fresh state (:233-262).

On the GPU a chunk is uploaded once: ss_denoise_audio leaves the denoised chunk resident and
ss_transcribe_resident consumes it.
"""
from __future__ import annotations

import base64
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .asr import AsrParams, TranscribeSegment, WhisperAsr
from .audio import DenoiseConfig, denoise_audio

SAMPLE_RATE = 16000                 # asr.rs:14
CHUNK_SIZE = SAMPLE_RATE * 10       # asr.rs:16  (bytes!)
OVERLAP_SIZE = SAMPLE_RATE          # asr.rs:18  (bytes!)
_SENTENCE_END = "。！？.!?"


@dataclass
class Segment:                      # proto/asr.proto:39-43
    start: int
    end: int
    text: bytes


@dataclass
class TranscribeResponse:           # proto/asr.proto:32-37
    end: int
    text: bytes
    device_id: str
    segments: List[Segment] = field(default_factory=list)


class StreamContext:                # asr.rs:24-60
    def __init__(self):
        self.block_index = 0
        self.last_text = ""
        self.last_end_time = 0.0

    def calculate_segment_time(self, segment_start: float, segment_end: float):
        block_base_time = float(self.block_index) * 5.0
        abs_start = int((block_base_time + segment_start) * 1000.0)
        abs_end = int((block_base_time + segment_end) * 1000.0)
        last_end_ms = int(self.last_end_time * 1000.0)
        if abs_start < last_end_ms:
            diff = last_end_ms - abs_start
            abs_start = last_end_ms
            abs_end += diff
        self.last_end_time = abs_end / 1000.0
        return abs_start, abs_end

    def next_block(self):
        self.block_index += 1


def _split_sentences(text: str) -> List[str]:
    out, cur = [], []
    for ch in text:
        if ch in _SENTENCE_END:
            out.append("".join(cur)); cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return [s for s in out if s.strip()]


def process_text(new_text: str, last_text: str, segments: List[TranscribeSegment]) -> Optional[str]:
    """AsrService::process_text (asr.rs:69-136); lengths are UTF-8 byte lengths as in Rust"""
    if not last_text:
        return new_text
    if segments:
        if segments[-1].text not in last_text:
            return segments[-1].text
    nl, ll = len(new_text.encode("utf-8")), len(last_text.encode("utf-8"))
    if nl > ll and new_text.startswith(last_text):
        added = new_text[len(last_text):]
        if added.strip():
            return added.strip()
    if nl > ll * 2 or ll > nl * 2:
        return new_text
    if new_text != last_text:
        new_s, last_s = _split_sentences(new_text), _split_sentences(last_text)
        if len(new_s) > len(last_s):
            content = "".join(new_s[len(last_s):]).strip()
            if content:
                if new_text and new_text[-1] in _SENTENCE_END:
                    content += new_text[-1]
                return content
        elif new_s and last_s:
            if new_s[-1].strip() != last_s[-1].strip():
                result = new_s[-1].strip()
                if new_text and new_text[-1] in _SENTENCE_END:
                    result += new_text[-1]
                return result
    return None


def pcm16_to_f32(data: bytes, exact: bool) -> np.ndarray:
    """asr.rs:188-194 (chunks_exact(2)) / :235-245 (chunks(2): a trailing odd byte becomes one 0.0 sample)"""
    n = len(data) // 2
    x = np.frombuffer(data[:2 * n], dtype="<i2").astype(np.float32) / np.float32(32767.0)
    if not exact and len(data) % 2:
        x = np.concatenate([x, np.zeros(1, np.float32)])
    return x


class AsrStreamSession:
    """One gRPC Transcribe stream (asr.rs:146-281): `feed` takes one TranscribeRequest (base64 audio + end flag)
    and returns the TranscribeResponses the handler would send for it."""

    def __init__(self, engine: WhisperAsr, denoise: Optional[DenoiseConfig] = None):
        self.engine = engine
        self.state = engine.create_state()                              # asr.rs:164
        self.params = AsrParams(language="zh", stream_mode=True, min_segment_length=5)      # asr.rs:154-157
        self.denoise = denoise or DenoiseConfig()
        self.audio_buffer = bytearray()
        self.device_id = ""
        self.ctx = StreamContext()
        self.n_chunks = 0
        self.closed = False

    def close(self):
        if not self.closed:
            self.state.close(); self.closed = True

    def feed(self, audio_b64: bytes, end: int = 0, device_id: str = "") -> List[TranscribeResponse]:
        out: List[TranscribeResponse] = []
        if not self.device_id:
            self.device_id = device_id
        try:
            decoded = base64.b64decode(audio_b64, validate=True)
        except Exception:      # noqa: BLE001  (asr.rs:176-183: log and skip the message)
            return out
        self.audio_buffer.extend(decoded)
        if len(self.audio_buffer) >= CHUNK_SIZE:                       # asr.rs:187 (one chunk per message at most)
            float_data = pcm16_to_f32(bytes(self.audio_buffer[:CHUNK_SIZE]), exact=True)
            denoise_audio(self.engine, self.state, float_data, self.denoise, fetch=False)      # asr.rs:196
            try:
                result = self.engine.transcribe_resident(self.state, self.params)               # asr.rs:198
            except Exception:      # noqa: BLE001  (asr.rs:228: log, keep the stream alive)
                result = None
            if result is not None:
                for seg in result.segments:
                    new_text = process_text(seg.text, self.ctx.last_text, [seg])
                    if new_text is not None:
                        self.ctx.last_text = seg.text
                        start, end_t = self.ctx.calculate_segment_time(seg.start, seg.end)
                        out.append(TranscribeResponse(end=0, text=new_text.encode("utf-8"), device_id=self.device_id,
                                                      segments=[Segment(start, end_t, seg.text.encode("utf-8"))]))
                self.ctx.next_block()
            self.n_chunks += 1
            del self.audio_buffer[:CHUNK_SIZE - OVERLAP_SIZE]          # asr.rs:231
        if end == 1 and len(self.audio_buffer) > 0:                    # asr.rs:234-263
            float_data = pcm16_to_f32(bytes(self.audio_buffer), exact=False)
            try:
                result = self.engine.transcribe(float_data, self.params)      # fresh state, no denoise
            except Exception:      # noqa: BLE001
                result = None
            if result is not None:
                final_text = process_text(result.full_text, self.ctx.last_text, result.segments)
                if final_text is not None:
                    segs = []
                    for seg in result.segments:
                        start, end_t = self.ctx.calculate_segment_time(seg.start, seg.end)
                        segs.append(Segment(start, end_t, seg.text.encode("utf-8")))
                    out.append(TranscribeResponse(end=1, text=final_text.encode("utf-8"), device_id=self.device_id, segments=segs))
        return out


def encode_messages(pcm_f32: np.ndarray, message_bytes: int = 32 * 1024):
    """client side as examples/asr_client.rs:142,169-179: f32 -> PCM16LE -> base64 messages of 32 KiB; last has end = 1"""
    raw = (np.clip(pcm_f32, -1.0, 1.0) * 32767.0).astype("<i2").tobytes()
    msgs = [raw[i:i + message_bytes] for i in range(0, len(raw), message_bytes)] or [b""]
    return [(base64.b64encode(m), 1 if i == len(msgs) - 1 else 0) for i, m in enumerate(msgs)]
