"""Host mirror of the reference's audio pre-processing in front of the transcribe path
(/root/reference/src/audio/mod.rs).  Same names, argument meaning and quirks as the
reference; the arithmetic runs on the GPU through the C ABI (ss_denoise_audio, csrc/denoise.cu) - there is
no CPU fallback.

    DenoiseConfig          == audio/mod.rs:41-62
    denoise_audio          == audio/mod.rs:507-528  (gRPC handler: grpc/handlers/asr.rs:196, per 5 s chunk)
    StreamAudioProcessor   == audio/mod.rs:80-155   (REST path, 2048-sample frames)

Quirks kept: the output of denoise_audio is scaled by frame_size x 10 (unnormalised
inverse FFT, x10 in overlap_add); StreamAudioProcessor's first-frame noise floor is NaN so its VAD gain is
0.1 for ever; an all-zero chunk normalises to NaN.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from . import _native

NOISE_TYPES = ("Stationary", "NonStationary", "Mixed")          # audio/mod.rs:531-536


class SsDenoiseConfig(C.Structure):
    _fields_ = [("frame_size", C.c_int), ("overlap", C.c_float), ("strength", C.c_float), ("noise_gate", C.c_float),
                ("enable_noise_reduction", C.c_int), ("threshold", C.c_float)]


@dataclass
class DenoiseConfig:
    frame_size: int = 2048
    overlap: float = 0.75
    strength: float = 0.2
    noise_gate: float = 0.003
    enable_noise_reduction: bool = True
    threshold: float = 0.002

    def _native(self) -> SsDenoiseConfig:
        return SsDenoiseConfig(self.frame_size, self.overlap, self.strength, self.noise_gate,
                               int(self.enable_noise_reduction), self.threshold)


def denoise_audio(engine, state, samples, config: Optional[DenoiseConfig] = None, fetch: bool = True):
    """denoise_audio(&samples, &config) on the engine's GPU.  The denoised chunk stays resident as `state`'s
    PCM (follow with engine.transcribe_resident(state, params) to skip the second upload).
    Returns (samples f32 or None when fetch=False, noise type name, normalised spectral variance)."""
    cfg = (config or DenoiseConfig())._native()
    x = np.ascontiguousarray(samples, dtype=np.float32)
    out = np.empty_like(x) if fetch else None
    t, nv = C.c_int(), C.c_float()
    with state._lock:
        _native.check(_native.lib().ss_denoise_audio(engine._h, state._h, x.ctypes.data, x.size, C.byref(cfg),
                                                     out.ctypes.data if fetch else None, C.byref(t), C.byref(nv)))
    return out, NOISE_TYPES[t.value], float(nv.value)


def denoise_frames(engine, state, frames: np.ndarray, config: Optional[DenoiseConfig] = None) -> np.ndarray:
    """steps 4-5 of StreamAudioProcessor::process_frame (audio/mod.rs:131-139) for all rows of `frames`
    ([n_frames, frame_size] f32, already scaled by the VAD gain) in one launch: denoise_audio on each single frame + noise gate"""
    cfg = config or DenoiseConfig()
    x = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, cfg.frame_size)
    out = np.empty_like(x)
    c = cfg._native()
    with state._lock:
        _native.check(_native.lib().ss_denoise_frames(engine._h, state._h, x.ctypes.data, x.shape[0], C.byref(c), out.ctypes.data))
    return out


class StreamAudioProcessor:
    """StreamAudioProcessor::{new, process_chunk, finish} (audio/mod.rs:80-155).  `callback` receives every
    processed 2048-sample frame, in order.  The scalar recurrences (VAD gain, noise floor) run on the host
    in f32 exactly as written in the reference; the denoise of every frame runs on the GPU."""

    def __init__(self, engine, state, config: Optional[DenoiseConfig], callback: Callable[[np.ndarray], None]):
        self.engine, self.state = engine, state
        self.config = config or DenoiseConfig()
        self.frame_size = 2048                      # mod.rs:84 (fixed, independent of config.frame_size)
        self.buffer = np.zeros(0, np.float32)
        self.callback = callback
        self.prev_energy = np.float32(0.0)
        self.noise_floor = np.float32(0.0)

    @staticmethod
    def _estimate_noise_floor(frame: np.ndarray) -> np.float32:      # mod.rs:744-762
        e = sorted(np.float32(np.sum(c * c, dtype=np.float32) / np.float32(c.size)) for c in
                   (frame[i:i + 1024] for i in range(0, frame.size, 1024)))
        cnt = int(np.float32(len(e)) * np.float32(0.1))
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.float32(np.float32(sum(e[:cnt], np.float32(0))) / np.float32(cnt))

    def _gain(self, frame: np.ndarray) -> np.float32:      # mod.rs:111-128 (steps 1-2 + the state update)
        pre = frame.copy()
        pre[1:] = frame[1:] - np.float32(0.97) * frame[:-1]
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            energy = np.float32(np.cumsum(pre * pre, dtype=np.float32)[-1] / np.float32(frame.size))      # sequential f32 sum, as Iterator::sum
            threshold = np.float32(self.noise_floor * np.float32(1.2) + self.prev_energy * np.float32(0.1))
            if energy > threshold:
                gain = np.float32(1.0)
            else:
                r = np.float32(energy / threshold)
                gain = np.float32(0.1) if np.isnan(r) else max(r, np.float32(0.1))
            self.prev_energy = energy
            mn = energy if np.isnan(self.noise_floor) else min(energy, self.noise_floor)
            self.noise_floor = np.float32(self.noise_floor * np.float32(0.95) + mn * np.float32(0.05))
        return gain

    def _process_frames(self, frames: List[np.ndarray]) -> None:
        """process_frame for a run of frames: the scalar recurrences in order on the host, then ONE launch for steps 3-5"""
        if not frames:
            return
        scaled = np.empty((len(frames), self.frame_size), np.float32)
        for k, frame in enumerate(frames):
            if self.noise_floor == 0.0:                       # mod.rs:102-104
                self.noise_floor = self._estimate_noise_floor(frame)
            scaled[k] = frame * self._gain(frame)             # step 3
        cfg = DenoiseConfig(self.frame_size, self.config.overlap, self.config.strength, self.config.noise_gate,
                            self.config.enable_noise_reduction, self.config.threshold)
        if self.frame_size != self.config.frame_size:
            # the reference frames by its fixed 2048 (mod.rs:84) but denoises with config.frame_size: only equal sizes take the
            # single-window route; anything else goes through the general entry point frame by frame
            outs = []
            for row in scaled:
                o, _, _ = denoise_audio(self.engine, self.state, row, self.config) if self.config.enable_noise_reduction else (row.copy(), None, None)
                o[np.abs(o) < np.float32(self.config.noise_gate)] = 0.0
                outs.append(o)
            out = np.stack(outs)
        else:
            out = denoise_frames(self.engine, self.state, scaled, cfg)
        for row in out:
            self.callback(row.copy())

    def process_chunk(self, chunk) -> None:      # mod.rs:92-109
        x = np.ascontiguousarray(chunk, dtype=np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            x = x / (np.max(np.abs(x)) if x.size else np.float32(1.0))      # normalize_audio :408-411
        self.buffer = np.concatenate([self.buffer, x.astype(np.float32)])
        frames = []
        while self.buffer.size >= self.frame_size:
            frames.append(self.buffer[:self.frame_size].copy())
            self.buffer = self.buffer[self.frame_size:]
        self._process_frames(frames)

    def finish(self) -> None:      # mod.rs:143-155
        if self.buffer.size:
            frame = np.zeros(self.frame_size, np.float32)
            frame[:self.buffer.size] = self.buffer
            self.buffer = np.zeros(0, np.float32)
            self._process_frames([frame])


def collect_frames() -> tuple:
    """convenience: (list, callback) pair for StreamAudioProcessor"""
    frames: List[np.ndarray] = []
    return frames, frames.append
