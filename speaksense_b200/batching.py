"""Micro-batching front end: concurrent sessions, one decoder step.

The reference serves every gRPC stream / REST task from its own tokio worker and calls `transcribe_with_state` there
(/root/reference/src/grpc/handlers/asr.rs:198, src/schedule/processors/transcribe.rs:112): N concurrent streams are N
independent `state.full` calls.  On one GPU those calls take turns, each streaming the decoder weights once per token.
`BatchingEngine` sits where the handlers hold their `Arc<dyn AsrEngine>`: it has the engine's interface, but calls that
arrive while the device is busy (or within `linger_s` of each other) are merged into one `ss_transcribe_batch`, which decodes
them with one batched step per token (csrc/decoder_batch.cu).  Every caller still gets exactly the result its own call would
have produced; a failure of one clip (e.g. invalid UTF-8 in a segment, whisper.rs:85) fails that call only.

INTEGRATION.md §2d describes the same collector for the Rust host (mpsc queue + oneshot replies).
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .asr import AsrEngine, AsrParams, TranscribeResult


@dataclass
class _Request:
    state: object
    pcm: Optional[np.ndarray]          # None: the state's resident PCM
    params: AsrParams
    future: Future


def _params_key(p: AsrParams):
    return (p.language, bool(p.speaker_diarization), bool(p.stream_mode), int(getattr(p, "beam_size", 0) or 0),
            bool(getattr(p, "debug_keep_logits", False)))


class BatchingEngine(AsrEngine):
    """Same interface as WhisperAsr for the calls the stream / REST mirrors make; merges concurrent calls."""

    def __init__(self, engine, max_batch: int = 32, linger_s: float = 0.0005):
        self.engine = engine
        self.max_batch = max_batch
        self.linger_s = linger_s
        self.info = getattr(engine, "info", None)
        self.n_batches = 0
        self.n_requests = 0
        self.max_seen = 0
        self._q: "queue.Queue[Optional[_Request]]" = queue.Queue()
        self._closed = False
        self._broken: Optional[BaseException] = None      # the worker died: every later call fails with this
        self._mu = threading.Lock()                        # orders _submit's enqueue against close()'s sentinel
        self._thread = threading.Thread(target=self._run, name="ss-batching", daemon=True)
        self._thread.start()

    # the denoise mirror (audio.denoise_audio) talks to the native engine handle directly
    @property
    def _h(self):
        return self.engine._h

    def create_state(self):
        return self.engine.create_state()

    def upload_pcm(self, state, audio):
        return self.engine.upload_pcm(state, audio)

    def transcribe_with_state(self, state, audio, params: AsrParams) -> TranscribeResult:
        return self._submit(state, np.ascontiguousarray(audio, dtype=np.float32), params)

    def transcribe_resident(self, state, params: AsrParams) -> TranscribeResult:
        return self._submit(state, None, params)

    def close(self, close_engine: bool = False):
        with self._mu:
            first = not self._closed
            self._closed = True
            if first:
                self._q.put(None)      # nothing can be enqueued behind the sentinel: _submit checks _closed under the same lock
        if first:
            self._thread.join()
            self._fail_queued(RuntimeError("BatchingEngine is closed"))
        if close_engine:
            self.engine.close()

    # ------------------------------------------------------------------------------------------
    def _submit(self, state, pcm, params) -> TranscribeResult:
        fut: Future = Future()
        with self._mu:
            if self._broken is not None:
                raise RuntimeError("BatchingEngine worker died: %r" % (self._broken,))
            if self._closed:
                raise RuntimeError("BatchingEngine is closed")
            self._q.put(_Request(state, pcm, params, fut))
        return fut.result()

    def _fail_queued(self, exc: BaseException):
        while True:
            try:
                r = self._q.get_nowait()
            except queue.Empty:
                return
            if r is not None and not r.future.done():
                r.future.set_exception(exc)

    def _single(self, r: _Request) -> TranscribeResult:
        if r.pcm is None:
            return self.engine.transcribe_resident(r.state, r.params)
        return self.engine.transcribe_with_state(r.state, r.pcm, r.params)

    def _execute(self, group: List[_Request]):
        """one group of compatible requests on distinct states"""
        self.n_batches += 1
        self.n_requests += len(group)
        self.max_seen = max(self.max_seen, len(group))
        if len(group) > 1:
            try:
                # a clip whose read-back fails (whisper.rs:85) fails alone inside the native batch call: its slot comes back as the
                # error, the other clips keep their results - nothing is decoded twice and no state is advanced twice
                results = self.engine.transcribe_batch([r.state for r in group], [r.pcm for r in group], group[0].params,
                                                       return_exceptions=True)
                for r, res in zip(group, results):
                    if isinstance(res, BaseException):
                        r.future.set_exception(res)
                    else:
                        r.future.set_result(res)
                return
            except TypeError:      # an engine without per-clip error reporting (test stubs): clip by clip below
                pass
            except Exception as e:      # noqa: BLE001  the call failed as a whole (bad language, out of memory, ...): so does every clip
                for r in group:
                    r.future.set_exception(e)
                return
        for r in group:
            try:
                r.future.set_result(self._single(r))
            except Exception as e:      # noqa: BLE001
                r.future.set_exception(e)

    def _run(self):
        pending: List[_Request] = []
        try:
            self._loop(pending)
        except BaseException as e:      # noqa: BLE001  never leave a caller blocked in fut.result()
            with self._mu:
                self._broken = e
            for r in pending:
                if not r.future.done():
                    r.future.set_exception(e)
            self._fail_queued(e)

    def _loop(self, pending: List[_Request]):
        stop = False
        while not (stop and not pending):
            if not pending:
                r = self._q.get()
                if r is None:
                    stop = True
                    continue
                pending.append(r)
            deadline = time.monotonic() + self.linger_s
            while len(pending) < self.max_batch and not stop:
                try:
                    r = self._q.get(timeout=max(0.0, deadline - time.monotonic()))
                except queue.Empty:
                    break
                if r is None:
                    stop = True
                    break
                pending.append(r)
            # the first request decides the group: same parameters, every state at most once (a second call on a state
            # waits for the next round, as the state's mutex would make it, whisper.rs:51-54)
            key = _params_key(pending[0].params)
            group, rest, seen = [], [], set()
            for r in pending:
                if len(group) < self.max_batch and _params_key(r.params) == key and id(r.state) not in seen:
                    group.append(r)
                    seen.add(id(r.state))
                else:
                    rest.append(r)
            pending[:] = rest
            try:
                self._execute(group)
            except BaseException as e:      # noqa: BLE001
                for r in group:
                    if not r.future.done():
                        r.future.set_exception(e)
                raise
