"""gRPC transport of the streaming caller: service `asr.Asr`, method
`Transcribe(stream TranscribeRequest) returns (stream TranscribeResponse)` exactly as /root/reference/proto/asr.proto:1-43
defines it, served by AsrStreamSession (the handler logic of src/grpc/handlers/asr.rs:146-281).

There is no protoc / grpc_tools offline, so the message classes are built at import time from a FileDescriptorProto that
restates proto/asr.proto field by field (same package, names, numbers and types: wire-compatible with the reference's
tonic server and its examples/asr_client.rs), and the method is registered through grpcio's generic handler API.

    server = serve(engine, "127.0.0.1:7300")          # one AsrStreamSession (= one ss_state) per stream
    ...
    server.stop(0)
"""
from __future__ import annotations

from concurrent import futures
from typing import Callable, Iterable, Iterator, Optional

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from . import stream as _stream

SERVICE = "asr.Asr"
METHOD = "/asr.Asr/Transcribe"


def _build_messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "speaksense_b200/asr.proto"
    fd.package = "asr"
    fd.syntax = "proto3"
    T = descriptor_pb2.FieldDescriptorProto
    en = fd.enum_type.add(); en.name = "AudioFormat"                                     # proto/asr.proto:10-20
    for i, n in enumerate(["AAC", "MP3", "WAV", "OGG", "FLAC", "AMR", "OPUS", "PCM", "UNKNOWN"]):
        v = en.value.add(); v.name = n; v.number = i

    def field(msg, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None):
        f = msg.field.add(); f.name = name; f.number = number; f.type = ftype; f.label = label
        if type_name:
            f.type_name = type_name

    req = fd.message_type.add(); req.name = "TranscribeRequest"                           # :22-31
    field(req, "type", 1, T.TYPE_ENUM, type_name=".asr.AudioFormat")
    field(req, "end", 2, T.TYPE_INT32)
    field(req, "audio", 3, T.TYPE_BYTES)
    field(req, "device_id", 4, T.TYPE_STRING)
    resp = fd.message_type.add(); resp.name = "TranscribeResponse"                        # :33-38
    field(resp, "end", 1, T.TYPE_INT32)
    field(resp, "text", 2, T.TYPE_BYTES)
    field(resp, "device_id", 3, T.TYPE_STRING)
    field(resp, "segments", 4, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".asr.Segment")
    seg = fd.message_type.add(); seg.name = "Segment"                                     # :40-44
    field(seg, "start", 1, T.TYPE_INT64)
    field(seg, "end", 2, T.TYPE_INT64)
    field(seg, "text", 3, T.TYPE_BYTES)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("asr." + n))      # noqa: E731
    return get("TranscribeRequest"), get("TranscribeResponse"), get("Segment")


TranscribeRequest, TranscribeResponse, Segment = _build_messages()
AUDIO_FORMAT_PCM = 7


def _to_pb(r: _stream.TranscribeResponse):
    return TranscribeResponse(end=r.end, text=r.text, device_id=r.device_id,
                              segments=[Segment(start=s.start, end=s.end, text=s.text) for s in r.segments])


class AsrService:
    """AsrService (asr.rs:20-22,62-66,146-281): one session per Transcribe stream"""

    def __init__(self, engine, session_factory: Optional[Callable] = None):
        self.engine = engine
        self.session_factory = session_factory or (lambda: _stream.AsrStreamSession(engine))

    def Transcribe(self, request_iterator: Iterable, context) -> Iterator:
        try:
            session = self.session_factory()
        except Exception as e:      # noqa: BLE001   asr.rs:163-166: Status::internal(e.to_string())
            context.abort(grpc.StatusCode.INTERNAL, str(e))
            return
        try:
            for req in request_iterator:
                for r in session.feed(req.audio, req.end, req.device_id):
                    yield _to_pb(r)
                if req.end == 1 and len(session.audio_buffer) > 0:      # asr.rs:264: the handler leaves its loop after the tail
                    break
        finally:
            session.close()


def serve(engine, address: str = "127.0.0.1:7300", max_workers: int = 8, session_factory: Optional[Callable] = None):
    """start a grpc.Server speaking proto/asr.proto on `address` (the reference listens on GRPC_ADDR, src/main.rs).
    `engine` may be a BatchingEngine (speaksense_b200/batching.py): chunks of concurrent streams that are ready together then
    share one batched decoder step per token."""
    service = AsrService(engine, session_factory)
    handler = grpc.method_handlers_generic_handler(SERVICE, {
        "Transcribe": grpc.stream_stream_rpc_method_handler(
            service.Transcribe, request_deserializer=TranscribeRequest.FromString,
            response_serializer=lambda m: m.SerializeToString())})
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((handler,))
    port = server.add_insecure_port(address)
    server.start()
    server.bound_port = port
    return server


def transcribe_stream(address: str, messages, device_id: str = "client"):
    """client side (examples/asr_client.rs): send (base64 audio, end) messages, yield TranscribeResponse protobufs"""
    with grpc.insecure_channel(address) as ch:
        call = ch.stream_stream(METHOD, request_serializer=lambda m: m.SerializeToString(),
                                response_deserializer=TranscribeResponse.FromString)
        reqs = (TranscribeRequest(type=AUDIO_FORMAT_PCM, end=e, audio=a, device_id=device_id) for a, e in messages)
        for resp in call(reqs):
            yield resp
