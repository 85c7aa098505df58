// asr_host.cpp - the reference's calling pattern (src/grpc/handlers/asr.rs:154-164,198; src/schedule/processors/
// transcribe.rs:69,100,112) written against the C++ mirror of the trait (include/speaksense_asr.hpp):
//
//     g++ -std=c++17 -Iinclude examples/asr_host.cpp -Lspeaksense_b200/lib -lspeaksense_whisper -Wl,-rpath,$PWD/speaksense_b200/lib -o asr_host
//     ./asr_host ggml-large-v3.bin clip.f32 [clip2.f32 ...]      # raw little-endian f32, mono, 16 kHz
//
// One clip: create_state + transcribe_with_state (stream mode, language zh - what both production callers set).  Several
// clips: one state each, one transcribe_batch call.  Without a model argument it only exercises the host-side text rules
// and the error path of WhisperAsr::new (what the CPU test-suite runs: there is no GPU there and the engine has no fallback).
#include <cstdio>
#include <fstream>

#include "speaksense_asr.hpp"

using namespace speaksense;

static std::vector<float> read_f32(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw AsrError(0, std::string("cannot open ") + path);
    const std::streamsize n = f.tellg();
    std::vector<float> v((size_t)n / sizeof(float));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(float)));
    return v;
}

static void print_result(const TranscribeResult &r) {
    for (const TranscribeSegment &s : r.segments)
        std::printf("[%8.0f -> %8.0f] speaker %zu: %s\n", s.start, s.end, s.speaker_id, s.text.c_str());
    std::printf("full_text: %s\n", r.full_text.c_str());
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::printf("promo %d %d\n", (int)WhisperAsr::is_promotional_text("\xe6\xac\xa2\xe8\xbf\x8e\xe8\xae\xa2\xe9\x98\x85"), (int)WhisperAsr::is_promotional_text("hello"));
        std::printf("punct [%s]\n", WhisperAsr::add_punctuation("hello").c_str());
        try {
            WhisperAsr engine(argc > 1 ? argv[1] : "/nonexistent/ggml-model.bin");
            std::printf("opened\n");
        } catch (const AsrError &e) {
            std::printf("error %d: %s\n", e.code, e.what());
        }
        return 0;
    }
    try {
        WhisperAsr engine(argv[1]);                                   // main.rs:38
        AsrParams params = AsrParams::create();
        params.set_language(std::string("zh"));                       // grpc/handlers/asr.rs:154-157
        params.set_stream_mode(true);
        params.set_min_segment_length(5);
        if (argc == 3) {
            StatePtr state = engine.create_state();                   // asr.rs:164
            print_result(engine.transcribe_with_state(state, read_f32(argv[2]), params));      // asr.rs:198
        } else {
            std::vector<StatePtr> states;
            std::vector<std::vector<float>> clips;
            for (int i = 2; i < argc; i++) { states.push_back(engine.create_state()); clips.push_back(read_f32(argv[i])); }
            for (const TranscribeResult &r : engine.transcribe_batch(states, clips, params)) print_result(r);
        }
    } catch (const AsrError &e) {
        std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
