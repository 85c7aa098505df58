// speaksense_asr.hpp - C++ host mirror of SpeakSense's transcribe trait above the C ABI (speaksense_whisper.h).
//
// The reference's host language is Rust (no rustc / cargo in this image), so the host side above the boundary is given in
// C++ with the reference's own names, argument meaning and error behaviour:
//
//   /root/reference/src/asr/mod.rs:9-42    AsrParams { language, speaker_diarization, stream_mode, min_segment_length }
//   /root/reference/src/asr/mod.rs:44-56   TranscribeSegment { text, speaker_id, start, end } / TranscribeResult
//   /root/reference/src/asr/mod.rs:58-73   trait AsrEngine { create_state, transcribe_with_state, transcribe (fresh state) }
//   /root/reference/src/asr/whisper.rs:16-129  WhisperAsr: new(model_path), create_state, transcribe_with_state
//
// anyhow::Result<T> becomes T or a thrown speaksense::AsrError whose what() is the string the reference would have built
// ("failed to open whisper model: ...", whisper.rs:24; "Failed to create whisper state: ...", :32).  Arc<Mutex<Box<WhisperState>>>
// becomes std::shared_ptr<WhisperState> carrying its own mutex; transcribe_with_state holds it for the whole inference
// (whisper.rs:51-54).  Audio is taken by value like the reference's Vec<f32>.  Header-only; link with -lspeaksense_whisper.
// (speaksense_b200/asr.py is the same mirror in Python - the one the test-suite drives; INTEGRATION.md has the Rust stub.)
#pragma once
#include <algorithm>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "speaksense_whisper.h"

namespace speaksense {

struct AsrError : std::runtime_error {      // anyhow::Error, stringified
    int code;                               // the ss_status behind it (0 for host-side failures)
    AsrError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct AsrParams {                          // mod.rs:9-42
    std::optional<std::string> language;
    bool speaker_diarization = false;
    bool stream_mode = false;
    size_t min_segment_length = 10;
    int beam_size = 0;                      // extension (BASELINE config 5); 0 / 1 == the reference's Greedy{best_of: 5}

    static AsrParams create() { return AsrParams{}; }                       // AsrParams::new()  mod.rs:18-25
    void set_language(std::optional<std::string> l) { language = std::move(l); }
    void set_speaker_diarization(bool enable) { speaker_diarization = enable; }
    void set_stream_mode(bool enable) { stream_mode = enable; }
    void set_min_segment_length(size_t length) { min_segment_length = length; }

    ss_params native() const {
        ss_params p;
        ss_params_default(&p);
        p.language = language ? language->c_str() : nullptr;      // valid while *this lives
        p.speaker_diarization = speaker_diarization;
        p.stream_mode = stream_mode;
        p.min_segment_length = (int)min_segment_length;
        p.beam_size = beam_size;
        return p;
    }
};

struct TranscribeSegment {                  // mod.rs:44-50
    std::string text;
    size_t speaker_id = 0;
    double start = 0, end = 0;              // whisper t0 / t1 in 10 ms ticks, as f64 (whisper.rs:107-108)
    bool operator==(const TranscribeSegment &o) const { return text == o.text && speaker_id == o.speaker_id && start == o.start && end == o.end; }
};

struct TranscribeResult {                   // mod.rs:52-56
    std::vector<TranscribeSegment> segments;
    std::string full_text;
    bool operator==(const TranscribeResult &o) const { return segments == o.segments && full_text == o.full_text; }
};

class WhisperState {                        // Arc<Mutex<Box<whisper_rs::WhisperState>>>
public:
    explicit WhisperState(ss_state *h) : h_(h) {}
    ~WhisperState() { ss_state_free(h_); }
    WhisperState(const WhisperState &) = delete;
    WhisperState &operator=(const WhisperState &) = delete;
    ss_state *handle() const { return h_; }
    std::mutex &mutex() { return mu_; }

private:
    ss_state *h_;
    std::mutex mu_;
};
using StatePtr = std::shared_ptr<WhisperState>;

class AsrEngine {                           // trait AsrEngine: Send + Sync   mod.rs:58-73
public:
    virtual ~AsrEngine() = default;
    virtual StatePtr create_state() = 0;
    virtual TranscribeResult transcribe_with_state(StatePtr state, std::vector<float> audio, const AsrParams &params) = 0;
    virtual TranscribeResult transcribe(std::vector<float> audio, const AsrParams &params) {      // default method: new state each call
        StatePtr state = create_state();
        return transcribe_with_state(std::move(state), std::move(audio), params);
    }
};

class WhisperAsr : public AsrEngine {       // whisper.rs:16-129
public:
    explicit WhisperAsr(const std::string &model_path, int device = 0) {      // WhisperAsr::new  whisper.rs:21-28
        const int rc = ss_engine_open(model_path.c_str(), device, &h_);
        if (rc != 0) throw AsrError(rc, std::string("failed to open whisper model: ") + ss_last_error());
    }
    ~WhisperAsr() override { ss_engine_close(h_); }
    WhisperAsr(const WhisperAsr &) = delete;
    WhisperAsr &operator=(const WhisperAsr &) = delete;
    ss_engine *handle() const { return h_; }

    StatePtr create_state() override {                                         // whisper.rs:30-39
        ss_state *s = nullptr;
        const int rc = ss_state_new(h_, &s);
        if (rc != 0) throw AsrError(rc, std::string("Failed to create whisper state: ") + ss_last_error());
        return std::make_shared<WhisperState>(s);
    }

    TranscribeResult transcribe_with_state(StatePtr state, std::vector<float> audio, const AsrParams &params) override {      // whisper.rs:45-129
        if (!state) throw AsrError(SS_ERR_INVALID, "Failed to lock state: null state");
        std::lock_guard<std::mutex> guard(state->mutex());
        const ss_params p = params.native();
        const int rc = ss_transcribe(h_, state->handle(), audio.data(), audio.size(), &p);
        if (rc != 0) throw AsrError(rc, ss_last_error());
        return read_result(*state);
    }

    // several sessions at once (one batched decoder step per token; results as if each had been transcribed alone)
    std::vector<TranscribeResult> transcribe_batch(const std::vector<StatePtr> &states, const std::vector<std::vector<float>> &audios,
                                                   const AsrParams &params) {
        if (states.size() != audios.size()) throw AsrError(SS_ERR_INVALID, "transcribe_batch: states and audios differ in length");
        std::vector<ss_state *> hs;
        std::vector<const float *> ptrs;
        std::vector<size_t> lens;
        for (size_t i = 0; i < states.size(); i++) {
            if (!states[i]) throw AsrError(SS_ERR_INVALID, "transcribe_batch: null state");
            hs.push_back(states[i]->handle()); ptrs.push_back(audios[i].data()); lens.push_back(audios[i].size());
        }
        std::vector<std::unique_lock<std::mutex>> guards;      // distinct states; locked in address order
        {
            std::vector<WhisperState *> order;
            for (auto &s : states) order.push_back(s.get());
            std::sort(order.begin(), order.end());
            order.erase(std::unique(order.begin(), order.end()), order.end());
            for (WhisperState *s : order) guards.emplace_back(s->mutex());
        }
        const ss_params p = params.native();
        const int rc = ss_transcribe_batch(h_, hs.data(), ptrs.data(), lens.data(), (int)hs.size(), &p);
        if (rc != 0) throw AsrError(rc, ss_last_error());
        std::vector<TranscribeResult> out;
        for (auto &s : states) out.push_back(read_result(*s));
        return out;
    }

    static bool is_promotional_text(const std::string &text) { return ss_is_promotional_text(text.c_str()) != 0; }      // whisper.rs:41-43
    static std::string add_punctuation(const std::string &text) {                                                          // whisper.rs:175-201
        std::string out(text.size() + 8, '\0');
        const int n = ss_add_punctuation(text.c_str(), &out[0], out.size());
        if (n < 0) throw AsrError(n, ss_last_error());
        out.resize((size_t)n);
        return out;
    }

private:
    static TranscribeResult read_result(const WhisperState &st) {
        TranscribeResult r;
        const ss_state *s = st.handle();
        const int n = ss_n_segments(s);
        for (int i = 0; i < n; i++)
            r.segments.push_back(TranscribeSegment{ss_segment_text(s, i), (size_t)ss_segment_speaker_id(s, i), ss_segment_start(s, i), ss_segment_end(s, i)});
        r.full_text = ss_full_text(s);
        return r;
    }
    ss_engine *h_ = nullptr;
};

}  // namespace speaksense
