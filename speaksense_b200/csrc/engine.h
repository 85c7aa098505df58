// engine.h - host side of the engine: sessions (ss_state), KV caches, the whisper_full decode loop,
// segment assembly and the reference's Rust-side post-processing.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "kernels.h"

namespace ss {

struct Engine {
    Model model;
    int device = 0;
    std::string path;
    std::atomic<int> n_states{0};      // live sessions on this device (ss_state_new_on(e, -1) picks the least loaded replica)
    // operand buffers of the batched decoder (engine_batch.cc): one batch at a time per engine
    std::mutex batch_mu;
    void *batch_scratch = nullptr;
    cudaStream_t batch_stream = nullptr;
    cudaEvent_t batch_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // 0, 1: termination polls; 2, 3: timing
    int *batch_h_flags = nullptr;                                     // pinned [2]: finished-sequence counts read back by the polls
    int sms = 0;
    // activations of the batched encoder pass (SS_BATCH_ENCODER, on by default): [clips * n_audio_ctx] rows
    void *enc_scratch = nullptr;
    cudaEvent_t enc_ev[3] = {nullptr, nullptr, nullptr};              // start / end of the pass (timing), done (other streams wait on it)
    // candidates of a batched beam-search step (SS_BATCH_BEAM, on by default): device / pinned [kMaxBatch][8]
    TokData *beam_cand = nullptr, *beam_h_cand = nullptr;
    ~Engine();
};

struct FullParams {   // == build_params (whisper.rs:131-173) + overrides (:60-71)
    std::string language = "en";
    bool tdrz_enable = false, no_context = false, single_segment = false;
    int best_of = 5, beam_size = 0;
    float temperature = 0.0f, temperature_inc = 0.2f, entropy_thold = 2.4f, logprob_thold = -1.0f;
    float max_initial_ts = 1.0f, length_penalty = -1.0f;
    bool suppress_blank = true;
    int n_max_text_ctx = 16384, max_tokens = 0;
    bool keep_logits = false;
};

constexpr int kMaxDecoders = 8;   // WHISPER_MAX_DECODERS

struct Sequence {
    std::vector<TokData> tokens;
    int result_len = 0;
    double sum_logprobs_all = 0, sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
};

struct Decoder {
    MegaParams mp{};                  // host copy of the device-resident descriptor
    MegaParams *d_mp = nullptr;
    void *d_ll = nullptr; size_t ll_bytes = 0;   // flagged exchange arena
    bool mp_dirty = true;
    DecCtl *h_ctl = nullptr;          // pinned mirror
    TokData *h_tok = nullptr;         // pinned
    Sequence seq;
    int seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;
    std::vector<float> probs, logits, logprobs;   // host-sampled fallback / beam path
    __half *alt_k = nullptr, *alt_v = nullptr;    // second self-KV buffer: target of the beam-search cache shuffle
    std::mt19937 rng{0};
};

struct RawSegment { int64_t t0, t1; std::string text; bool speaker_turn_next; };
struct OutSegment { std::string text; int speaker_id; double start, end; };

struct State {
    std::shared_ptr<Engine> engine;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // audio / mel
    float *d_pcm = nullptr; size_t pcm_cap = 0; size_t n_resident = 0; float *h_pcm = nullptr; size_t h_pcm_cap = 0;
    float *d_pcm_alt = nullptr; float *d_dn = nullptr; size_t dn_cap = 0;   // denoise: output buffer (swapped with d_pcm), scratch
    float *d_mel = nullptr; size_t mel_cap = 0; int n_len = 0, n_len_org = 0; int *d_max = nullptr;
    // encoder scratch (one window)
    __half *win = nullptr, *x1 = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *ff = nullptr, *enc16 = nullptr;
    float *x = nullptr, *enc_out = nullptr;
    __half *cross_k = nullptr, *cross_v = nullptr;
    // the encoder pass of one window (conv stem .. cross-KV: ~230 launches over buffers that never move) captured as a CUDA graph at
    // the first window and replayed afterwards (SS_ENC_GRAPH=0: plain launches)
    void *enc_graph = nullptr; int enc_graph_launches = 0;
    // decoders
    std::vector<std::unique_ptr<Decoder>> dec;
    float *keep = nullptr; int keep_cap = 0; std::vector<float> h_keep; int n_keep = 0;
    float *h_logits = nullptr;   // pinned [n_vocab]
    int mega_grid = 0;
    // results
    std::vector<int> prompt_past;
    std::vector<RawSegment> raw;
    std::vector<OutSegment> out;
    std::string full_text;
    std::vector<TokData> result_tokens;
    int n_fallbacks = 0, n_decoded = 0, n_windows = 0, n_launches = 0;
    int last_status = 0; std::string last_error;      // outcome of the last transcribe on this state (a batch call fails clip by clip)
    float ms_mel = 0, ms_enc = 0, ms_dec = 0;

    ~State();
};

std::shared_ptr<Engine> engine_open(const std::string &path, int device);
std::shared_ptr<Engine> engine_open_dist(const char *path, int device, int rank, int world, const unsigned char *nccl_id);
// one process, several GPUs (the reference is ONE server process: main.rs:38-39): the file is parsed once, the arena goes to
// devices[0] and from there to the others with an in-process ncclBroadcast (ncclCommInitAll + group call) over NVLink
std::vector<std::shared_ptr<Engine>> engine_open_multi(const std::string &path, const int *devices, int n_devices);
void nccl_unique_id(unsigned char out[128]);
State *state_new(const std::shared_ptr<Engine> &e);

void upload_pcm(State &s, const float *pcm, size_t n);
// reference denoise_audio on the device: host PCM in, the denoised chunk stays resident (and is copied to `out` if given)
int denoise_audio(State &s, const float *pcm, size_t n, int frame_size, float overlap, float strength, float *out, float *nv_out);
// StreamAudioProcessor step 4-5 for `n_frames` frames (host in / host out)
void denoise_frames(State &s, const float *frames, int n_frames, int frame_size, float strength, float noise_gate, float *out);
void run_log_mel(State &s, const float *pcm, size_t n);   // pcm == nullptr: use the resident PCM
float bench_decode_steps(State &s, int n_steps, int n_past0);
void run_encode(State &s, int seek);
void run_decode_forced(State &s, const int *tokens, int n, int n_past, float *logits_out);
int transcribe(State &s, const float *pcm, size_t n, const FullParams &fp, bool stream_mode);
// ss_transcribe_batch: the clips advance window by window together, their temperature-0 greedy decodes share one batched
// decoder step (decoder_batch.cu); fallbacks run per clip.  Results land in each State exactly as transcribe() leaves them.
int transcribe_batch(State *const *states, const float *const *pcm, const size_t *n, int batch, const FullParams &fp, bool stream_mode);
bool batch_decode_enabled();   // default on; SS_BATCH_DECODE=0 disables

// Rust-side post-processing of whisper.rs:84-128 on s.raw -> s.out / s.full_text
int postprocess(State &s, bool stream_mode);
bool is_valid_utf8(const std::string &t);
bool is_promotional_text(const std::string &t);
std::string add_punctuation(const std::string &text);

}  // namespace ss
