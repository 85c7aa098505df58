// model.h - ggml legacy .bin loader and the packed HBM weight arena.
// Replaces WhisperContext::new_with_params (/root/reference/src/asr/whisper.rs:23); file layout per
// SURVEY.md Appendix A.1.
#pragma once
#include <string>
#include <vector>

#include "common.h"

namespace ss {

struct HParams {
    int n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
};

struct Lin {            // y = W x + b ; W row-major [n_out][n_in] f16 (K-major), b f32 (may be null)
    const __half *w = nullptr;
    const float *b = nullptr;
    int n_out = 0, n_in = 0;
};
struct LNp { const float *w = nullptr, *b = nullptr; };

struct EncLayer { LNp attn_ln; Lin qkv, o; LNp mlp_ln; Lin fc1, fc2; };
struct DecLayer { LNp attn_ln; Lin qkv, o; LNp cross_ln; Lin cq, ckv, co; LNp mlp_ln; Lin fc1, fc2; };

struct Vocab {
    std::vector<std::string> id_to_token;
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, blank;
    bool multilingual;
    int n_lang;
};

int lang_id(const char *code);   // OpenAI order; -1 if unknown
constexpr int kNumLangTable = 100;

struct Model {
    HParams hp{};
    Vocab vocab;
    int device = 0;
    unsigned char *arena = nullptr;   // device
    size_t arena_bytes = 0;
    size_t meta_bytes = 0;            // leading raw-file prefix (header + filters + vocab)
    // mel
    const float *filters = nullptr;   // [n_mels][201]
    const int2 *filt_range = nullptr; // [n_mels] {first nonzero bin, one past last}
    // encoder
    const float *e_pos = nullptr;
    Lin conv1, conv2;                 // weights permuted to [out][k][c] so K index = k*C + c
    std::vector<EncLayer> enc;
    LNp ln_post;
    // decoder
    const float *d_pos = nullptr;
    const __half *tok_emb = nullptr;
    std::vector<DecLayer> dec;
    LNp d_ln;

    ~Model();
};

struct ModelProbe { HParams hp; size_t arena_bytes; uint64_t fnv1a; int eot, beg, n_vocab_strings; };
// host-only: parse + pack like engine_open would, report sizes / checksum (no device)
ModelProbe probe_model(const std::string &path);

// host-only: hparams + vocabulary of a ggml file into `m` (what the host-side decode logic needs; no tensors, no device)
void load_model_meta(const std::string &path, Model &m);

// Build the host image of the arena from a ggml file (rank 0 / single GPU).
std::vector<unsigned char> build_arena_image(const std::string &path);
// Bind a Model to an arena image already resident on `device` (all ranks): parses the meta prefix
// (copied back from the device) and recomputes the same offsets the packer used.
void bind_model(Model &m, unsigned char *d_arena, size_t bytes, int device);

}  // namespace ss
