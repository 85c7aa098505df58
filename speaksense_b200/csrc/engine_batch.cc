// engine_batch.cc - ss_transcribe_batch: several clips through `state.full` together (BASELINE.json configs 3/4).
//
// The clips of a batch are independent sessions (one ss_state each, exactly as the reference would hold one
// WhisperState per stream / task: /root/reference/src/asr/whisper.rs:30-39,75), so the control flow of whisper_full
//  stays per clip; what is shared is the expensive part of every window - the temperature-0 greedy
// decode - which runs as ONE batched decoder step per token for all clips that are in a window (decoder_batch.cu:
// the weights are streamed once per step instead of once per clip).  Rounds:
//     every unfinished clip: encode its current window (own stream), arm its decoder-0 control block
//     all of them together : batched greedy decode until every sequence has completed / failed / hit the cap
//     every clip           : whisper_full's scoring and success test; clips that fail walk the temperature ladder
//                            (5 sampled decoders, host sampling) on their own as in engine.cc; segments; seek advance
// Beam search, kept logits and a non-zero base temperature take the clip-by-clip path (engine.cc `transcribe`).
//
// The per-window logic below restates the corresponding parts of `transcribe` in engine.cc instead of sharing them: this
// file was written at the very end of round 1 and `transcribe` was left untouched as the verified single-clip path.  The
// batched path is parity-green against it (tests/test_gpu_batch.py); folding the two into one is round-2 housekeeping.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "engine_internal.h"

#include <random>

namespace ss {

Engine::~Engine() {
    if (batch_scratch || batch_stream || batch_h_flags) cudaSetDevice(device);
    if (batch_stream) { cudaStreamSynchronize(batch_stream); cudaStreamDestroy(batch_stream); }
    for (auto &e : batch_ev) if (e) cudaEventDestroy(e);
    if (batch_scratch) cudaFree(batch_scratch);
    if (batch_h_flags) cudaFreeHost(batch_h_flags);
    if (enc_scratch) { cudaSetDevice(device); cudaFree(enc_scratch); }
    if (beam_cand) { cudaSetDevice(device); cudaFree(beam_cand); }
    if (beam_h_cand) cudaFreeHost(beam_h_cand);
    for (auto &e : enc_ev) if (e) cudaEventDestroy(e);
}

bool batch_decode_enabled() {      // on by default; SS_BATCH_DECODE=0 gives ss_transcribe_batch the clip-by-clip decode back
    const char *e = getenv("SS_BATCH_DECODE");
    return !(e && e[0] == '0');
}

// Smallest batch that takes the batched decoder.  A batched step is a chain of ~390 short kernels (~2 ms per step at large-v3
// whatever the batch), the batch-1 persistent kernel needs 0.71 ms per token and clip: below ~3 clips clip-by-clip is the
// faster way (estimate from profiles/r1e_batch_launches_summary.txt; to be re-measured: tools/gpu_round_check.sh).
static int batch_decode_min() {
    const char *e = getenv("SS_BATCH_MIN");
    const int v = e ? atoi(e) : 4;
    return std::max(2, v);
}

namespace {

constexpr int kXsplitMax = 8;
// Decoder steps between two reads of the finished-sequence count.  The host enqueues at most two such groups ahead of the count it has
// seen, so between kPollEvery and 3 * kPollEvery - 1 steps run after the last sequence has finished (their kernels leave early, but a
// step of ~390 of them still costs ~1 ms).  2 keeps >= 4 steps (>= 6 ms of device work at any batch size) queued and wastes 2..5 steps
// instead of 4..11; SS_BATCH_POLL overrides (1..16).
static int poll_every() {
    static const int v = [] { const char *e = getenv("SS_BATCH_POLL"); const int x = e ? atoi(e) : 2; return x >= 1 && x <= 16 ? x : 2; }();
    return v;
}
#define kPollEvery (poll_every())

struct ClipRun {
    State *s = nullptr;
    int seek = 0, seek_end = 0;
    bool finished = false;
    std::vector<int> prompt;       // prompt of the temperature in progress
    int best_decoder_id = 0;
};

struct Common {      // what whisper_full derives once per call
    std::vector<float> temps;
    std::vector<int> prompt_init;
    int n_decoders = 1, n_max = 0, tid0_init = -1;
};

void ensure_batch_resources(Engine &E) {
    if (E.batch_scratch) return;
    const HParams &hp = E.model.hp;
    const size_t bytes = decode_batch_scratch_bytes(hp.n_text_state, hp.n_vocab, hp.n_text_head, kXsplitMax);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) SS_THROW(-5, "cudaMalloc of %zu bytes (batched decoder operands) failed", bytes);
    CUDA_CHECK(cudaMemset(p, 0, bytes));      // rows of absent sequences stay finite (they are multiplied, never stored)
    E.batch_scratch = p;
    CUDA_CHECK(cudaStreamCreateWithFlags(&E.batch_stream, cudaStreamNonBlocking));
    for (auto &e : E.batch_ev) CUDA_CHECK(cudaEventCreate(&e));
    if (cudaMallocHost(&E.batch_h_flags, 2 * sizeof(int)) != cudaSuccess) SS_THROW(-5, "cudaMallocHost failed");
    CUDA_CHECK(cudaDeviceGetAttribute(&E.sms, cudaDevAttrMultiProcessorCount, E.device));
}

void build_prompt(const State &s, const FullParams &P, const Common &C, float t_cur, std::vector<int> &prompt) {
    const HParams &hp = s.engine->model.hp; const Vocab &v = s.engine->model.vocab;
    prompt.clear();
    if (!s.prompt_past.empty() && t_cur < 0.5f && P.n_max_text_ctx > 0) {
        const int n_take = std::min({P.n_max_text_ctx, hp.n_text_ctx / 2, (int)s.prompt_past.size()});
        prompt.push_back(v.prev);
        prompt.insert(prompt.end(), s.prompt_past.end() - n_take, s.prompt_past.end());
    }
    prompt.insert(prompt.end(), C.prompt_init.begin(), C.prompt_init.end());
}

void reset_decoders(State &s, int n_cur) {
    while ((int)s.dec.size() < n_cur) new_decoder(s, false);
    for (int j = 0; j < n_cur; j++) {
        Decoder &dc = *s.dec[j];
        dc.seq = Sequence{}; dc.seq.sum_logprobs = -INFINITY; dc.seq.avg_logprobs = -INFINITY; dc.seq.score = -INFINITY;
        dc.seek_delta = 100 * kChunkSec; dc.failed = false; dc.completed = false; dc.has_ts = false;
    }
}

// whisper_full's ranking of the decoders of one temperature and its success test; returns true when the window is settled
bool score_and_test(State &s, const FullParams &P, const Common &C, size_t it, int n_cur, int &best_decoder_id) {
    double best_score = -INFINITY;
    for (int j = 0; j < n_cur; j++) {
        Decoder &dc = *s.dec[j];
        if (dc.failed) continue;
        dc.seq.tokens.resize(std::min<size_t>(dc.seq.tokens.size(), (size_t)dc.seq.result_len));
        sequence_score(P, dc.seq);
        if (dc.seq.entropy < P.entropy_thold) { dc.failed = true; continue; }
        if (best_score < dc.seq.score) { best_score = dc.seq.score; best_decoder_id = j; }
    }
    if (it != C.temps.size() - 1) {
        const Decoder &dc = *s.dec[best_decoder_id];
        if (dc.failed || dc.seq.avg_logprobs < P.logprob_thold) return false;
    }
    return true;
}

// t > 0: best_of sampled decoders of one clip, host-side sampling (== the non-beam fallback branch of engine.cc)
void decode_sampled(State &s, const FullParams &P, const Common &C, float t_cur, int n_cur, const std::vector<int> &prompt, int seek, int seek_end) {
    const Model &m = s.engine->model; const Vocab &v = m.vocab;
    const int n_prompt = (int)prompt.size(), n_max = C.n_max;
    for (int j = 0; j < n_cur; j++) set_sampling(*s.dec[j], P, C.tid0_init);
    step_host_sampled(s, *s.dec[0], prompt.data(), n_prompt, 0);
    s.n_decoded += n_prompt - 1;
    process_logits_host(m, P, *s.dec[0], s.h_logits, t_cur);
    for (int j = 1; j < n_cur; j++) {
        kv_copy(s, *s.dec[0], *s.dec[j], n_prompt);
        s.dec[j]->probs = s.dec[0]->probs; s.dec[j]->logits = s.dec[0]->logits; s.dec[j]->logprobs = s.dec[0]->logprobs;
    }
    for (int i = 0; i < n_max; i++) {
        for (int j = 0; j < n_cur; j++) {
            Decoder &dc = *s.dec[j];
            if (dc.completed || dc.failed) continue;
            dc.seq.tokens.push_back(sample_token_host(m, dc, false));
            dc.seq.sum_logprobs_all += dc.seq.tokens.back().plog;
        }
        for (int j = 0; j < n_cur; j++) {
            Decoder &dc = *s.dec[j];
            if (dc.completed || dc.failed) continue;
            const TokData &tk = dc.seq.tokens.back();
            if (tk.id > v.beg) {
                const int sd_new = 2 * (tk.id - v.beg);
                if (dc.has_ts && dc.seek_delta > sd_new && dc.seq.result_len < i) { dc.failed = true; continue; }
                dc.seek_delta = sd_new; dc.seq.result_len = i + 1; dc.has_ts = true;
            }
            if (tk.id == v.eot || (P.max_tokens > 0 && i >= P.max_tokens) || (dc.has_ts && seek + dc.seek_delta + 100 >= seek_end)) {
                if (dc.seq.result_len == 0) {
                    if (seek + dc.seek_delta + 100 >= seek_end) dc.seq.result_len = i + 1;
                    else { dc.failed = true; continue; }
                }
                if (P.single_segment) { dc.seq.result_len = i + 1; dc.seek_delta = 100 * kChunkSec; }
                dc.completed = true; continue;
            }
            if (i == n_max - 1 && (dc.seq.result_len == 0 || dc.seek_delta < 100 * kChunkSec / 2)) { dc.failed = true; continue; }
        }
        bool all = true;
        for (int j = 0; j < n_cur; j++) if (!(s.dec[j]->completed || s.dec[j]->failed)) all = false;
        if (all) break;
        const int n_past = n_prompt + i;
        for (int j = 0; j < n_cur; j++) {
            Decoder &dc = *s.dec[j];
            if (dc.failed || dc.completed) continue;
            const int tok = dc.seq.tokens.back().id;
            step_host_sampled(s, dc, &tok, 1, n_past);
            process_logits_host(m, P, dc, s.h_logits, t_cur);
        }
    }
}

// the winner's tokens -> prompt_past, result tokens, raw segments; returns seek_delta
int emit_window(State &s, const FullParams &P, const Common &C, const ClipRun &r) {
    const Vocab &v = s.engine->model.vocab;
    const Decoder &bd = *s.dec[r.best_decoder_id];
    const int seek = r.seek, seek_delta = bd.seek_delta, result_len = bd.seq.result_len;
    const auto &tc = bd.seq.tokens;
    const std::vector<int> &prompt = r.prompt;
    std::vector<int> keep;
    if (prompt.front() == v.prev) keep.assign(prompt.begin() + 1, prompt.end() - C.prompt_init.size());
    s.prompt_past = keep;
    for (int i = 0; i < result_len && i < (int)tc.size(); i++) s.prompt_past.push_back(tc[i].id);
    if (!tc.empty()) {
        s.result_tokens.insert(s.result_tokens.end(), tc.begin(), tc.end());
        int64_t t0 = seek + 2 * (tc.front().tid - v.beg);
        std::string text; bool turn = false;
        for (int i = 0; i < (int)tc.size(); i++) {
            if (tc[i].id < v.eot) text += v.id_to_token[tc[i].id];
            if (P.tdrz_enable && tc[i].id == v.solm) turn = true;
            if (tc[i].id > v.beg && !P.single_segment) {
                const int64_t t1 = seek + 2 * (tc[i].tid - v.beg);
                if (!text.empty()) s.raw.push_back({t0, t1, text, turn});
                text.clear();
                while (i < (int)tc.size() && tc[i].id > v.beg) i++;
                i--;
                t0 = t1; turn = false;
            }
        }
        if (!text.empty()) s.raw.push_back({t0, (int64_t)seek + seek_delta, text, turn});
    }
    return seek_delta;
}

// ---- batched encoder pass (default; SS_BATCH_ENCODER=0 turns it off) ----
// The windows of all clips of a round as ONE pass over [clips * 1500] rows: the conv stem and the cross-KV projection stay
// per clip (their operands live in each State), the 32 layers run once with M = clips * 1500 - the N = 1280 GEMMs, 120 tiles
// on 148 SMs for one clip, become 375 * clips tiles - and the fused attention kernel takes the clip as its outer batch.
// SS_BATCH_ENCODER=0: the clips' encoders run concurrently on their own streams (run_encode) - round 1's path; measured on
// 32 x 30 s clips (profiles/r2a_batch_bench*.json): log-mel + encoders 157 ms per clip stream -> 131 ms batched.
bool batch_encoder_enabled() {
    const char *e = getenv("SS_BATCH_ENCODER");
    return !(e && e[0] == '0');
}

struct EncBuf { float *x, *enc_out; __half *xn, *qkv, *att, *ff, *enc16; };

size_t a256(size_t n) { return (n + 255) & ~(size_t)255; }

EncBuf bind_enc_scratch(Engine &E) {
    const HParams &hp = E.model.hp;
    const size_t rows = (size_t)kMaxBatch * hp.n_audio_ctx, d = hp.n_audio_state;
    const size_t sz[7] = {a256(rows * d * 4), a256(rows * d * 4), a256(rows * d * 2), a256(rows * 3 * d * 2), a256(rows * d * 2),
                          a256(rows * 4 * d * 2), a256(rows * d * 2)};
    if (!E.enc_scratch) {
        size_t total = 0;
        for (size_t b : sz) total += b;
        if (cudaMalloc(&E.enc_scratch, total) != cudaSuccess) SS_THROW(-5, "cudaMalloc of %zu bytes (batched encoder activations) failed", total);
        for (auto &e : E.enc_ev) CUDA_CHECK(cudaEventCreate(&e));
    }
    uint8_t *p = static_cast<uint8_t *>(E.enc_scratch);
    EncBuf b;
    b.x = reinterpret_cast<float *>(p); p += sz[0];
    b.enc_out = reinterpret_cast<float *>(p); p += sz[1];
    b.xn = reinterpret_cast<__half *>(p); p += sz[2];
    b.qkv = reinterpret_cast<__half *>(p); p += sz[3];
    b.att = reinterpret_cast<__half *>(p); p += sz[4];
    b.ff = reinterpret_cast<__half *>(p); p += sz[5];
    b.enc16 = reinterpret_cast<__half *>(p);
    return b;
}

// encoder + cross-KV of the current window of every clip of `grp` (<= kMaxBatch) on the engine's batch stream; every clip's
// own stream is made to wait for the pass.  Same arithmetic, kernels and epilogues as run_encode (engine.cc).
void run_encode_batch(Engine &E, std::vector<ClipRun *> &grp) {
    NvtxRange nvtx("ss.batch.encode");
    const Model &m = E.model; const HParams &hp = m.hp;
    const int R = (int)grp.size(), T = hp.n_audio_ctx, d = hp.n_audio_state, H = hp.n_audio_head, C = hp.n_mels;
    const EncBuf B = bind_enc_scratch(E);
    cudaStream_t st = E.batch_stream;
    int launches = 0; int *nl = &launches;
    CUDA_CHECK(cudaEventRecord(E.enc_ev[0], st));
    for (int b = 0; b < R; b++) {      // conv stem per clip -> rows [b * T, (b + 1) * T) of the residual stream
        State &s = *grp[b]->s;
        mel_window_enqueue(m, s.d_mel, s.n_len, grp[b]->seek, s.win, st, nl);
        {
            GemmOperand A; A.ptr = s.win; A.rows = 2 * T; A.ld = C;
            GemmOperand W; W.ptr = m.conv1.w; W.rows = d; W.ld = 3 * C;
            GemmEpilogue ep; ep.bias = m.conv1.b; ep.gelu = 1; ep.out = s.x1; ep.out_type = GEMM_OUT_F16; ep.out_ld = d; ep.out_row_offset = 1;
            gemm_enqueue(A, W, 2 * T, d, 3 * C, false, ep, st, nl);
        }
        {
            GemmOperand A; A.ptr = s.x1; A.rows = T; A.ld = 2 * d;
            GemmOperand W; W.ptr = m.conv2.w; W.rows = d; W.ld = 3 * d;
            GemmEpilogue ep; ep.bias = m.conv2.b; ep.gelu = 1; ep.pos = m.e_pos; ep.pos_rows = T; ep.out = B.x + (size_t)b * T * d; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
            gemm_enqueue(A, W, T, d, 3 * d, false, ep, st, nl);
        }
    }
    const int M = R * T;
    for (int il = 0; il < hp.n_audio_layer; il++) {
        const EncLayer &L = m.enc[il];
        layernorm_f16_enqueue(B.x, B.xn, M, d, L.attn_ln, st, nl);
        {
            GemmOperand A; A.ptr = B.xn; A.rows = M; A.ld = d;
            GemmOperand W; W.ptr = L.qkv.w; W.rows = 3 * d; W.ld = d;
            GemmEpilogue ep; ep.bias = L.qkv.b; ep.out = B.qkv; ep.out_ld = 3 * d;
            gemm_enqueue(A, W, M, 3 * d, d, false, ep, st, nl);
        }
        attention_enqueue(B.qkv, B.att, R, T, H, 1.0f / sqrtf(64.0f), st, nl);
        {
            GemmOperand A; A.ptr = B.att; A.rows = M; A.ld = d;
            GemmOperand W; W.ptr = L.o.w; W.rows = d; W.ld = d;
            GemmEpilogue ep; ep.bias = L.o.b; ep.residual = 1; ep.out = B.x; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
            gemm_enqueue(A, W, M, d, d, false, ep, st, nl);
        }
        layernorm_f16_enqueue(B.x, B.xn, M, d, L.mlp_ln, st, nl);
        {
            GemmOperand A; A.ptr = B.xn; A.rows = M; A.ld = d;
            GemmOperand W; W.ptr = L.fc1.w; W.rows = 4 * d; W.ld = d;
            GemmEpilogue ep; ep.bias = L.fc1.b; ep.gelu = 1; ep.out = B.ff; ep.out_ld = 4 * d;
            gemm_enqueue(A, W, M, 4 * d, d, false, ep, st, nl);
        }
        {
            GemmOperand A; A.ptr = B.ff; A.rows = M; A.ld = 4 * d;
            GemmOperand W; W.ptr = L.fc2.w; W.rows = d; W.ld = 4 * d;
            GemmEpilogue ep; ep.bias = L.fc2.b; ep.residual = 1; ep.out = B.x; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
            gemm_enqueue(A, W, M, d, 4 * d, false, ep, st, nl);
        }
    }
    layernorm_f32_enqueue(B.x, B.enc_out, M, d, m.ln_post, st, nl);
    f32_to_f16_enqueue(B.enc_out, B.enc16, (size_t)M * d, st, nl);
    const int dd = hp.n_text_state, Ld = hp.n_text_layer;
    const float s4 = powf((float)(dd / hp.n_text_head), -0.25f);
    const long wstride = Ld > 1 ? (long)(m.dec[1].ckv.w - m.dec[0].ckv.w) : (long)2 * dd * d;
    const long bstride = Ld > 1 ? (long)(m.dec[1].ckv.b - m.dec[0].ckv.b) : (long)2 * dd;
    for (int il = 0; il < Ld; il++)
        if (m.dec[il].ckv.w != m.dec[0].ckv.w + (long)il * wstride || m.dec[il].ckv.b != m.dec[0].ckv.b + (long)il * bstride)
            SS_THROW(-9, "decoder layers are not equally spaced in the weight arena");
    for (int b = 0; b < R; b++) {      // cross-attention K / V of every decoder layer into the clip's own cache
        State &s = *grp[b]->s;
        GemmOperand A; A.ptr = B.enc16 + (size_t)b * T * d; A.rows = T; A.ld = d;
        GemmOperand W; W.ptr = m.dec[0].ckv.w; W.rows = 2 * dd; W.ld = d; W.batch0 = Ld; W.stride0 = wstride;
        GemmEpilogue ep; ep.a_broadcast = 1; ep.bias = m.dec[0].ckv.b; ep.bias_stride0 = bstride;
        ep.alpha = s4; ep.alpha_cols = dd;
        ep.out = s.cross_k; ep.head_major = 1; ep.head_rows = T; ep.out_stride0 = (long)2 * T * dd;
        gemm_enqueue(A, W, T, 2 * dd, d, false, ep, st, nl);
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(E.enc_ev[1], st));
    CUDA_CHECK(cudaEventRecord(E.enc_ev[2], st));
    for (int b = 0; b < R; b++) {
        State &s = *grp[b]->s;
        CUDA_CHECK(cudaStreamWaitEvent(s.stream, E.enc_ev[2], 0));
        s.n_launches += ceil_div(launches, R);
    }
}

// batched greedy decode of the armed decoder-0 sequences of `grp` (<= kMaxBatch clips)
void decode_group(Engine &E, const FullParams &P, const Common &C, std::vector<ClipRun *> &grp) {
    const Model &m = E.model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    const int nb = (int)grp.size();
    BatchParams bp{};
    bp.B = nb; bp.d = hp.n_text_state; bp.H = hp.n_text_head; bp.L = hp.n_text_layer; bp.T = hp.n_audio_ctx; bp.ctx = hp.n_text_ctx; bp.n_vocab = hp.n_vocab;
    bp.s4 = powf((float)(bp.d / bp.H), -0.25f);
    bp.tok_emb = m.tok_emb; bp.d_pos = m.d_pos; bp.lnf_w = m.d_ln.w; bp.lnf_b = m.d_ln.b;
    bp.eot = v.eot; bp.sot = v.sot; bp.translate = v.translate; bp.transcribe = v.transcribe; bp.solm = v.solm; bp.prev = v.prev;
    bp.nosp = v.nosp; bp.not_ = v.not_; bp.beg = v.beg; bp.blank = v.blank;
    bp.suppress_blank = P.suppress_blank; bp.tdrz = P.tdrz_enable; bp.tid0_init = C.tid0_init;
    decode_batch_bind(bp, E.batch_scratch);
    cudaStream_t bs = E.batch_stream;
    int max_steps = 0, min_prompt = 1 << 30;
    for (int k = 0; k < nb; k++) {
        State &s = *grp[k]->s; Decoder &dc = *s.dec[0];
        bp.seq[k] = BatchSeq{dc.mp.ctl, dc.mp.self_k, dc.mp.self_v, s.cross_k, s.cross_v, dc.mp.tok_out};
        CUDA_CHECK(cudaStreamWaitEvent(bs, s.ev[2], 0));      // its encoder, cross-KV and control block are on its own stream
        const int np = (int)grp[k]->prompt.size();
        max_steps = std::max(max_steps, np + C.n_max - 1); min_prompt = std::min(min_prompt, np);
    }
    for (int k = nb; k < kMaxBatch; k++) bp.seq[k] = bp.seq[0];      // never dereferenced (b < B guards); keeps the block defined
    const MegaParams &w = grp[0]->s->dec[0]->mp;                     // weight pointers (identical for every decoder of the engine)
    const int xsplit = decode_batch_xsplit(nb, bp.H, E.sms);
    CUDA_CHECK(cudaMemsetAsync(bp.n_done, 0, sizeof(int), bs));
    CUDA_CHECK(cudaEventRecord(E.batch_ev[2], bs));
    int launches = 0;
    BatchStepGraph step_graph;      // the step captured once, replayed per token
    E.batch_h_flags[0] = E.batch_h_flags[1] = 0;
    for (int t = 0; t < max_steps; t++) {
        if (t % kPollEvery == 0) {
            const int G = t / kPollEvery;
            if (G >= 2) {      // the count read back two groups ago: at most 2 * kPollEvery steps of early-exit kernels are wasted
                CUDA_CHECK(cudaEventSynchronize(E.batch_ev[G & 1]));
                if (E.batch_h_flags[G & 1] >= nb) break;
            }
        }
        decode_batch_step_graph(step_graph, bp, w, /*need_logits=*/t >= min_prompt - 1, xsplit, bs, &launches);
        if (t % kPollEvery == kPollEvery - 1) {
            const int G = t / kPollEvery;
            CUDA_CHECK(cudaMemcpyAsync(&E.batch_h_flags[G & 1], bp.n_done, sizeof(int), cudaMemcpyDeviceToHost, bs));
            CUDA_CHECK(cudaEventRecord(E.batch_ev[G & 1], bs));
        }
    }
    CUDA_CHECK(cudaEventRecord(E.batch_ev[3], bs));
    CUDA_CHECK(cudaStreamSynchronize(bs));
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, E.batch_ev[2], E.batch_ev[3]));
    // results back into each clip's decoder 0, as the device-side greedy path of engine.cc leaves them
    for (int k = 0; k < nb; k++) {
        State &s = *grp[k]->s; Decoder &dc = *s.dec[0];
        CUDA_CHECK(cudaMemcpyAsync(dc.h_ctl, dc.mp.ctl, offsetof(DecCtl, prompt), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        const int ns = dc.h_ctl->n_sampled, np = (int)grp[k]->prompt.size();
        if (ns > 0) {
            CUDA_CHECK(cudaMemcpyAsync(dc.h_tok, dc.mp.tok_out, (size_t)ns * sizeof(TokData), cudaMemcpyDeviceToHost, s.stream));
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
        }
        dc.seq.tokens.assign(dc.h_tok, dc.h_tok + ns);
        for (int i = 0; i < ns; i++) dc.seq.sum_logprobs_all += dc.h_tok[i].plog;
        dc.seq.result_len = dc.h_ctl->result_len; dc.seek_delta = dc.h_ctl->seek_delta;
        dc.failed = dc.h_ctl->failed; dc.completed = dc.h_ctl->completed; dc.has_ts = dc.h_ctl->has_ts;
        s.n_decoded += np - 1 + ns;
        s.n_launches += ceil_div(launches, nb);
        s.ms_dec += ms / nb;      // the batch's device time, shared equally
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// beam search on the batched step (default; SS_BATCH_BEAM=0 turns it off).
// The SS_BATCH_BEAM=0 path (engine.cc) launches the batch-1 kernel once per live beam and token and filters / sorts 51 866
// logits per beam on the host; here the live beams of the window are the sequences of ONE batched step (they share the
// state's cross-KV cache, each decoder keeps its own self-KV cache and control block) and bd_topk_kernel leaves k
// candidates per beam.  Candidate ranking, duplicate skipping, the self-KV shuffle (beam_advance) and the per-token
// bookkeeping stay on the host, exactly as in engine.cc.
// ------------------------------------------------------------------------------------------------
bool batch_beam_enabled() {
    const char *e = getenv("SS_BATCH_BEAM");
    return !(e && e[0] == '0');
}
bool batch_beam_supported(const State &s) {
    const HParams &hp = s.engine->model.hp;
    return hp.n_text_state == hp.n_text_head * 64 && hp.n_text_state <= 1280 && hp.n_text_ctx <= 512 && hp.n_audio_ctx <= 1536;
}

void decode_beam_batched(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek, int seek_end,
                         int n_max, int tid0_init) {
    Engine &E = *s.engine;
    const Model &m = E.model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    std::lock_guard<std::mutex> lk(E.batch_mu);      // the operand buffers are the engine's
    ensure_batch_resources(E);
    if (!E.beam_cand) {
        if (cudaMalloc(&E.beam_cand, (size_t)kMaxBatch * 8 * sizeof(TokData)) != cudaSuccess) SS_THROW(-5, "cudaMalloc failed");
        if (cudaMallocHost(&E.beam_h_cand, (size_t)kMaxBatch * 8 * sizeof(TokData)) != cudaSuccess) SS_THROW(-5, "cudaMallocHost failed");
    }
    const int n_prompt = (int)prompt.size(), K = P.beam_size;
    // the prompt goes through the batch-1 kernel on decoder 0 and is filtered on the host once, as in engine.cc
    for (int j = 0; j < n_cur; j++) set_sampling(*s.dec[j], P, tid0_init);
    step_host_sampled(s, *s.dec[0], prompt.data(), n_prompt, 0);
    s.n_decoded += n_prompt - 1;
    process_logits_host(m, P, *s.dec[0], s.h_logits, t_cur);
    for (int j = 1; j < n_cur; j++) {
        kv_copy(s, *s.dec[0], *s.dec[j], n_prompt);
        s.dec[j]->probs = s.dec[0]->probs; s.dec[j]->logits = s.dec[0]->logits; s.dec[j]->logprobs = s.dec[0]->logprobs;
    }
    BatchParams bp{};
    bp.d = hp.n_text_state; bp.H = hp.n_text_head; bp.L = hp.n_text_layer; bp.T = hp.n_audio_ctx; bp.ctx = hp.n_text_ctx; bp.n_vocab = hp.n_vocab;
    bp.s4 = powf((float)(bp.d / bp.H), -0.25f);
    bp.tok_emb = m.tok_emb; bp.d_pos = m.d_pos; bp.lnf_w = m.d_ln.w; bp.lnf_b = m.d_ln.b;
    bp.eot = v.eot; bp.sot = v.sot; bp.translate = v.translate; bp.transcribe = v.transcribe; bp.solm = v.solm; bp.prev = v.prev;
    bp.nosp = v.nosp; bp.not_ = v.not_; bp.beg = v.beg; bp.blank = v.blank;
    bp.suppress_blank = P.suppress_blank; bp.tdrz = P.tdrz_enable; bp.tid0_init = tid0_init;
    decode_batch_bind(bp, E.batch_scratch);
    const MegaParams &w = s.dec[0]->mp;
    const BeamStep bstep{t_cur, K, E.beam_cand};
    std::vector<std::vector<TokData>> dev_cands(n_cur);      // candidates of decoder j from the last batched step
    BatchStepGraph step_graph;      // re-captured only when the set of live beams changes
    std::vector<BeamCandidate> cands;
    for (int i = 0; i < n_max; i++) {
        cands.clear();
        for (int j = 0; j < n_cur; j++) {
            Decoder &dc = *s.dec[j];
            if (dc.completed || dc.failed) continue;
            const std::vector<TokData> list = i == 0 ? sample_topk_host(m, dc, K) : dev_cands[j];
            for (const TokData &t : list) {
                cands.push_back({j, dc.seek_delta, dc.has_ts, dc.seq});
                cands.back().seq.tokens.push_back(t);
                cands.back().seq.sum_logprobs_all += t.plog;
            }
        }
        beam_advance(s, cands, n_cur, i, n_prompt + i);
        for (int j = 0; j < n_cur; j++) {
            Decoder &dc = *s.dec[j];
            if (dc.completed || dc.failed) continue;
            const TokData &tk = dc.seq.tokens.back();
            if (tk.id > v.beg) {
                const int sd_new = 2 * (tk.id - v.beg);
                if (dc.has_ts && dc.seek_delta > sd_new && dc.seq.result_len < i) { dc.failed = true; continue; }
                dc.seek_delta = sd_new; dc.seq.result_len = i + 1; dc.has_ts = true;
            }
            if (tk.id == v.eot || (P.max_tokens > 0 && i >= P.max_tokens) || (dc.has_ts && seek + dc.seek_delta + 100 >= seek_end)) {
                if (dc.seq.result_len == 0) {
                    if (seek + dc.seek_delta + 100 >= seek_end) dc.seq.result_len = i + 1;
                    else { dc.failed = true; continue; }
                }
                if (P.single_segment) { dc.seq.result_len = i + 1; dc.seek_delta = 100 * kChunkSec; }
                dc.completed = true; continue;
            }
            if (i == n_max - 1 && (dc.seq.result_len == 0 || dc.seek_delta < 100 * kChunkSec / 2)) { dc.failed = true; continue; }
        }
        std::vector<int> live;
        for (int j = 0; j < n_cur; j++) if (!(s.dec[j]->completed || s.dec[j]->failed)) live.push_back(j);
        if (live.empty()) break;
        // ---- one batched step for the live beams (on the state's stream: it follows beam_advance's cache copies)
        const int n_past = n_prompt + i;
        bp.B = (int)live.size();
        for (int k = 0; k < bp.B; k++) {
            Decoder &dc = *s.dec[live[k]];
            DecCtl &c = *dc.h_ctl;
            memset(&c, 0, offsetof(DecCtl, prompt));
            const auto &tk = dc.seq.tokens;
            c.pos = n_past; c.pos0 = n_past; c.token = tk.back().id; c.n_prompt = 1; c.sample = 2;
            c.n_sampled = (int)tk.size(); c.last_id = tk.back().id; c.penult_id = tk.size() >= 2 ? tk[tk.size() - 2].id : -1;
            c.has_ts = dc.has_ts; c.seek_delta = dc.seek_delta; c.result_len = dc.seq.result_len;
            c.seek = seek; c.seek_end = seek_end; c.n_max = n_max;
            CUDA_CHECK(cudaMemcpyAsync(dc.mp.ctl, dc.h_ctl, offsetof(DecCtl, prompt), cudaMemcpyHostToDevice, s.stream));
            bp.seq[k] = BatchSeq{dc.mp.ctl, dc.mp.self_k, dc.mp.self_v, s.cross_k, s.cross_v, dc.mp.tok_out};
        }
        for (int k = bp.B; k < kMaxBatch; k++) bp.seq[k] = bp.seq[0];
        CUDA_CHECK(cudaMemsetAsync(bp.n_done, 0, sizeof(int), s.stream));
        decode_batch_step_graph(step_graph, bp, w, /*need_logits=*/true, decode_batch_xsplit(bp.B, bp.H, E.sms), s.stream, &s.n_launches, &bstep);
        CUDA_CHECK(cudaMemcpyAsync(E.beam_h_cand, E.beam_cand, (size_t)bp.B * 8 * sizeof(TokData), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        for (int k = 0; k < bp.B; k++) dev_cands[live[k]].assign(E.beam_h_cand + (size_t)k * 8, E.beam_h_cand + (size_t)k * 8 + K);
        s.n_decoded += bp.B;
    }
}

// ------------------------------------------------------------------------------------------------
// t > 0 fallback rungs on the batched step (default; SS_BATCH_SAMPLE=0 turns it off).
// whisper_full runs best_of (5) sampled decoders per rung: one batch-1 launch + a 207 KB logits read-back + a 51 866-wide
// filter / soft-max / discrete_distribution on the host PER DECODER AND TOKEN in the SS_BATCH_SAMPLE=0 path (engine.cc).  Here the
// decoders are the sequences of ONE batched step and the whole rung runs on the device like the temperature-0 loop: the prompt goes
// through the batch-1 kernel once (decoder 0, its self-KV rows copied to the others), then every step filters, divides by the
// temperature, draws (bd_sample_kernel, sample == 3) and applies whisper_full's per-token bookkeeping per sequence.  The uniforms
// come from each decoder's own std::mt19937 through std::generate_canonical<double, 53> - what std::discrete_distribution consumes -
// generated ahead on a copy of the generator; afterwards the real generator is advanced by the draws the decoder actually used.
// ------------------------------------------------------------------------------------------------
bool batch_sample_enabled() {
    const char *e = getenv("SS_BATCH_SAMPLE");
    return !(e && e[0] == '0');
}

static void decode_sampled_batched_locked(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek,
                                         int seek_end, int n_max, int tid0_init);
void decode_sampled_batched(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek, int seek_end,
                            int n_max, int tid0_init) {
    Engine &E = *s.engine;
    std::lock_guard<std::mutex> lk(E.batch_mu);      // the operand buffers are the engine's
    ensure_batch_resources(E);
    decode_sampled_batched_locked(s, P, t_cur, n_cur, prompt, seek, seek_end, n_max, tid0_init);
}
// (the caller holds the engine's batch mutex: transcribe_batch runs the ladder of a failed clip while it owns the operand buffers)
static void decode_sampled_batched_locked(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek,
                                         int seek_end, int n_max, int tid0_init) {
    Engine &E = *s.engine;
    const Model &m = E.model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    const int n_prompt = (int)prompt.size();
    for (int j = 0; j < n_cur; j++) { set_sampling(*s.dec[j], P, tid0_init); ensure_params(s, *s.dec[j]); }
    // prompt[0 .. n_prompt - 2] through the batch-1 kernel on decoder 0 (no logits needed), its self-KV rows to the other decoders;
    // the last prompt token is the first step of the batched loop (every sequence recomputes that one row itself)
    if (n_prompt > 1) {
        Decoder &d0 = *s.dec[0];
        DecCtl &c = *d0.h_ctl;
        memset(&c, 0, offsetof(DecCtl, prompt));
        c.pos = 0; c.pos0 = 0; c.token = prompt[0]; c.n_prompt = n_prompt; c.sample = 0; c.last_id = -1; c.penult_id = -1;   // n_prompt: no LM head before the end
        for (int i = 0; i < n_prompt - 1; i++) c.prompt[i] = prompt[i];
        c.prompt[n_prompt - 1] = prompt[n_prompt - 1];
        CUDA_CHECK(cudaMemcpyAsync(d0.mp.ctl, d0.h_ctl, offsetof(DecCtl, u), cudaMemcpyHostToDevice, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));      // the pinned block is re-armed for the batched loop below
        decode_mega_launch(d0.d_mp, d0.d_ll, d0.ll_bytes, n_prompt - 1, s.mega_grid, s.stream);
        s.n_launches += 1;
        for (int j = 1; j < n_cur; j++) kv_copy(s, d0, *s.dec[j], n_prompt - 1);
    }
    s.n_decoded += n_prompt - 1;
    BatchParams bp{};
    bp.B = n_cur; bp.d = hp.n_text_state; bp.H = hp.n_text_head; bp.L = hp.n_text_layer; bp.T = hp.n_audio_ctx; bp.ctx = hp.n_text_ctx; bp.n_vocab = hp.n_vocab;
    bp.s4 = powf((float)(bp.d / bp.H), -0.25f);
    bp.tok_emb = m.tok_emb; bp.d_pos = m.d_pos; bp.lnf_w = m.d_ln.w; bp.lnf_b = m.d_ln.b;
    bp.eot = v.eot; bp.sot = v.sot; bp.translate = v.translate; bp.transcribe = v.transcribe; bp.solm = v.solm; bp.prev = v.prev;
    bp.nosp = v.nosp; bp.not_ = v.not_; bp.beg = v.beg; bp.blank = v.blank;
    bp.suppress_blank = P.suppress_blank; bp.tdrz = P.tdrz_enable; bp.tid0_init = tid0_init;
    decode_batch_bind(bp, E.batch_scratch);
    for (int j = 0; j < n_cur; j++) {
        Decoder &dc = *s.dec[j];
        DecCtl &c = *dc.h_ctl;
        memset(&c, 0, offsetof(DecCtl, prompt));
        c.pos = n_prompt - 1; c.pos0 = n_prompt - 1; c.token = prompt[n_prompt - 1]; c.n_prompt = 1; c.sample = 3; c.temperature = t_cur;
        c.last_id = -1; c.penult_id = -1; c.seek = seek; c.seek_end = seek_end; c.n_max = n_max; c.seek_delta = 100 * kChunkSec;
        c.prompt[0] = prompt[n_prompt - 1];
        std::mt19937 ahead = dc.rng;      // a copy: the decoder's generator moves on by the draws it really consumed (below)
        for (int i = 0; i < n_max; i++) c.u[i] = std::generate_canonical<double, 53>(ahead);
        CUDA_CHECK(cudaMemcpyAsync(dc.mp.ctl, dc.h_ctl, sizeof(DecCtl), cudaMemcpyHostToDevice, s.stream));
        bp.seq[j] = BatchSeq{dc.mp.ctl, dc.mp.self_k, dc.mp.self_v, s.cross_k, s.cross_v, dc.mp.tok_out};
    }
    for (int k = n_cur; k < kMaxBatch; k++) bp.seq[k] = bp.seq[0];
    const MegaParams &w = s.dec[0]->mp;
    const int xsplit = decode_batch_xsplit(n_cur, bp.H, E.sms);
    CUDA_CHECK(cudaMemsetAsync(bp.n_done, 0, sizeof(int), s.stream));
    BatchStepGraph step_graph;
    E.batch_h_flags[0] = E.batch_h_flags[1] = 0;
    for (int t = 0; t < n_max; t++) {
        if (t % kPollEvery == 0) {
            const int G = t / kPollEvery;
            if (G >= 2) {
                CUDA_CHECK(cudaEventSynchronize(E.batch_ev[G & 1]));
                if (E.batch_h_flags[G & 1] >= n_cur) break;
            }
        }
        decode_batch_step_graph(step_graph, bp, w, /*need_logits=*/true, xsplit, s.stream, &s.n_launches);
        if (t % kPollEvery == kPollEvery - 1) {
            const int G = t / kPollEvery;
            CUDA_CHECK(cudaMemcpyAsync(&E.batch_h_flags[G & 1], bp.n_done, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CUDA_CHECK(cudaEventRecord(E.batch_ev[G & 1], s.stream));
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    for (int j = 0; j < n_cur; j++) {
        Decoder &dc = *s.dec[j];
        CUDA_CHECK(cudaMemcpyAsync(dc.h_ctl, dc.mp.ctl, offsetof(DecCtl, prompt), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        const int ns = dc.h_ctl->n_sampled;
        if (ns > 0) {
            CUDA_CHECK(cudaMemcpyAsync(dc.h_tok, dc.mp.tok_out, (size_t)ns * sizeof(TokData), cudaMemcpyDeviceToHost, s.stream));
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
        }
        dc.seq.tokens.assign(dc.h_tok, dc.h_tok + ns);
        for (int i = 0; i < ns; i++) dc.seq.sum_logprobs_all += dc.h_tok[i].plog;
        dc.seq.result_len = dc.h_ctl->result_len; dc.seek_delta = dc.h_ctl->seek_delta;
        dc.failed = dc.h_ctl->failed; dc.completed = dc.h_ctl->completed; dc.has_ts = dc.h_ctl->has_ts;
        dc.rng.discard(2ull * (unsigned long long)ns);      // generate_canonical<double, 53> takes two 32-bit words per draw
        s.n_decoded += ns;
    }
}

int transcribe_batch(State *const *states, const float *const *pcm, const size_t *n, int batch, const FullParams &P, bool stream_mode) {
    NvtxRange nvtx("ss.transcribe_batch");
    if (batch <= 0) return 0;
    Engine &E = *states[0]->engine;
    const Model &m = E.model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    bool batched = batch >= batch_decode_min() && P.beam_size <= 1 && !P.keep_logits && P.temperature < 1e-6f && hp.n_text_state == hp.n_text_head * 64 &&
                   hp.n_text_state <= 1280 && hp.n_text_ctx <= 512 && hp.n_audio_ctx <= 1536;
    for (int i = 0; i < batch && batched; i++)
        for (int j = 0; j < i; j++) if (states[i] == states[j]) batched = false;      // one state twice: the calls must serialise
    if (!batched) {
        int bad = -1;
        for (int i = 0; i < batch; i++) {      // clip by clip; a failing clip does not stop the others
            states[i]->last_status = 0; states[i]->last_error.clear();
            try { const int r = transcribe(*states[i], pcm[i], n[i], P, stream_mode); if (r) { states[i]->last_status = r; states[i]->last_error = "transcribe failed"; } }
            catch (const Error &e) { states[i]->last_status = e.code; states[i]->last_error = e.what(); }
            if (states[i]->last_status && bad < 0) bad = i;
        }
        if (bad >= 0) SS_THROW(states[bad]->last_status, "clip %d: %s (the other clips of the batch completed)", bad, states[bad]->last_error.c_str());
        return 0;
    }
    std::lock_guard<std::mutex> lk(E.batch_mu);
    CUDA_CHECK(cudaSetDevice(E.device));
    ensure_batch_resources(E);

    int lang = 0;
    if (v.multilingual) { lang = lang_id(P.language.c_str()); if (lang < 0) SS_THROW(-6, "unknown language '%s'", P.language.c_str()); }
    Common C;
    if (P.temperature_inc > 0.0f) { for (float t = P.temperature; t < 1.0f + 1e-6f; t += P.temperature_inc) C.temps.push_back(t); }
    else C.temps.push_back(P.temperature);
    C.n_decoders = std::max(1, P.best_of);
    if (P.best_of > kMaxDecoders) SS_THROW(-1, "best_of above %d", kMaxDecoders);
    C.prompt_init = {v.sot};
    if (v.multilingual) { C.prompt_init.push_back(v.sot + 1 + lang); C.prompt_init.push_back(v.transcribe); }
    C.n_max = hp.n_text_ctx / 2 - 4;
    const float precision = (float)kChunkSec / hp.n_audio_ctx;
    C.tid0_init = P.max_initial_ts > 0.0f ? (int)std::round(P.max_initial_ts / precision) : -1;

    std::vector<ClipRun> runs(batch);
    // A clip whose read-back fails (invalid UTF-8 in a segment: whisper.rs:85 fails that call) fails alone: its State records the
    // error, the other clips of the batch complete and keep their results (ss_state_status); the batch call reports the first error.
    auto finish_clip = [&](ClipRun &r) {
        try { postprocess(*r.s, stream_mode); }
        catch (const Error &e) { r.s->last_status = e.code; r.s->last_error = e.what(); }
        r.finished = true;
    };
    for (int i = 0; i < batch; i++) { states[i]->last_status = 0; states[i]->last_error.clear(); }
    for (int i = 0; i < batch; i++) {      // log-mel of every clip, each on its own stream
        State &s = *states[i];
        runs[i].s = &s;
        s.raw.clear(); s.out.clear(); s.full_text.clear(); s.result_tokens.clear();
        s.n_fallbacks = 0; s.n_decoded = 0; s.n_windows = 0; s.n_launches = 0; s.n_keep = 0; s.h_keep.clear();
        s.ms_mel = s.ms_enc = s.ms_dec = 0;
        CUDA_CHECK(cudaEventRecord(s.ev[0], s.stream));
        run_log_mel(s, pcm[i], n[i]);
        CUDA_CHECK(cudaEventRecord(s.ev[1], s.stream));
    }
    for (int i = 0; i < batch; i++) {
        State &s = *states[i];
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        { float ms; cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.ms_mel += ms; }
        runs[i].seek = 0; runs[i].seek_end = s.n_len_org;
        if (runs[i].seek_end < 100) { finish_clip(runs[i]); continue; }
        if (P.no_context) s.prompt_past.clear();
    }

    while (true) {
        std::vector<ClipRun *> act;
        for (auto &r : runs) {
            if (r.finished) continue;
            if (r.seek + 100 >= r.seek_end) { finish_clip(r); continue; }
            act.push_back(&r);
        }
        if (act.empty()) break;
        const bool enc_batched = batch_encoder_enabled() && act.size() >= 2;
        for (size_t g0 = 0; enc_batched && g0 < act.size(); g0 += kMaxBatch) {
            std::vector<ClipRun *> grp(act.begin() + g0, act.begin() + std::min(act.size(), g0 + (size_t)kMaxBatch));
            run_encode_batch(E, grp);
            CUDA_CHECK(cudaStreamSynchronize(E.batch_stream));      // the activations are reused by the next group
            float ms = 0.f; cudaEventElapsedTime(&ms, E.enc_ev[0], E.enc_ev[1]);
            for (ClipRun *r : grp) r->s->ms_enc += ms / (float)grp.size();
        }
        for (ClipRun *r : act) {      // encoder + cross-KV of the window, temperature-0 control block
            State &s = *r->s;
            CUDA_CHECK(cudaEventRecord(s.ev[0], s.stream));
            if (!enc_batched) run_encode(s, r->seek);
            CUDA_CHECK(cudaEventRecord(s.ev[1], s.stream));
            s.n_windows++;
            if (r->seek > 0 && r->seek + 500 >= r->seek_end) s.prompt_past.clear();
            r->best_decoder_id = 0;
            reset_decoders(s, 1);
            build_prompt(s, P, C, C.temps[0], r->prompt);
            Decoder &dc = *s.dec[0];
            DecCtl &c = *dc.h_ctl;
            memset(&c, 0, sizeof c);
            const int n_prompt = (int)r->prompt.size();
            c.pos = 0; c.pos0 = 0; c.token = r->prompt[0]; c.n_prompt = n_prompt; c.sample = 1; c.last_id = -1; c.penult_id = -1;
            c.seek = r->seek; c.seek_end = r->seek_end; c.n_max = C.n_max; c.seek_delta = 100 * kChunkSec;
            for (int i = 0; i < n_prompt; i++) c.prompt[i] = r->prompt[i];
            upload_ctl(s, dc);
            CUDA_CHECK(cudaEventRecord(s.ev[2], s.stream));
        }
        for (size_t g0 = 0; g0 < act.size(); g0 += kMaxBatch) {
            std::vector<ClipRun *> grp(act.begin() + g0, act.begin() + std::min(act.size(), g0 + (size_t)kMaxBatch));
            if (grp.size() == 1) {      // a lone straggler: the batch-1 persistent kernel is the faster one
                State &s = *grp[0]->s; Decoder &dc = *s.dec[0];
                set_sampling(dc, P, C.tid0_init);
                ensure_params(s, dc);
                CUDA_CHECK(cudaEventRecord(s.ev[2], s.stream));
                decode_mega_launch(dc.d_mp, dc.d_ll, dc.ll_bytes, (int)grp[0]->prompt.size() + C.n_max - 1, s.mega_grid, s.stream);
                s.n_launches += 1;
                CUDA_CHECK(cudaEventRecord(s.ev[3], s.stream));
                CUDA_CHECK(cudaMemcpyAsync(dc.h_ctl, dc.mp.ctl, offsetof(DecCtl, prompt), cudaMemcpyDeviceToHost, s.stream));
                CUDA_CHECK(cudaStreamSynchronize(s.stream));
                { float ms; cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]); s.ms_dec += ms; }
                const int ns = dc.h_ctl->n_sampled;
                if (ns > 0) {
                    CUDA_CHECK(cudaMemcpyAsync(dc.h_tok, dc.mp.tok_out, (size_t)ns * sizeof(TokData), cudaMemcpyDeviceToHost, s.stream));
                    CUDA_CHECK(cudaStreamSynchronize(s.stream));
                }
                dc.seq.tokens.assign(dc.h_tok, dc.h_tok + ns);
                for (int i = 0; i < ns; i++) dc.seq.sum_logprobs_all += dc.h_tok[i].plog;
                dc.seq.result_len = dc.h_ctl->result_len; dc.seek_delta = dc.h_ctl->seek_delta;
                dc.failed = dc.h_ctl->failed; dc.completed = dc.h_ctl->completed; dc.has_ts = dc.h_ctl->has_ts;
                s.n_decoded += (int)grp[0]->prompt.size() - 1 + ns;
            } else decode_group(E, P, C, grp);
        }
        for (ClipRun *r : act) {      // scoring, temperature ladder, segments, seek advance - per clip
            State &s = *r->s;
            if (!enc_batched) { float ms; cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.ms_enc += ms; }
            bool settled = score_and_test(s, P, C, 0, 1, r->best_decoder_id);
            for (size_t it = 1; it < C.temps.size() && !settled; it++) {
                const float t_cur = C.temps[it];
                const int n_cur = std::max(1, t_cur > 0.0f ? C.n_decoders : 1);
                s.n_fallbacks++;
                reset_decoders(s, n_cur);
                build_prompt(s, P, C, t_cur, r->prompt);
                CUDA_CHECK(cudaEventRecord(s.ev[2], s.stream));
                if (t_cur > 0.0f && batch_sample_enabled() && C.n_max <= kMaxDraws)      // (this function's batch path implies the supported shapes)
                    decode_sampled_batched_locked(s, P, t_cur, n_cur, r->prompt, r->seek, r->seek_end, C.n_max, C.tid0_init);
                else decode_sampled(s, P, C, t_cur, n_cur, r->prompt, r->seek, r->seek_end);
                CUDA_CHECK(cudaEventRecord(s.ev[3], s.stream));
                CUDA_CHECK(cudaStreamSynchronize(s.stream));
                { float ms; cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]); s.ms_dec += ms; }
                settled = score_and_test(s, P, C, it, n_cur, r->best_decoder_id);
            }
            r->seek += emit_window(s, P, C, *r);
        }
    }
    for (int i = 0; i < batch; i++)
        if (states[i]->last_status) SS_THROW(states[i]->last_status, "clip %d: %s (the other clips of the batch completed)", i, states[i]->last_error.c_str());
    return 0;
}

}  // namespace ss
