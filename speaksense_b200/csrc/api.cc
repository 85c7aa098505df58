// api.cc - the C ABI declared in include/speaksense_whisper.h.
#include <cstring>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/speaksense_whisper.h"
#include "engine.h"
#include "engine_internal.h"

using namespace ss;

// `e` is replicas[0]; an engine opened with ss_engine_open / ss_engine_open_dist has one replica (one device)
struct ss_engine {
    std::shared_ptr<Engine> e;
    std::vector<std::shared_ptr<Engine>> replicas;
    explicit ss_engine(std::shared_ptr<Engine> one) : e(one), replicas{one} {}
    explicit ss_engine(std::vector<std::shared_ptr<Engine>> all) : e(all.at(0)), replicas(std::move(all)) {}
    bool owns(const ss_state *s) const;
};
struct ss_state { State *s; };

static thread_local std::string g_err;

bool ss_engine::owns(const ss_state *s) const {
    for (const auto &r : replicas) if (s->s->engine.get() == r.get()) return true;
    return false;
}

template <typename F>
static int guard(F &&f) {
    try {
        return f();
    } catch (const ss::Error &e) {
        g_err = e.what(); return e.code;
    } catch (const std::bad_alloc &) {
        g_err = "host out of memory"; return SS_ERR_OOM;
    } catch (const std::exception &e) {
        g_err = e.what(); return SS_ERR_INTERNAL;
    }
}

static FullParams make_params(const ss_params *p) {
    FullParams fp;   // build_params defaults (whisper.rs:131-173)
    if (p) {
        if (p->language && p->language[0]) fp.language = p->language;      // whisper.rs:60-63
        fp.tdrz_enable = p->speaker_diarization != 0;                        // whisper.rs:137-140
        if (p->stream_mode) { fp.single_segment = false; fp.no_context = true; }   // whisper.rs:65-69 (audio_ctx 0 == full ctx)
        fp.beam_size = p->beam_size;
        fp.keep_logits = p->debug_keep_logits != 0;
    }
    return fp;
}

extern "C" {

void ss_params_default(ss_params *p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->language = nullptr; p->speaker_diarization = 0; p->stream_mode = 0; p->min_segment_length = 10;   // mod.rs:18-25
}
int ss_abi_version(void) { return SS_ABI_VERSION; }
const char *ss_last_error(void) { return g_err.c_str(); }
const char *ss_build_info(void) { return "speaksense_b200 whisper engine; sm_100a; tcgen05+TMA GEMM; CUDA " SS_STR(CUDART_VERSION); }

int ss_engine_open(const char *path, int device, ss_engine **out) {
    return guard([&]() -> int {
        if (!path || !out) SS_THROW(SS_ERR_INVALID, "null argument");
        auto e = engine_open(path, device);
        *out = new ss_engine{e};
        return 0;
    });
}
int ss_engine_open_multi(const char *path, const int *devices, int n_devices, ss_engine **out) {
    return guard([&]() -> int {
        if (!path || !out || !devices || n_devices <= 0) SS_THROW(SS_ERR_INVALID, "null argument");
        *out = new ss_engine{engine_open_multi(path, devices, n_devices)};
        return 0;
    });
}
int ss_engine_n_devices(const ss_engine *e) { return e ? (int)e->replicas.size() : 0; }
int ss_engine_device(const ss_engine *e, int i) { return (e && i >= 0 && i < (int)e->replicas.size()) ? e->replicas[i]->device : -1; }
int ss_engine_n_states(const ss_engine *e, int i) { return (e && i >= 0 && i < (int)e->replicas.size()) ? e->replicas[i]->n_states.load() : -1; }
int ss_engine_arena_fnv1a(const ss_engine *e, int i, uint64_t *out) {
    return guard([&]() -> int {
        if (!e || !out || i < 0 || i >= (int)e->replicas.size()) SS_THROW(SS_ERR_INVALID, "bad argument");
        const Model &m = e->replicas[i]->model;
        CUDA_CHECK(cudaSetDevice(m.device));
        const size_t blk = (size_t)64 << 20;
        std::vector<unsigned char> buf(std::min(blk, m.arena_bytes));
        uint64_t h = 1469598103934665603ull;
        for (size_t off = 0; off < m.arena_bytes; off += blk) {
            const size_t nb = std::min(blk, m.arena_bytes - off);
            CUDA_CHECK(cudaMemcpy(buf.data(), m.arena + off, nb, cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < nb; k++) { h ^= buf[k]; h *= 1099511628211ull; }
        }
        *out = h;
        return 0;
    });
}
int ss_nccl_unique_id(unsigned char out[128]) {
    return guard([&]() -> int { if (!out) SS_THROW(SS_ERR_INVALID, "null argument"); nccl_unique_id(out); return 0; });
}
int ss_engine_open_dist(const char *path, int device, int rank, int world, const unsigned char nccl_id[128], ss_engine **out) {
    return guard([&]() -> int {
        if (!out || (world > 1 && !nccl_id) || rank < 0 || rank >= (world > 0 ? world : 1)) SS_THROW(SS_ERR_INVALID, "bad argument");
        auto e = engine_open_dist(path, device, rank, world, nccl_id);
        *out = new ss_engine{e};
        return 0;
    });
}
void ss_engine_close(ss_engine *e) { delete e; }
int ss_engine_info(const ss_engine *e, int *n_vocab, int *n_audio_state, int *n_audio_layer, int *n_text_layer, int *n_mels, int64_t *weight_bytes) {
    if (!e) { g_err = "null engine"; return SS_ERR_INVALID; }
    const HParams &hp = e->e->model.hp;
    if (n_vocab) *n_vocab = hp.n_vocab;
    if (n_audio_state) *n_audio_state = hp.n_audio_state;
    if (n_audio_layer) *n_audio_layer = hp.n_audio_layer;
    if (n_text_layer) *n_text_layer = hp.n_text_layer;
    if (n_mels) *n_mels = hp.n_mels;
    if (weight_bytes) *weight_bytes = (int64_t)e->e->model.arena_bytes;
    return 0;
}

int ss_state_new(ss_engine *e, ss_state **out) {
    return guard([&]() -> int {
        if (!e || !out) SS_THROW(SS_ERR_INVALID, "null argument");
        return ss_state_new_on(e, e->replicas.size() > 1 ? -1 : e->e->device, out);
    });
}
int ss_state_new_on(ss_engine *e, int device, ss_state **out) {
    return guard([&]() -> int {
        if (!e || !out) SS_THROW(SS_ERR_INVALID, "null argument");
        std::shared_ptr<Engine> pick;
        if (device < 0) {      // least loaded replica (ties: the first in the ss_engine_open_multi order)
            for (const auto &r : e->replicas) if (!pick || r->n_states.load() < pick->n_states.load()) pick = r;
        } else {
            for (const auto &r : e->replicas) if (r->device == device) pick = r;
            if (!pick) SS_THROW(SS_ERR_INVALID, "engine has no replica on CUDA device %d", device);
        }
        *out = new ss_state{state_new(pick)};
        return 0;
    });
}
int ss_state_device(const ss_state *s) { return s ? s->s->engine->device : -1; }
int ss_state_status(const ss_state *s) { return s ? s->s->last_status : SS_ERR_INVALID; }
const char *ss_state_error(const ss_state *s) { return s ? s->s->last_error.c_str() : ""; }
void ss_state_free(ss_state *s) { if (s) { delete s->s; delete s; } }

int ss_transcribe(ss_engine *e, ss_state *s, const float *pcm, size_t n, const ss_params *p) {
    return guard([&]() -> int {
        if (!e || !s || (!pcm && n)) SS_THROW(SS_ERR_INVALID, "null argument");
        if (!e->owns(s)) SS_THROW(SS_ERR_INVALID, "state belongs to another engine");
        State &st = *s->s;
        st.last_status = 0; st.last_error.clear();
        try { return transcribe(st, pcm, n, make_params(p), p && p->stream_mode); }
        catch (const ss::Error &er) { st.last_status = er.code; st.last_error = er.what(); throw; }
    });
}
int ss_upload_pcm(ss_engine *e, ss_state *s, const float *pcm, size_t n) {
    return guard([&]() -> int {
        if (!e || !s || (!pcm && n)) SS_THROW(SS_ERR_INVALID, "null argument");
        upload_pcm(*s->s, pcm, n);
        CUDA_CHECK(cudaStreamSynchronize(s->s->stream));
        return 0;
    });
}
void ss_denoise_config_default(ss_denoise_config *c) {
    if (!c) return;
    c->frame_size = 2048; c->overlap = 0.75f; c->strength = 0.2f; c->noise_gate = 0.003f; c->enable_noise_reduction = 1; c->threshold = 0.002f;
}
int ss_denoise_audio(ss_engine *e, ss_state *s, const float *pcm, size_t n, const ss_denoise_config *cfg, float *out, int *noise_type, float *spectral_variance) {
    return guard([&]() -> int {
        if (!e || !s || !pcm) SS_THROW(SS_ERR_INVALID, "null argument");
        if (!e->owns(s)) SS_THROW(SS_ERR_INVALID, "state belongs to another engine");
        ss_denoise_config c; ss_denoise_config_default(&c);
        if (cfg) c = *cfg;
        const int t = denoise_audio(*s->s, pcm, n, c.frame_size, c.overlap, c.strength, out, spectral_variance);
        if (noise_type) *noise_type = t;
        return 0;
    });
}
int ss_denoise_frames(ss_engine *e, ss_state *s, const float *frames, int n_frames, const ss_denoise_config *cfg, float *out) {
    return guard([&]() -> int {
        if (!e || !s || !frames || !out || n_frames < 0) SS_THROW(SS_ERR_INVALID, "null argument");
        if (!e->owns(s)) SS_THROW(SS_ERR_INVALID, "state belongs to another engine");
        ss_denoise_config c; ss_denoise_config_default(&c);
        if (cfg) c = *cfg;
        if (!c.enable_noise_reduction) {      // mod.rs:131-133: only the noise gate applies
            for (size_t i = 0; i < (size_t)n_frames * c.frame_size; i++) out[i] = fabsf(frames[i]) < c.noise_gate ? 0.0f : frames[i];
            return 0;
        }
        denoise_frames(*s->s, frames, n_frames, c.frame_size, c.strength, c.noise_gate, out);
        return 0;
    });
}
int ss_transcribe_resident(ss_engine *e, ss_state *s, const ss_params *p) {
    return guard([&]() -> int {
        if (!e || !s) SS_THROW(SS_ERR_INVALID, "null argument");
        if (!s->s->d_pcm) SS_THROW(SS_ERR_INVALID, "no resident PCM: call ss_upload_pcm first");
        return transcribe(*s->s, nullptr, s->s->n_resident, make_params(p), p && p->stream_mode);
    });
}
int ss_bench_decode_steps(ss_engine *e, ss_state *s, int n_steps, int n_past0, float *ms_per_step) {
    return guard([&]() -> int {
        if (!e || !s || !ms_per_step) SS_THROW(SS_ERR_INVALID, "null argument");
        *ms_per_step = bench_decode_steps(*s->s, n_steps, n_past0);
        return 0;
    });
}
int ss_transcribe_batch(ss_engine *e, ss_state *const *states, const float *const *pcm, const size_t *n, int batch, const ss_params *p) {
    return guard([&]() -> int {
        if (!e || !states || !pcm || !n || batch < 0) SS_THROW(SS_ERR_INVALID, "null argument");
        std::vector<size_t> len(batch);
        for (int i = 0; i < batch; i++) {
            if (!states[i] || !e->owns(states[i])) SS_THROW(SS_ERR_INVALID, "bad state %d", i);
            len[i] = n[i];
            if (!pcm[i]) {      // NULL clip: the state's resident PCM (ss_upload_pcm / ss_denoise_audio), as ss_transcribe_resident
                if (!states[i]->s->d_pcm) SS_THROW(SS_ERR_INVALID, "clip %d: no resident PCM: call ss_upload_pcm / ss_denoise_audio first", i);
                len[i] = states[i]->s->n_resident;
            }
        }
        if (batch_decode_enabled()) {      // one batched decoder step per token for all clips (engine_batch.cc); SS_BATCH_DECODE=0: clip by clip
            // states of a multi-device engine are grouped by the device they live on; the groups run concurrently, one host thread each
            std::vector<std::vector<int>> groups;
            for (const auto &r : e->replicas) {
                std::vector<int> g;
                for (int i = 0; i < batch; i++) if (states[i]->s->engine.get() == r.get()) g.push_back(i);
                if (!g.empty()) groups.push_back(std::move(g));
            }
            const FullParams fp = make_params(p);
            const bool sm = p && p->stream_mode;
            auto run_group = [&](const std::vector<int> &g) -> int {
                std::vector<State *> st; std::vector<const float *> pc; std::vector<size_t> ln;
                for (int i : g) { st.push_back(states[i]->s); pc.push_back(pcm[i]); ln.push_back(len[i]); }
                return transcribe_batch(st.data(), pc.data(), ln.data(), (int)g.size(), fp, sm);
            };
            if (groups.size() <= 1) return groups.empty() ? 0 : run_group(groups[0]);
            std::vector<int> rcs(groups.size(), 0); std::vector<std::string> errs(groups.size());
            std::vector<std::thread> th;
            for (size_t k = 0; k < groups.size(); k++)
                th.emplace_back([&, k]() {
                    try { rcs[k] = run_group(groups[k]); }
                    catch (const ss::Error &er) { rcs[k] = er.code; errs[k] = er.what(); }
                    catch (const std::exception &er) { rcs[k] = SS_ERR_INTERNAL; errs[k] = er.what(); }
                });
            for (auto &t : th) t.join();
            for (size_t k = 0; k < groups.size(); k++) if (rcs[k]) { if (!errs[k].empty()) g_err = errs[k]; return rcs[k]; }
            return 0;
        }
        int rc = 0;
        for (int i = 0; i < batch; i++) {
            const int r = transcribe(*states[i]->s, pcm[i], len[i], make_params(p), p && p->stream_mode);
            if (r && !rc) rc = r;
        }
        return rc;
    });
}

int ss_n_segments_raw(const ss_state *s) { return s ? (int)s->s->raw.size() : 0; }
const char *ss_segment_text_raw(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->raw.size()) ? s->s->raw[i].text.c_str() : nullptr; }
int64_t ss_segment_t0_raw(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->raw.size()) ? s->s->raw[i].t0 : -1; }
int64_t ss_segment_t1_raw(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->raw.size()) ? s->s->raw[i].t1 : -1; }
int ss_segment_speaker_turn_next_raw(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->raw.size()) ? s->s->raw[i].speaker_turn_next : 0; }

int ss_n_segments(const ss_state *s) { return s ? (int)s->s->out.size() : 0; }
const char *ss_segment_text(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->out.size()) ? s->s->out[i].text.c_str() : nullptr; }
double ss_segment_start(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->out.size()) ? s->s->out[i].start : -1.0; }
double ss_segment_end(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->out.size()) ? s->s->out[i].end : -1.0; }
int ss_segment_speaker_id(const ss_state *s, int i) { return (s && i >= 0 && i < (int)s->s->out.size()) ? s->s->out[i].speaker_id : 0; }
const char *ss_full_text(const ss_state *s) { return s ? s->s->full_text.c_str() : nullptr; }

int ss_n_result_tokens(const ss_state *s) { return s ? (int)s->s->result_tokens.size() : 0; }
int ss_result_token(const ss_state *s, int i, float *p, float *plog) {
    if (!s || i < 0 || i >= (int)s->s->result_tokens.size()) return -1;
    const TokData &t = s->s->result_tokens[i];
    if (p) *p = t.p;
    if (plog) *plog = t.plog;
    return t.id;
}
int ss_n_fallbacks(const ss_state *s) { return s ? s->s->n_fallbacks : 0; }
int ss_n_decoded(const ss_state *s) { return s ? s->s->n_decoded : 0; }
int ss_n_windows(const ss_state *s) { return s ? s->s->n_windows : 0; }
int ss_n_kernel_launches(const ss_state *s) { return s ? s->s->n_launches : 0; }
int ss_stage_ms(const ss_state *s, float *mel_ms, float *enc_ms, float *dec_ms) {
    if (!s) return SS_ERR_INVALID;
    if (mel_ms) *mel_ms = s->s->ms_mel;
    if (enc_ms) *enc_ms = s->s->ms_enc;
    if (dec_ms) *dec_ms = s->s->ms_dec;
    return 0;
}
const float *ss_debug_logits(const ss_state *s, int step, int *n_vocab) {
    if (!s || step < 0 || step >= s->s->n_keep) return nullptr;
    const int nv = s->s->engine->model.hp.n_vocab;
    if (n_vocab) *n_vocab = nv;
    return s->s->h_keep.data() + (size_t)step * nv;
}

int ss_is_promotional_text(const char *utf8) { return utf8 ? (is_promotional_text(utf8) ? 1 : 0) : 0; }
int ss_add_punctuation(const char *utf8, char *out, size_t out_cap) {
    if (!utf8 || !out) { g_err = "null argument"; return SS_ERR_INVALID; }
    const std::string r = add_punctuation(utf8);
    if (r.size() + 1 > out_cap) { g_err = "output buffer too small"; return SS_ERR_INVALID; }
    memcpy(out, r.c_str(), r.size() + 1);
    return (int)r.size();
}
int ss_is_valid_utf8(const char *bytes, size_t n) { return bytes ? (is_valid_utf8(std::string(bytes, n)) ? 1 : 0) : 0; }
int ss_model_probe(const char *path, int hparams_out[11], int64_t *arena_bytes, uint64_t *arena_fnv1a, int *token_eot, int *token_beg,
                   int *n_vocab_strings) {
    return guard([&]() -> int {
        if (!path) SS_THROW(SS_ERR_INVALID, "null argument");
        ModelProbe pr = probe_model(path);
        if (hparams_out) memcpy(hparams_out, &pr.hp, sizeof(int) * 11);
        if (arena_bytes) *arena_bytes = (int64_t)pr.arena_bytes;
        if (arena_fnv1a) *arena_fnv1a = pr.fnv1a;
        if (token_eot) *token_eot = pr.eot;
        if (token_beg) *token_beg = pr.beg;
        if (n_vocab_strings) *n_vocab_strings = pr.n_vocab_strings;
        return 0;
    });
}

int ss_log_mel(ss_engine *e, ss_state *s, const float *pcm, size_t n, float *mel_out, size_t mel_cap, int *n_len, int *n_len_org) {
    return guard([&]() -> int {
        if (!e || !s || (!pcm && n)) SS_THROW(SS_ERR_INVALID, "null argument");
        State &st = *s->s;
        run_log_mel(st, pcm, n);
        CUDA_CHECK(cudaStreamSynchronize(st.stream));
        if (n_len) *n_len = st.n_len;
        if (n_len_org) *n_len_org = st.n_len_org;
        if (mel_out) {
            const size_t need = (size_t)st.engine->model.hp.n_mels * st.n_len;
            if (mel_cap < need) SS_THROW(SS_ERR_INVALID, "mel_out too small: need %zu floats", need);
            CUDA_CHECK(cudaMemcpy(mel_out, st.d_mel, need * sizeof(float), cudaMemcpyDeviceToHost));
        }
        return 0;
    });
}
int ss_encode(ss_engine *e, ss_state *s, int seek, float *enc_out, size_t enc_cap) {
    return guard([&]() -> int {
        if (!e || !s) SS_THROW(SS_ERR_INVALID, "null argument");
        State &st = *s->s;
        if (!st.d_mel) SS_THROW(SS_ERR_INVALID, "ss_encode needs a preceding ss_log_mel on this state");
        run_encode(st, seek);
        CUDA_CHECK(cudaStreamSynchronize(st.stream));
        if (enc_out) {
            const HParams &hp = st.engine->model.hp;
            const size_t need = (size_t)hp.n_audio_ctx * hp.n_audio_state;
            if (enc_cap < need) SS_THROW(SS_ERR_INVALID, "enc_out too small: need %zu floats", need);
            CUDA_CHECK(cudaMemcpy(enc_out, st.enc_out, need * sizeof(float), cudaMemcpyDeviceToHost));
        }
        return 0;
    });
}
int ss_decode(ss_engine *e, ss_state *s, const int *tokens, int n, int n_past, float *logits_out) {
    return guard([&]() -> int {
        if (!e || !s || !tokens) SS_THROW(SS_ERR_INVALID, "null argument");
        run_decode_forced(*s->s, tokens, n, n_past, logits_out);
        return 0;
    });
}

int ss_debug_process_logits(const char *path, const int *ids, int n_ids, int has_ts, int seek_delta, const float *raw, float temperature,
                            float *out) {
    return guard([&]() -> int {
        if (!path || !raw || !out || n_ids < 0 || (n_ids > 0 && !ids)) SS_THROW(SS_ERR_INVALID, "bad argument");
        Model m;
        load_model_meta(path, m);
        Decoder dc;
        for (int k = 0; k < n_ids; k++) { TokData t{}; t.id = ids[k]; dc.seq.tokens.push_back(t); }
        dc.has_ts = has_ts != 0; dc.seek_delta = seek_delta;
        const FullParams P;      // build_params' constants (whisper.rs:131-173)
        process_logits_host(m, P, dc, raw, temperature);
        memcpy(out, dc.logits.data(), (size_t)m.hp.n_vocab * sizeof(float));
        return 0;
    });
}

int ss_debug_sequence_score(const int *ids, const float *plogs, int n, int result_len, float length_penalty, double out[4]) {
    return guard([&]() -> int {
        if (!ids || !plogs || !out || n < 0 || result_len < 0 || result_len > n) SS_THROW(SS_ERR_INVALID, "bad argument");
        Sequence q;
        for (int k = 0; k < n; k++) { TokData t{}; t.id = ids[k]; t.plog = plogs[k]; q.tokens.push_back(t); }
        q.result_len = result_len;
        FullParams P; P.length_penalty = length_penalty;
        sequence_score(P, q);
        out[0] = q.sum_logprobs; out[1] = q.avg_logprobs; out[2] = q.entropy; out[3] = q.score;
        return 0;
    });
}

int ss_debug_beam_assign(const int *ids, const int *len, int max_len, const double *sums, const int *decoder_idx, int n_cands,
                         const int *live, int n_cur, int i, int *out) {
    return guard([&]() -> int {
        if (n_cands < 0 || n_cur < 0 || max_len < 0 || !out || (n_cands > 0 && (!ids || !len || !sums || !decoder_idx)) || (n_cur > 0 && !live))
            SS_THROW(SS_ERR_INVALID, "bad argument");
        std::vector<BeamCandidate> cands((size_t)n_cands);
        for (int c = 0; c < n_cands; c++) {
            if (len[c] < 0 || len[c] > max_len) SS_THROW(SS_ERR_INVALID, "candidate %d: length %d outside [0, %d]", c, len[c], max_len);
            cands[c].decoder_idx = decoder_idx[c];
            cands[c].seek_delta = c;      // carries the original index through the sort
            cands[c].has_ts = false;
            for (int a = 0; a < len[c]; a++) { TokData t{}; t.id = ids[(size_t)c * max_len + a]; cands[c].seq.tokens.push_back(t); }
            cands[c].seq.sum_logprobs_all = sums[c];
        }
        std::vector<char> lv((size_t)n_cur);
        for (int j = 0; j < n_cur; j++) lv[j] = live[j] != 0;
        const std::vector<int> pick = beam_pick(cands, lv, i);
        for (int j = 0; j < n_cur; j++) out[j] = pick[j] >= 0 ? cands[pick[j]].seek_delta : -1;
        return 0;
    });
}

int ss_debug_gemm(int device, const uint16_t *a, const uint16_t *b, float *d, int M, int N, int K, int b_mn_major) {
    return guard([&]() -> int {
        if (!a || !b || !d || M <= 0 || N <= 0 || K <= 0) SS_THROW(SS_ERR_INVALID, "bad argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) SS_THROW(SS_ERR_NO_DEVICE, "no CUDA device available (no CPU fallback)");
        CUDA_CHECK(cudaSetDevice(device));
        gemm_init();
        __half *da, *db; float *dd;
        const size_t nb = b_mn_major ? (size_t)K * N : (size_t)N * K;
        CUDA_CHECK(cudaMalloc(&da, (size_t)M * K * 2)); CUDA_CHECK(cudaMalloc(&db, nb * 2)); CUDA_CHECK(cudaMalloc(&dd, (size_t)M * N * 4));
        CUDA_CHECK(cudaMemcpy(da, a, (size_t)M * K * 2, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(db, b, nb * 2, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemset(dd, 0xff, (size_t)M * N * 4));
        GemmOperand A; A.ptr = da; A.rows = M; A.ld = K;
        GemmOperand B; B.ptr = db;
        if (b_mn_major) { B.rows = K; B.ld = N; } else { B.rows = N; B.ld = K; }
        GemmEpilogue ep; ep.out = dd; ep.out_type = GEMM_OUT_F32; ep.out_ld = N;
        int launches = 0;
        gemm_enqueue(A, B, M, N, K, b_mn_major != 0, ep, 0, &launches);
        CUDA_CHECK(cudaDeviceSynchronize());
        CUDA_CHECK(cudaMemcpy(d, dd, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
        cudaFree(da); cudaFree(db); cudaFree(dd);
        return 0;
    });
}

}  // extern "C"
