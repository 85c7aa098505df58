// engine.cc - sessions, encoder orchestration, the whisper_full decode loop and post-processing.
//
// Host-side restatement of `state.full(params, &audio)` (/root/reference/src/asr/whisper.rs:75) for
// the B200 engine: whisper.cpp's whisper_full_with_state control flow with
// every arithmetic stage on the device.  The temperature-0 greedy decoder runs entirely on the GPU
// (CUDA-graph replay, on-device logits filter / argmax / state update); the t>0 fallback decoders are
// sampled on the host with std::mt19937 + std::discrete_distribution exactly like whisper.cpp.
#include "engine_internal.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

namespace ss {

namespace {

template <typename T>
T *dmalloc(size_t n) {
    T *p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) SS_THROW(-5, "cudaMalloc of %zu bytes failed", n * sizeof(T));
    return p;
}
template <typename T>
T *hmalloc(size_t n) {
    T *p = nullptr;
    if (cudaMallocHost(&p, n * sizeof(T)) != cudaSuccess) SS_THROW(-5, "cudaMallocHost of %zu bytes failed", n * sizeof(T));
    return p;
}

void check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) SS_THROW(-3, "no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= n) SS_THROW(-3, "CUDA device %d out of range (%d visible)", device, n);
    cudaDeviceProp pr;
    CUDA_CHECK(cudaGetDeviceProperties(&pr, device));
    if (pr.major != 10) SS_THROW(-3, "device %d is sm_%d%d; this engine is built for sm_100a (B200) only", device, pr.major, pr.minor);
    CUDA_CHECK(cudaSetDevice(device));
}

std::shared_ptr<Engine> finish_open(unsigned char *d_arena, size_t bytes, int device, const std::string &path) {
    auto e = std::make_shared<Engine>();
    e->device = device; e->path = path;
    bind_model(e->model, d_arena, bytes, device);
    gemm_init();
    attention_init();
    return e;
}

}  // namespace

std::shared_ptr<Engine> engine_open(const std::string &path, int device) {
    check_device(device);
    std::vector<unsigned char> img = build_arena_image(path);
    unsigned char *d = dmalloc<unsigned char>(img.size());
    CUDA_CHECK(cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice));
    return finish_open(d, img.size(), device, path);
}

// ------------------------------------------------------------------------------------------------
// NCCL weight broadcast: the only collective of the path, at init.
// libnccl is dlopen'ed so that the library loads on hosts without NCCL / without a GPU.
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, unsigned char[128] /* by value in the real ABI */, int) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
struct NcclId { char internal[128]; };
typedef int (*CommInitRankFn)(void **, int, NcclId, int);

NcclApi &nccl() {
    static NcclApi api;
    if (api.h) return api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
    if (!api.h) SS_THROW(-8, "cannot dlopen libnccl.so.2: %s", dlerror());
    api.GetUniqueId = reinterpret_cast<int (*)(void *)>(dlsym(api.h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.h, "ncclCommInitRank"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.h, "ncclBroadcast"));
    api.CommDestroy = reinterpret_cast<int (*)(void *)>(dlsym(api.h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<const char *(*)(int)>(dlsym(api.h, "ncclGetErrorString"));
    api.CommInitAll = reinterpret_cast<int (*)(void **, int, const int *)>(dlsym(api.h, "ncclCommInitAll"));
    api.GroupStart = reinterpret_cast<int (*)()>(dlsym(api.h, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(api.h, "ncclGroupEnd"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.Broadcast || !api.CommDestroy) SS_THROW(-8, "libnccl lacks expected symbols");
    return api;
}
#define NCCL_CHECK(expr)                                                                                   \
    do {                                                                                                   \
        int _r = (expr);                                                                                   \
        if (_r != 0) SS_THROW(-8, "NCCL error %d (%s) at %s:%d", _r, nccl().GetErrorString ? nccl().GetErrorString(_r) : "?", __FILE__, __LINE__); \
    } while (0)
}  // namespace

void nccl_unique_id(unsigned char out[128]) {
    NcclId id;
    NCCL_CHECK(nccl().GetUniqueId(&id));
    memcpy(out, id.internal, 128);
}

std::shared_ptr<Engine> engine_open_dist(const char *path, int device, int rank, int world, const unsigned char *nccl_id) {
    check_device(device);
    if (world <= 1) { if (!path) SS_THROW(-1, "model path required"); return engine_open(path, device); }
    NcclApi &api = nccl();
    NcclId id; memcpy(id.internal, nccl_id, 128);
    // rank 0 parses the file BEFORE the rendezvous; a failure there is broadcast as size 0 so that the other ranks return an
    // error instead of waiting in ncclBroadcast for ever
    std::vector<unsigned char> img;
    unsigned long long sz = 0;
    std::string load_error;
    if (rank == 0) {
        try {
            if (!path) SS_THROW(-1, "rank 0 needs the model path");
            img = build_arena_image(path); sz = img.size();
        } catch (const Error &e) { load_error = e.what(); sz = 0; }
    }
    struct Res {      // released on every path
        NcclApi &api; void *comm = nullptr; cudaStream_t st = nullptr; unsigned long long *d_sz = nullptr; unsigned char *d = nullptr;
        ~Res() { if (comm) api.CommDestroy(comm); if (d_sz) cudaFree(d_sz); if (st) cudaStreamDestroy(st); if (d) cudaFree(d); }
    } r{api};
    NCCL_CHECK(reinterpret_cast<CommInitRankFn>(api.CommInitRank)(&r.comm, world, id, rank));
    CUDA_CHECK(cudaStreamCreate(&r.st));
    r.d_sz = dmalloc<unsigned long long>(1);
    if (rank == 0) CUDA_CHECK(cudaMemcpy(r.d_sz, &sz, 8, cudaMemcpyHostToDevice));
    NCCL_CHECK(api.Broadcast(r.d_sz, r.d_sz, 8, /*ncclUint8*/ 1, 0, r.comm, r.st));
    CUDA_CHECK(cudaStreamSynchronize(r.st));
    CUDA_CHECK(cudaMemcpy(&sz, r.d_sz, 8, cudaMemcpyDeviceToHost));
    if (sz == 0) SS_THROW(-2, "rank 0 could not load the model%s%s", load_error.empty() ? "" : ": ", load_error.c_str());
    r.d = dmalloc<unsigned char>(sz);
    if (rank == 0) CUDA_CHECK(cudaMemcpy(r.d, img.data(), sz, cudaMemcpyHostToDevice));
    NCCL_CHECK(api.Broadcast(r.d, r.d, sz, 1, 0, r.comm, r.st));
    CUDA_CHECK(cudaStreamSynchronize(r.st));
    unsigned char *arena = r.d; r.d = nullptr;      // ownership moves to the engine
    return finish_open(arena, sz, device, path ? path : "<nccl broadcast>");
}

std::vector<std::shared_ptr<Engine>> engine_open_multi(const std::string &path, const int *devices, int n) {
    if (!devices || n <= 0) SS_THROW(-1, "engine_open_multi: no devices");
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) if (devices[i] == devices[j]) SS_THROW(-1, "engine_open_multi: device %d listed twice", devices[i]);
    std::vector<std::shared_ptr<Engine>> out;
    if (n == 1) { out.push_back(engine_open(path, devices[0])); return out; }
    for (int i = 0; i < n; i++) check_device(devices[i]);
    const std::vector<unsigned char> img = build_arena_image(path);
    const size_t sz = img.size();
    NcclApi &api = nccl();
    if (!api.CommInitAll || !api.GroupStart || !api.GroupEnd) SS_THROW(-8, "libnccl lacks ncclCommInitAll / ncclGroupStart / ncclGroupEnd");
    struct Res {
        NcclApi &api; std::vector<void *> comm; std::vector<cudaStream_t> st; std::vector<unsigned char *> d; std::vector<int> dev;
        ~Res() {
            for (size_t i = 0; i < dev.size(); i++) {
                cudaSetDevice(dev[i]);
                if (i < comm.size() && comm[i]) api.CommDestroy(comm[i]);
                if (i < st.size() && st[i]) cudaStreamDestroy(st[i]);
                if (i < d.size() && d[i]) cudaFree(d[i]);
            }
        }
    } r{api};
    r.dev.assign(devices, devices + n); r.comm.assign(n, nullptr); r.st.assign(n, nullptr); r.d.assign(n, nullptr);
    for (int i = 0; i < n; i++) {
        CUDA_CHECK(cudaSetDevice(devices[i]));
        r.d[i] = dmalloc<unsigned char>(sz);
        CUDA_CHECK(cudaStreamCreate(&r.st[i]));
    }
    CUDA_CHECK(cudaSetDevice(devices[0]));
    CUDA_CHECK(cudaMemcpy(r.d[0], img.data(), sz, cudaMemcpyHostToDevice));
    NCCL_CHECK(api.CommInitAll(r.comm.data(), n, devices));
    NCCL_CHECK(api.GroupStart());
    for (int i = 0; i < n; i++) {
        CUDA_CHECK(cudaSetDevice(devices[i]));
        NCCL_CHECK(api.Broadcast(r.d[i], r.d[i], sz, /*ncclUint8*/ 1, 0, r.comm[i], r.st[i]));
    }
    NCCL_CHECK(api.GroupEnd());
    for (int i = 0; i < n; i++) { CUDA_CHECK(cudaSetDevice(devices[i])); CUDA_CHECK(cudaStreamSynchronize(r.st[i])); }
    for (int i = 0; i < n; i++) {
        CUDA_CHECK(cudaSetDevice(devices[i]));
        unsigned char *arena = r.d[i]; r.d[i] = nullptr;
        out.push_back(finish_open(arena, sz, devices[i], path));
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
Decoder *new_decoder(State &s, bool with_keep) {
    const Model &m = s.engine->model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    if (hp.n_text_layer > kMaxLayers || hp.n_text_state > 1280) SS_THROW(-2, "decoder larger than large-v3 is not supported");
    auto d = std::make_unique<Decoder>();
    const size_t dd = hp.n_text_state, kv = (size_t)hp.n_text_layer * hp.n_text_ctx * dd;
    MegaParams &b = d->mp;
    b.d = hp.n_text_state; b.H = hp.n_text_head; b.L = hp.n_text_layer; b.T = hp.n_audio_ctx; b.ctx = hp.n_text_ctx; b.n_vocab = hp.n_vocab;
    b.xsplit = std::max(1, std::min(8, s.mega_grid / b.H));
    b.s4 = powf((float)(b.d / b.H), -0.25f);
    if (ceil_div(hp.n_vocab, s.mega_grid) > 512 || ceil_div(b.T, b.xsplit) > 512 || b.ctx > 500 || ceil_div(4 * b.d, s.mega_grid) > 250 ||
        b.H > s.mega_grid || s.mega_grid * 8 > 1280 || (b.d & 127))
        SS_THROW(-3, "device has too few / too many SMs (%d) for the decode kernel's per-CTA work buffers", s.mega_grid);
    b.tok_emb = m.tok_emb; b.d_pos = m.d_pos; b.lnf_w = m.d_ln.w; b.lnf_b = m.d_ln.b;
    for (int i = 0; i < hp.n_text_layer; i++) {
        const DecLayer &L = m.dec[i]; MegaLayer &o = b.layer[i];
        const Lin *lins[6] = {&L.qkv, &L.o, &L.cq, &L.co, &L.fc1, &L.fc2};
        for (int k = 0; k < 6; k++) { o.w[k] = lins[k]->w; o.b[k] = lins[k]->b; }
        const LNp *lns[3] = {&L.attn_ln, &L.cross_ln, &L.mlp_ln};
        for (int k = 0; k < 3; k++) { o.lnw[k] = lns[k]->w; o.lnb[k] = lns[k]->b; }
    }
    b.ctl = dmalloc<DecCtl>(1);
    {   // flagged exchange arena
        const size_t words = 9 * dd + 4 * dd + (size_t)hp.n_text_head * 8 * 66 + (size_t)s.mega_grid * 8;
        d->ll_bytes = words * sizeof(ss_u64);
        ss_u64 *p = dmalloc<ss_u64>(words);
        d->d_ll = p;
        b.xA = p; p += dd; b.xB = p; p += dd; b.xC = p; p += dd; b.q1 = p; p += dd; b.kcur = p; p += dd; b.vcur = p; p += dd;
        b.att1 = p; p += dd; b.q2 = p; p += dd; b.att2 = p; p += dd; b.hbuf = p; p += 4 * dd;
        b.part = p; p += (size_t)hp.n_text_head * 8 * 66; b.stats = p;
    }
    b.logits = dmalloc<float>(hp.n_vocab);
    b.tok_out = dmalloc<TokData>(hp.n_text_ctx);
    b.self_k = dmalloc<__half>(kv); b.self_v = dmalloc<__half>(kv);
    CUDA_CHECK(cudaMemset(b.self_k, 0, kv * 2)); CUDA_CHECK(cudaMemset(b.self_v, 0, kv * 2));
    b.cross_k = s.cross_k; b.cross_v = s.cross_v;
    b.keep = with_keep ? s.keep : nullptr; b.keep_cap = with_keep ? s.keep_cap : 0;
    b.prof = getenv("SS_MEGA_TRACE") ? dmalloc<long long>((size_t)s.mega_grid * 1024) : getenv("SS_MEGA_PROF") ? dmalloc<long long>((size_t)s.mega_grid * 96) : nullptr;
    if (b.prof && getenv("SS_MEGA_TRACE")) CUDA_CHECK(cudaMemset(b.prof, 0, (size_t)s.mega_grid * 1024 * 8));
    b.eot = v.eot; b.sot = v.sot; b.translate = v.translate; b.transcribe = v.transcribe; b.solm = v.solm; b.prev = v.prev;
    b.nosp = v.nosp; b.not_ = v.not_; b.beg = v.beg; b.blank = v.blank;
    b.suppress_blank = 1; b.tdrz = 0; b.tid0_init = -1;
    d->d_mp = dmalloc<MegaParams>(1); d->mp_dirty = true;
    d->h_ctl = hmalloc<DecCtl>(1); d->h_tok = hmalloc<TokData>(hp.n_text_ctx);
    memset(d->h_ctl, 0, sizeof(DecCtl));
    Decoder *raw = d.get();
    s.dec.push_back(std::move(d));
    return raw;
}

State *state_new(const std::shared_ptr<Engine> &e) {
    CUDA_CHECK(cudaSetDevice(e->device));
    auto s = std::make_unique<State>();
    s->engine = e;
    e->n_states.fetch_add(1);
    const HParams &hp = e->model.hp;
    const size_t T = hp.n_audio_ctx, d = hp.n_audio_state, dd = hp.n_text_state;
    CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto &ev : s->ev) CUDA_CHECK(cudaEventCreate(&ev));
    s->d_max = dmalloc<int>(1);
    s->win = dmalloc<__half>((2 * T + 2) * hp.n_mels);
    s->x1 = dmalloc<__half>((2 * T + 2) * d);
    CUDA_CHECK(cudaMemset(s->x1, 0, (2 * T + 2) * d * 2));
    s->x = dmalloc<float>(T * d); s->xn = dmalloc<__half>(T * d); s->qkv = dmalloc<__half>(T * 3 * d);
    s->att = dmalloc<__half>(T * d); s->ff = dmalloc<__half>(T * 4 * d);
    s->enc_out = dmalloc<float>(T * d); s->enc16 = dmalloc<__half>(T * d);
    s->cross_k = dmalloc<__half>((size_t)hp.n_text_layer * 2 * T * dd);   // [layer][K | V][head][T][64]
    s->cross_v = s->cross_k + T * dd;
    s->h_logits = hmalloc<float>(hp.n_vocab);
    decode_mega_configure();
    s->mega_grid = decode_mega_grid(e->device);
    new_decoder(*s, true);
    return s.release();
}

State::~State() {
    if (!engine) return;
    engine->n_states.fetch_sub(1);
    cudaSetDevice(engine->device);
    if (stream) cudaStreamSynchronize(stream);
    if (enc_graph) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(enc_graph));
    for (auto &d : dec) {
        MegaParams &b = d->mp;
        cudaFree(b.ctl); cudaFree(d->d_ll); cudaFree(b.logits); cudaFree(b.tok_out); if (b.prof) cudaFree(b.prof);
        cudaFree(b.self_k); cudaFree(b.self_v); cudaFree(d->d_mp);
        if (d->alt_k) cudaFree(d->alt_k);
        if (d->alt_v) cudaFree(d->alt_v);
        cudaFreeHost(d->h_ctl); cudaFreeHost(d->h_tok);
    }
    void *ptrs[] = {d_pcm, d_mel, d_max, win, x1, xn, qkv, att, ff, enc16, x, enc_out, cross_k, keep};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h_pcm) cudaFreeHost(h_pcm);
    if (h_logits) cudaFreeHost(h_logits);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (d_pcm_alt) cudaFree(d_pcm_alt);
    if (d_dn) cudaFree(d_dn);
    if (stream) cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------
void upload_pcm(State &s, const float *pcm, size_t n) {
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    if (n + 1 > s.pcm_cap) {
        // (pointers and capacities are reset BEFORE the throwing allocation: a failed cudaMalloc must not leave dangling pointers)
        if (s.d_pcm_alt) { cudaFree(s.d_pcm_alt); s.d_pcm_alt = nullptr; }
        if (s.d_pcm) { cudaFree(s.d_pcm); s.d_pcm = nullptr; }
        s.pcm_cap = 0; s.n_resident = 0;
        const size_t want_cap = std::max<size_t>(n + 1, (size_t)kSampleRate * kChunkSec);
        s.d_pcm = dmalloc<float>(want_cap);
        s.pcm_cap = want_cap;
    }
    if (n + 1 > s.h_pcm_cap) {
        if (s.h_pcm) { cudaFreeHost(s.h_pcm); s.h_pcm = nullptr; }
        s.h_pcm_cap = 0;
        const size_t want_h = std::max<size_t>(n + 1, (size_t)kSampleRate * kChunkSec);
        s.h_pcm = hmalloc<float>(want_h);
        s.h_pcm_cap = want_h;
    }
    if (n) memcpy(s.h_pcm, pcm, n * sizeof(float));      // caller memory is pageable: stage through pinned
    if (n) CUDA_CHECK(cudaMemcpyAsync(s.d_pcm, s.h_pcm, n * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    s.n_resident = n;
}

int denoise_audio(State &s, const float *pcm, size_t n, int frame_size, float overlap, float strength, float *out, float *nv_out) {
    NvtxRange nvtx("ss.denoise");
    upload_pcm(s, pcm, n);
    const size_t need = denoise_scratch_floats(n, frame_size, overlap);
    if (need > s.dn_cap || !s.d_pcm_alt) {
        if (s.d_dn) { cudaFree(s.d_dn); s.d_dn = nullptr; }
        if (s.d_pcm_alt) { cudaFree(s.d_pcm_alt); s.d_pcm_alt = nullptr; }
        s.dn_cap = 0;
        s.d_dn = dmalloc<float>(need); s.dn_cap = need; s.d_pcm_alt = dmalloc<float>(s.pcm_cap);
    }
    const int type = denoise_enqueue(s.d_pcm, n, frame_size, overlap, strength, s.d_pcm_alt, s.d_dn, s.stream, &s.n_launches, nv_out);
    std::swap(s.d_pcm, s.d_pcm_alt);      // the denoised chunk is now the resident PCM (ss_transcribe_resident)
    if (out) {
        CUDA_CHECK(cudaMemcpyAsync(s.h_pcm, s.d_pcm, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        memcpy(out, s.h_pcm, n * sizeof(float));
    } else CUDA_CHECK(cudaStreamSynchronize(s.stream));
    return type;
}

void denoise_frames(State &s, const float *frames, int n_frames, int frame_size, float strength, float noise_gate, float *out) {
    const size_t n = (size_t)n_frames * frame_size;
    upload_pcm(s, frames, n);
    if (!s.d_pcm_alt) s.d_pcm_alt = dmalloc<float>(s.pcm_cap);
    denoise_frames_enqueue(s.d_pcm, n_frames, frame_size, strength, noise_gate, s.d_pcm_alt, s.stream, &s.n_launches);
    CUDA_CHECK(cudaMemcpyAsync(s.h_pcm, s.d_pcm_alt, n * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    memcpy(out, s.h_pcm, n * sizeof(float));
}

void run_log_mel(State &s, const float *pcm, size_t n) {
    NvtxRange nvtx("ss.log_mel");
    const Model &m = s.engine->model;
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    if (pcm == nullptr && n == s.n_resident && s.d_pcm) {   // PCM already resident in HBM (ss_upload_pcm)
        s.n_len = mel_n_len(n); s.n_len_org = mel_n_len_org(n);
        const size_t need0 = (size_t)m.hp.n_mels * s.n_len;
        if (need0 > s.mel_cap) { if (s.d_mel) { cudaFree(s.d_mel); s.d_mel = nullptr; } s.mel_cap = 0; s.d_mel = dmalloc<float>(need0); s.mel_cap = need0; }
        mel_enqueue(m, s.d_pcm, n, s.d_mel, s.n_len, s.d_max, s.stream, &s.n_launches);
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (n + 1 > s.pcm_cap) {
        // (pointers and capacities are reset BEFORE the throwing allocation: a failed cudaMalloc must not leave dangling pointers)
        if (s.d_pcm_alt) { cudaFree(s.d_pcm_alt); s.d_pcm_alt = nullptr; }
        if (s.d_pcm) { cudaFree(s.d_pcm); s.d_pcm = nullptr; }
        s.pcm_cap = 0; s.n_resident = 0;
        const size_t want_cap = std::max<size_t>(n + 1, (size_t)kSampleRate * kChunkSec);
        s.d_pcm = dmalloc<float>(want_cap);
        s.pcm_cap = want_cap;
    }
    if (n + 1 > s.h_pcm_cap) {
        if (s.h_pcm) { cudaFreeHost(s.h_pcm); s.h_pcm = nullptr; }
        s.h_pcm_cap = 0;
        const size_t want_h = std::max<size_t>(n + 1, (size_t)kSampleRate * kChunkSec);
        s.h_pcm = hmalloc<float>(want_h);
        s.h_pcm_cap = want_h;
    }
    s.n_len = mel_n_len(n); s.n_len_org = mel_n_len_org(n);
    const size_t need = (size_t)m.hp.n_mels * s.n_len;
    if (need > s.mel_cap) { if (s.d_mel) { cudaFree(s.d_mel); s.d_mel = nullptr; } s.mel_cap = 0; s.d_mel = dmalloc<float>(need); s.mel_cap = need; }
    if (n) memcpy(s.h_pcm, pcm, n * sizeof(float));      // caller memory is pageable: stage through pinned
    if (n) CUDA_CHECK(cudaMemcpyAsync(s.d_pcm, s.h_pcm, n * sizeof(float), cudaMemcpyHostToDevice, s.stream));
    s.n_resident = n;
    mel_enqueue(m, s.d_pcm, n, s.d_mel, s.n_len, s.d_max, s.stream, &s.n_launches);
    CUDA_CHECK(cudaGetLastError());
}

// conv stem, encoder layers, ln_post and the cross-KV projection of the window staged in s.win: every launch works on buffers the
// State owns for its whole life, so the sequence can be captured once and replayed
static void encode_window_enqueue(State &s, cudaStream_t st, int *nl) {
    const Model &m = s.engine->model; const HParams &hp = m.hp;
    const int T = hp.n_audio_ctx, d = hp.n_audio_state, H = hp.n_audio_head, C = hp.n_mels;
    {   // conv1 (k3,s1,p1) + bias + GELU as implicit GEMM over overlapping rows of the padded window
        GemmOperand A; A.ptr = s.win; A.rows = 2 * T; A.ld = C;
        GemmOperand B; B.ptr = m.conv1.w; B.rows = d; B.ld = 3 * C;
        GemmEpilogue ep; ep.bias = m.conv1.b; ep.gelu = 1; ep.out = s.x1; ep.out_type = GEMM_OUT_F16; ep.out_ld = d; ep.out_row_offset = 1;
        gemm_enqueue(A, B, 2 * T, d, 3 * C, false, ep, st, nl);
    }
    {   // conv2 (k3,s2,p1) + bias + GELU + positional embedding -> residual stream (f32)
        GemmOperand A; A.ptr = s.x1; A.rows = T; A.ld = 2 * d;
        GemmOperand B; B.ptr = m.conv2.w; B.rows = d; B.ld = 3 * d;
        GemmEpilogue ep; ep.bias = m.conv2.b; ep.gelu = 1; ep.pos = m.e_pos; ep.pos_rows = T; ep.out = s.x; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
        gemm_enqueue(A, B, T, d, 3 * d, false, ep, st, nl);
    }
    for (int il = 0; il < hp.n_audio_layer; il++) {
        const EncLayer &L = m.enc[il];
        layernorm_f16_enqueue(s.x, s.xn, T, d, L.attn_ln, st, nl);
        {
            GemmOperand A; A.ptr = s.xn; A.rows = T; A.ld = d;
            GemmOperand B; B.ptr = L.qkv.w; B.rows = 3 * d; B.ld = d;
            GemmEpilogue ep; ep.bias = L.qkv.b; ep.out = s.qkv; ep.out_ld = 3 * d;
            gemm_enqueue(A, B, T, 3 * d, d, false, ep, st, nl);
        }
        attention_enqueue(s.qkv, s.att, 1, T, H, 1.0f / sqrtf(64.0f), st, nl);   // fused: S and P never leave the SM
        {
            GemmOperand A; A.ptr = s.att; A.rows = T; A.ld = d;
            GemmOperand B; B.ptr = L.o.w; B.rows = d; B.ld = d;
            GemmEpilogue ep; ep.bias = L.o.b; ep.residual = 1; ep.out = s.x; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
            gemm_enqueue(A, B, T, d, d, false, ep, st, nl);
        }
        layernorm_f16_enqueue(s.x, s.xn, T, d, L.mlp_ln, st, nl);
        {
            GemmOperand A; A.ptr = s.xn; A.rows = T; A.ld = d;
            GemmOperand B; B.ptr = L.fc1.w; B.rows = 4 * d; B.ld = d;
            GemmEpilogue ep; ep.bias = L.fc1.b; ep.gelu = 1; ep.out = s.ff; ep.out_ld = 4 * d;
            gemm_enqueue(A, B, T, 4 * d, d, false, ep, st, nl);
        }
        {
            GemmOperand A; A.ptr = s.ff; A.rows = T; A.ld = 4 * d;
            GemmOperand B; B.ptr = L.fc2.w; B.rows = d; B.ld = 4 * d;
            GemmEpilogue ep; ep.bias = L.fc2.b; ep.residual = 1; ep.out = s.x; ep.out_type = GEMM_OUT_F32; ep.out_ld = d;
            gemm_enqueue(A, B, T, d, 4 * d, false, ep, st, nl);
        }
    }
    layernorm_f32_enqueue(s.x, s.enc_out, T, d, m.ln_post, st, nl);
    f32_to_f16_enqueue(s.enc_out, s.enc16, (size_t)T * d, st, nl);
    // cross-attention K/V for every decoder layer, written head-major into the persistent cache
    const int dd = hp.n_text_state;
    const float s4 = powf((float)(dd / hp.n_text_head), -0.25f);
    {   // one launch for all decoder layers: A = encoder output (shared), B = the stacked [K | V] projection of layer l
        const int Ld = hp.n_text_layer;
        const long wstride = Ld > 1 ? (long)(m.dec[1].ckv.w - m.dec[0].ckv.w) : (long)2 * dd * d;
        const long bstride = Ld > 1 ? (long)(m.dec[1].ckv.b - m.dec[0].ckv.b) : (long)2 * dd;
        for (int il = 0; il < Ld; il++)
            if (m.dec[il].ckv.w != m.dec[0].ckv.w + (long)il * wstride || m.dec[il].ckv.b != m.dec[0].ckv.b + (long)il * bstride)
                SS_THROW(-9, "decoder layers are not equally spaced in the weight arena");
        GemmOperand A; A.ptr = s.enc16; A.rows = T; A.ld = d;
        GemmOperand B; B.ptr = m.dec[0].ckv.w; B.rows = 2 * dd; B.ld = d; B.batch0 = Ld; B.stride0 = wstride;
        GemmEpilogue ep; ep.a_broadcast = 1; ep.bias = m.dec[0].ckv.b; ep.bias_stride0 = bstride;   // K rows have zero bias
        ep.alpha = s4; ep.alpha_cols = dd;                                                        // K pre-scaled by head_dim^-1/4
        ep.out = s.cross_k; ep.head_major = 1; ep.head_rows = T; ep.out_stride0 = (long)2 * T * dd;
        gemm_enqueue(A, B, T, 2 * dd, d, false, ep, st, nl);
    }
    CUDA_CHECK(cudaGetLastError());
}

static bool enc_graph_enabled() {
    static const bool on = [] { const char *e = getenv("SS_ENC_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}

void run_encode(State &s, int seek) {
    NvtxRange nvtx("ss.encode");
    const Model &m = s.engine->model;
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    cudaStream_t st = s.stream;
    mel_window_enqueue(m, s.d_mel, s.n_len, seek, s.win, st, &s.n_launches);      // (depends on the window: stays outside the graph)
    // programmatic dependent launch only while this replica serves a single session (kernels.h: encoder_pdl_enabled); the decision
    // is baked into the graph at capture, i.e. at the session's first window
    struct PdlScope { explicit PdlScope(bool on) { encoder_pdl_scope(on); } ~PdlScope() { encoder_pdl_scope(true); } } scope(s.engine->n_states.load() <= 1);
    if (!enc_graph_enabled()) { encode_window_enqueue(s, st, &s.n_launches); return; }
    if (!s.enc_graph) {
        // One host thread drives a State at a time; thread-local capture leaves the other sessions' streams alone.  The tensor maps and
        // kernel arguments are baked into the graph: one cudaGraphLaunch replaces ~230 launches and ~360 tensor-map encodes per window.
        cudaGraph_t graph = nullptr;
        int n = 0;
        CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try { encode_window_enqueue(s, st, &n); }
        catch (...) { cudaStreamEndCapture(st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        CUDA_CHECK(cudaStreamEndCapture(st, &graph));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) SS_THROW(-4, "cudaGraphInstantiate of the encoder pass failed: %s", cudaGetErrorString(e));
        s.enc_graph = exec; s.enc_graph_launches = n;
    }
    CUDA_CHECK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(s.enc_graph), st));
    s.n_launches += s.enc_graph_launches;
}

// ------------------------------------------------------------------------------------------------
// decoder driving
// ------------------------------------------------------------------------------------------------
void ensure_params(State &s, Decoder &d) {
    if (!d.mp_dirty) return;
    d.mp.cross_k = s.cross_k; d.mp.cross_v = s.cross_v;
    CUDA_CHECK(cudaMemcpyAsync(d.d_mp, &d.mp, sizeof(MegaParams), cudaMemcpyHostToDevice, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));   // d.mp is pageable host memory
    d.mp_dirty = false;
}

// one launch of the persistent decode kernel: runs until the device says done or `max_steps` tokens
static void run_steps(State &s, Decoder &d, int max_steps) {
    NvtxRange nvtx("ss.decode");
    ensure_params(s, d);
    decode_mega_launch(d.d_mp, d.d_ll, d.ll_bytes, max_steps, s.mega_grid, s.stream);
    s.n_launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(d.h_ctl, d.mp.ctl, offsetof(DecCtl, prompt), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

void upload_ctl(State &s, Decoder &d) {
    CUDA_CHECK(cudaMemcpyAsync(d.mp.ctl, d.h_ctl, sizeof(DecCtl), cudaMemcpyHostToDevice, s.stream));
}

// replay the decode-step graph `n_steps` times back to back (teacher-forced dummy tokens, positions
// n_past0..) and return the device time per step measured with CUDA events on the state's stream.
float bench_decode_steps(State &s, int n_steps, int n_past0) {
    const Model &m = s.engine->model; const HParams &hp = m.hp;
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    if (n_steps <= 0 || n_past0 < 0 || n_past0 + n_steps > hp.n_text_ctx) SS_THROW(-1, "bench_decode_steps: bad range");
    Decoder &d = *s.dec[0];
    ensure_params(s, d);
    DecCtl &c = *d.h_ctl;
    memset(&c, 0, sizeof c);
    // teacher-forced dummy tokens with the LM head forced on for every step: n_steps whole steps, one launch
    c.pos = n_past0; c.pos0 = n_past0; c.n_prompt = n_steps; c.sample = 0; c.all_logits = 1; c.last_id = -1; c.penult_id = -1; c.n_max = hp.n_text_ctx;
    for (int i = 0; i < n_steps; i++) c.prompt[i] = 1000 + 7 * i;
    c.token = c.prompt[0];
    (void)m;
    upload_ctl(s, d);
    CUDA_CHECK(cudaEventRecord(s.ev[2], s.stream));
    decode_mega_launch(d.d_mp, d.d_ll, d.ll_bytes, n_steps, s.mega_grid, s.stream);
    CUDA_CHECK(cudaEventRecord(s.ev[3], s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    float total = 0.f; CUDA_CHECK(cudaEventElapsedTime(&total, s.ev[2], s.ev[3]));
    s.n_launches += 1;
    if (d.mp.prof && getenv("SS_MEGA_TRACE")) {      // SS_MEGA_TRACE=1 build of the decode kernel: raw per-CTA timestamps of one step (tools/mega_trace.py)
        std::vector<long long> h((size_t)s.mega_grid * 1024);
        CUDA_CHECK(cudaMemcpy(h.data(), d.mp.prof, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(getenv("SS_MEGA_TRACE"), "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    } else if (d.mp.prof) {
        const int PN = 96;
        std::vector<long long> h((size_t)s.mega_grid * PN);
        CUDA_CHECK(cudaMemcpy(h.data(), d.mp.prof, h.size() * 8, cudaMemcpyDeviceToHost));
        const char *names[16] = {"poll (flag wait)", "gemv tiles", "-", "total", "phase QKV", "phase self-attn", "phase O", "phase CQ",
                                 "phase cross-attn", "phase CO", "phase FC1", "phase FC2", "phase LM+sample", "gemv wait ring full", "cross K wait ring full", "-"};
        const char *kinds[9] = {"QKV", "O", "CQ", "cross", "self", "CO", "FC1", "FC2", "LM"};
        const char *gst[8] = {"poll", "LN stats", "normalise", "tiles", "barrier", "fold", "epilogue", "-"};
        const char *ast[8] = {"q poll", "scores", "softmax", "PV", "fold+store", "partials fold", "-", "-"};
        auto line = [&](int k, const char *nm) {
            long long mn = h[k], mx = h[k]; double sum = 0;
            for (int c = 0; c < s.mega_grid; c++) { long long v = h[(size_t)c * PN + k]; mn = std::min(mn, v); mx = std::max(mx, v); sum += (double)v; }
            if (mx == 0) return;
            fprintf(stderr, "[mega prof] %-26s cycles/step: min %8.0f mean %8.0f max %8.0f\n", nm, (double)mn / n_steps, sum / s.mega_grid / n_steps, (double)mx / n_steps);
        };
        for (int k = 0; k < 16; k++) line(k, names[k]);
        for (int kd = 0; kd < 9; kd++) for (int sg = 0; sg < 8; sg++) {
            char nm[64]; snprintf(nm, sizeof nm, "%s: %s", kinds[kd], (kd == 3 || kd == 4) ? ast[sg] : gst[sg]);
            line(24 + kd * 8 + sg, nm);
        }
    }
    return total / n_steps;
}

void run_decode_forced(State &s, const int *tokens, int n, int n_past, float *logits_out) {
    const Model &m = s.engine->model; const HParams &hp = m.hp;
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    if (n <= 0 || n > kMaxPrompt || n_past < 0 || n_past + n > hp.n_text_ctx) SS_THROW(-1, "decode: bad token count / position");
    Decoder &d = *s.dec[0];
    ensure_params(s, d);
    DecCtl &c = *d.h_ctl;
    memset(&c, 0, sizeof c);
    c.pos = n_past; c.pos0 = n_past; c.token = tokens[0]; c.n_prompt = n; c.sample = 0; c.last_id = -1; c.penult_id = -1;
    c.n_max = hp.n_text_ctx / 2 - 4;
    for (int i = 0; i < n; i++) {
        if (tokens[i] < 0 || tokens[i] >= hp.n_vocab) SS_THROW(-1, "decode: token id out of range");
        c.prompt[i] = tokens[i];
    }
    upload_ctl(s, d);
    run_steps(s, d, n);
    s.n_decoded += n;
    CUDA_CHECK(cudaMemcpyAsync(s.h_logits, d.mp.logits, (size_t)hp.n_vocab * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    if (logits_out) memcpy(logits_out, s.h_logits, (size_t)hp.n_vocab * sizeof(float));
}

void set_sampling(Decoder &d, const FullParams &P, int tid0_init) {
    if (d.mp.suppress_blank != (int)P.suppress_blank || d.mp.tdrz != (int)P.tdrz_enable || d.mp.tid0_init != tid0_init) {
        d.mp.suppress_blank = P.suppress_blank; d.mp.tdrz = P.tdrz_enable; d.mp.tid0_init = tid0_init; d.mp_dirty = true;
    }
}

// ---- host restatement of whisper_process_logits / whisper_sample_token for the t>0 fallback path
void process_logits_host(const Model &m, const FullParams &P, Decoder &dc, const float *raw, float temperature) {
    const Vocab &v = m.vocab; const int nv = m.hp.n_vocab;
    dc.logits.assign(raw, raw + nv); dc.logprobs.resize(nv); dc.probs.resize(nv);
    float *logits = dc.logits.data(), *logprobs = dc.logprobs.data(), *probs = dc.probs.data();
    const auto &tk = dc.seq.tokens;
    const bool is_initial = tk.empty();
    if (temperature > 0.0f) for (int i = 0; i < nv; i++) logits[i] /= temperature;
    if (P.suppress_blank && is_initial) { logits[v.eot] = -INFINITY; if (v.blank >= 0) logits[v.blank] = -INFINITY; }
    logits[v.not_] = -INFINITY; logits[v.sot] = -INFINITY; logits[v.nosp] = -INFINITY;
    if (!P.tdrz_enable) logits[v.solm] = -INFINITY;
    logits[v.translate] = -INFINITY; logits[v.transcribe] = -INFINITY; logits[v.prev] = -INFINITY;
    for (int i = 0; i < kNumLangSuppress; i++) { const int t = v.sot + 1 + i; if (t < nv) logits[t] = -INFINITY; }
    {
        const bool last_ts = !tk.empty() && tk.back().id >= v.beg;
        const bool penult_ts = tk.size() < 2 || tk[tk.size() - 2].id >= v.beg;
        if (last_ts) {
            if (penult_ts) for (int i = v.beg; i < nv; i++) logits[i] = -INFINITY;
            else for (int i = 0; i < v.eot; i++) logits[i] = -INFINITY;
        }
    }
    if (is_initial && P.max_initial_ts > 0.0f) {
        const float precision = (float)kChunkSec / m.hp.n_audio_ctx;
        const int tid0 = (int)std::round(P.max_initial_ts / precision);
        for (int i = v.beg + tid0 + 1; i < nv; i++) logits[i] = -INFINITY;
    }
    if (dc.has_ts) { const int tid0 = dc.seek_delta / 2; for (int i = v.beg; i < v.beg + tid0 && i < nv; i++) logits[i] = -INFINITY; }
    {
        const float mx = *std::max_element(logits, logits + nv);
        float lse = 0.0f;
        for (int i = 0; i < nv; i++) if (logits[i] > -INFINITY) lse += expf(logits[i] - mx);
        lse = logf(lse) + mx;
        for (int i = 0; i < nv; i++) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
    }
    {
        float ts_lp = -INFINITY;
        {
            float lse = 0.0f;
            const float mx = *std::max_element(logprobs + v.beg, logprobs + nv);
            for (int i = v.beg; i < nv; i++) if (logprobs[i] > -INFINITY) lse += expf(logprobs[i] - mx);
            if (lse > 0.0f) ts_lp = logf(lse) + mx;
        }
        const float mt = *std::max_element(logprobs, logprobs + v.beg);
        if (ts_lp > mt) for (int i = 0; i < v.beg; i++) { logits[i] = -INFINITY; logprobs[i] = -INFINITY; }
    }
    for (int i = 0; i < nv; i++) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

TokData sample_token_host(const Model &m, Decoder &dc, bool best) {
    const Vocab &v = m.vocab; const int nv = m.hp.n_vocab;
    TokData r{0, 0, 0.f, 0.f, 0.f, 0.f};
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = v.beg; i < nv; i++) { sum_ts += dc.probs[i]; if (max_ts < dc.probs[i]) { max_ts = dc.probs[i]; r.tid = i; } }
        r.pt = (float)(max_ts / (sum_ts + 1e-10)); r.ptsum = (float)sum_ts;
    }
    if (best) {
        for (int i = 0; i < nv; i++) if (r.p < dc.probs[i]) { r.id = i; r.p = dc.probs[i]; r.plog = dc.logprobs[i]; }
    } else {
        std::discrete_distribution<> dist(dc.probs.begin(), dc.probs.end());
        r.id = dist(dc.rng); r.p = dc.probs[r.id]; r.plog = dc.logprobs[r.id];
    }
    if (r.id >= v.beg) { r.tid = r.id; r.pt = r.p; }
    return r;
}

void sequence_score(const FullParams &P, Sequence &q) {
    if (q.result_len == 0) return;
    double result = 0.0;
    for (int i = 0; i < q.result_len; i++) result += q.tokens[i].plog;
    q.sum_logprobs = result; q.avg_logprobs = result / q.result_len;
    double penalty = q.result_len;
    if (P.length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, P.length_penalty);
    q.score = result / penalty;
    int cnt = 0; double entropy = 0.0;
    std::map<int, int> counts;
    for (int i = std::max(0, q.result_len - 32); i < q.result_len; i++) { counts[q.tokens[i].id]++; cnt++; }
    for (const auto &kv : counts) { const double p = kv.second / (double)cnt; entropy -= p * log(p); }
    q.entropy = entropy;
}

// one forward step of decoder `d` feeding `token` at position n_past; raw logits land in s.h_logits
void step_host_sampled(State &s, Decoder &d, const int *tokens, int n, int n_past) {
    ensure_params(s, d);
    DecCtl &c = *d.h_ctl;
    memset(&c, 0, offsetof(DecCtl, prompt));
    c.pos = n_past; c.pos0 = n_past; c.token = tokens[0]; c.n_prompt = n; c.sample = 0; c.last_id = -1; c.penult_id = -1;
    for (int i = 0; i < n; i++) c.prompt[i] = tokens[i];
    upload_ctl(s, d);
    run_steps(s, d, n);
    s.n_decoded += 1;
    CUDA_CHECK(cudaMemcpyAsync(s.h_logits, d.mp.logits, (size_t)s.engine->model.hp.n_vocab * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

void kv_copy(State &s, Decoder &from, Decoder &to, int n_pos) {
    const HParams &hp = s.engine->model.hp;
    const size_t pitch = (size_t)hp.n_text_ctx * 64 * 2, width = (size_t)n_pos * 64 * 2, height = (size_t)hp.n_text_layer * hp.n_text_head;
    CUDA_CHECK(cudaMemcpy2DAsync(to.mp.self_k, pitch, from.mp.self_k, pitch, width, height, cudaMemcpyDeviceToDevice, s.stream));
    CUDA_CHECK(cudaMemcpy2DAsync(to.mp.self_v, pitch, from.mp.self_v, pitch, width, height, cudaMemcpyDeviceToDevice, s.stream));
}

// ---- beam search (whisper_full's BEAM_SEARCH strategy; the reference itself always asks for Greedy{best_of:5},
// whisper.rs:132 - this is the beam_size>1 extension BASELINE config 5 names)

// whisper_sample_token_topk: the k most likely tokens (log-prob descending, id ascending on ties), each carrying the
// same timestamp summary (tid / pt / ptsum) sample_token_host computes
std::vector<TokData> sample_topk_host(const Model &m, const Decoder &dc, int k) {
    const Vocab &v = m.vocab; const int nv = m.hp.n_vocab;
    std::vector<int> ids(nv);
    for (int i = 0; i < nv; i++) ids[i] = i;
    k = std::min(k, nv);
    std::partial_sort(ids.begin(), ids.begin() + k, ids.end(), [&](int a, int b) {
        return dc.logprobs[a] > dc.logprobs[b] || (dc.logprobs[a] == dc.logprobs[b] && a < b);
    });
    double sum_ts = 0.0, max_ts = 0.0; int tid = 0;
    for (int i = v.beg; i < nv; i++) { sum_ts += dc.probs[i]; if (max_ts < dc.probs[i]) { max_ts = dc.probs[i]; tid = i; } }
    std::vector<TokData> out;
    for (int a = 0; a < k; a++) {
        TokData t{ids[a], tid, dc.probs[ids[a]], dc.logprobs[ids[a]], (float)(max_ts / (sum_ts + 1e-10)), (float)sum_ts};
        if (t.id >= v.beg) { t.tid = t.id; t.pt = t.p; }
        out.push_back(t);
    }
    return out;
}

static bool same_tokens(const Sequence &a, const Sequence &b) {
    if (a.tokens.size() != b.tokens.size()) return false;
    for (size_t i = 0; i < a.tokens.size(); i++) if (a.tokens[i].id != b.tokens[i].id) return false;
    return true;
}

// which candidate each live decoder continues with (engine_internal.h; pure host logic, probed on the CPU by ss_debug_beam_assign)
std::vector<int> beam_pick(std::vector<BeamCandidate> &cands, const std::vector<char> &live, int i) {
    std::stable_sort(cands.begin(), cands.end(), [](const BeamCandidate &a, const BeamCandidate &b) {
        return a.seq.sum_logprobs_all > b.seq.sum_logprobs_all;
    });
    std::vector<int> pick(live.size(), -1);
    size_t cur_c = 0;
    for (size_t j = 0; j < live.size(); j++) {
        if (!live[j] || cands.empty()) continue;
        if (cur_c >= cands.size()) cur_c = 0;
        const BeamCandidate &c = cands[cur_c];
        pick[j] = (int)cur_c++;
        while (cands.size() > cur_c && i > 0 && same_tokens(cands[cur_c].seq, c.seq)) ++cur_c;
    }
    return pick;
}

// hand the best candidates to the live decoders (skipping duplicates of the one just taken) and move each decoder's
// self-attention cache to follow its new sequence. The shuffle is two-phase like whisper.cpp's temporary sequence ids:
// every moved cache is first copied into the destination decoder's second buffer, then the buffers are swapped.
void beam_advance(State &s, std::vector<BeamCandidate> &cands, int n_cur, int i, int n_past) {
    std::vector<char> live(n_cur);
    for (int j = 0; j < n_cur; j++) live[j] = !(s.dec[j]->completed || s.dec[j]->failed);
    const std::vector<int> pick = beam_pick(cands, live, i);
    const HParams &hp = s.engine->model.hp;
    const size_t kv = (size_t)hp.n_text_layer * hp.n_text_ctx * hp.n_text_state;
    std::vector<int> src(n_cur, -1);
    for (int j = 0; j < n_cur; j++) {
        if (pick[j] < 0) continue;
        Decoder &dc = *s.dec[j];
        const BeamCandidate &c = cands[pick[j]];
        dc.seek_delta = c.seek_delta; dc.has_ts = c.has_ts; dc.seq = c.seq;
        src[j] = c.decoder_idx;
    }
    const size_t pitch = (size_t)hp.n_text_ctx * 64 * 2, width = (size_t)n_past * 64 * 2, height = (size_t)hp.n_text_layer * hp.n_text_head;
    for (int j = 0; j < n_cur; j++) {
        if (src[j] < 0 || src[j] == j) continue;
        Decoder &to = *s.dec[j]; const Decoder &from = *s.dec[src[j]];
        if (!to.alt_k) {
            to.alt_k = dmalloc<__half>(kv); to.alt_v = dmalloc<__half>(kv);
            CUDA_CHECK(cudaMemsetAsync(to.alt_k, 0, kv * 2, s.stream)); CUDA_CHECK(cudaMemsetAsync(to.alt_v, 0, kv * 2, s.stream));
        }
        CUDA_CHECK(cudaMemcpy2DAsync(to.alt_k, pitch, from.mp.self_k, pitch, width, height, cudaMemcpyDeviceToDevice, s.stream));
        CUDA_CHECK(cudaMemcpy2DAsync(to.alt_v, pitch, from.mp.self_v, pitch, width, height, cudaMemcpyDeviceToDevice, s.stream));
    }
    for (int j = 0; j < n_cur; j++) {
        if (src[j] < 0 || src[j] == j) continue;
        Decoder &to = *s.dec[j];
        std::swap(to.mp.self_k, to.alt_k); std::swap(to.mp.self_v, to.alt_v);
        to.mp_dirty = true;
    }
}

// ------------------------------------------------------------------------------------------------
// whisper_full
// ------------------------------------------------------------------------------------------------
int transcribe(State &s, const float *pcm, size_t n_samples, const FullParams &P, bool stream_mode) {
    NvtxRange nvtx("ss.transcribe");
    const Model &m = s.engine->model; const HParams &hp = m.hp; const Vocab &v = m.vocab;
    CUDA_CHECK(cudaSetDevice(s.engine->device));
    s.raw.clear(); s.out.clear(); s.full_text.clear(); s.result_tokens.clear();
    s.n_fallbacks = 0; s.n_decoded = 0; s.n_windows = 0; s.n_launches = 0; s.n_keep = 0; s.h_keep.clear();
    s.ms_mel = s.ms_enc = s.ms_dec = 0;
    const bool beam = P.beam_size > 1;
    if (P.beam_size > kMaxDecoders || P.best_of > kMaxDecoders) SS_THROW(-1, "beam_size / best_of above %d", kMaxDecoders);

    int lang = 0;
    if (v.multilingual) { lang = lang_id(P.language.c_str()); if (lang < 0) SS_THROW(-6, "unknown language '%s'", P.language.c_str()); }

    if (P.keep_logits && !s.keep) {
        s.keep_cap = hp.n_text_ctx / 2; s.keep = dmalloc<float>((size_t)s.keep_cap * hp.n_vocab);
        s.dec[0]->mp.keep = s.keep; s.dec[0]->mp.keep_cap = s.keep_cap; s.dec[0]->mp_dirty = true;
    }

    CUDA_CHECK(cudaEventRecord(s.ev[0], s.stream));
    run_log_mel(s, pcm, n_samples);
    CUDA_CHECK(cudaEventRecord(s.ev[1], s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    { float ms; cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.ms_mel += ms; }

    const int seek_start = 0, seek_end = s.n_len_org;
    if (seek_end < seek_start + 100) return postprocess(s, stream_mode);

    std::vector<float> temps;
    if (P.temperature_inc > 0.0f) { for (float t = P.temperature; t < 1.0f + 1e-6f; t += P.temperature_inc) temps.push_back(t); }
    else temps.push_back(P.temperature);

    const int n_decoders = std::max(1, beam ? std::max(P.best_of, P.beam_size) : P.best_of);
    if (P.no_context) s.prompt_past.clear();

    std::vector<int> prompt_init = {v.sot};
    if (v.multilingual) { prompt_init.push_back(v.sot + 1 + lang); prompt_init.push_back(v.transcribe); }

    const int n_max = hp.n_text_ctx / 2 - 4;
    const float precision = (float)kChunkSec / hp.n_audio_ctx;
    const int tid0_init = P.max_initial_ts > 0.0f ? (int)std::round(P.max_initial_ts / precision) : -1;
    int seek = seek_start;
    std::vector<int> prompt;

    while (true) {
        if (seek + 100 >= seek_end) break;
        CUDA_CHECK(cudaEventRecord(s.ev[0], s.stream));
        run_encode(s, seek);
        CUDA_CHECK(cudaEventRecord(s.ev[1], s.stream));
        s.n_windows++;
        if (seek > seek_start && seek + 500 >= seek_end) s.prompt_past.clear();

        int best_decoder_id = 0;
        for (size_t it = 0; it < temps.size(); it++) {
            const float t_cur = temps[it];
            const int n_cur = std::max(1, beam ? (t_cur > 0.0f ? P.best_of : P.beam_size) : (t_cur > 0.0f ? n_decoders : 1));
            if (it > 0) s.n_fallbacks++;
            while ((int)s.dec.size() < n_cur) new_decoder(s, false);
            for (int j = 0; j < n_cur; j++) {
                Decoder &dc = *s.dec[j];
                dc.seq = Sequence{}; dc.seq.sum_logprobs = -INFINITY; dc.seq.avg_logprobs = -INFINITY; dc.seq.score = -INFINITY;
                dc.seek_delta = 100 * kChunkSec; dc.failed = false; dc.completed = false; dc.has_ts = false;
            }
            prompt.clear();
            if (!s.prompt_past.empty() && t_cur < 0.5f && P.n_max_text_ctx > 0) {
                const int n_take = std::min({P.n_max_text_ctx, hp.n_text_ctx / 2, (int)s.prompt_past.size()});
                prompt.push_back(v.prev);
                prompt.insert(prompt.end(), s.prompt_past.end() - n_take, s.prompt_past.end());
            }
            prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
            const int n_prompt = (int)prompt.size();

            CUDA_CHECK(cudaEventRecord(s.ev[2], s.stream));
            if (!beam && t_cur < 1e-6f) {
                // ---------------- greedy at temperature 0: whole loop on the device ----------------
                Decoder &dc = *s.dec[0];
                set_sampling(dc, P, tid0_init);
                ensure_params(s, dc);
                DecCtl &c = *dc.h_ctl;
                memset(&c, 0, sizeof c);
                c.pos = 0; c.pos0 = 0; c.token = prompt[0]; c.n_prompt = n_prompt; c.sample = 1; c.last_id = -1; c.penult_id = -1;
                c.seek = seek; c.seek_end = seek_end; c.n_max = n_max; c.seek_delta = 100 * kChunkSec;
                c.keep_logits = (P.keep_logits && it == 0) ? 1 : 0; c.n_kept = 0;
                for (int i = 0; i < n_prompt; i++) c.prompt[i] = prompt[i];
                upload_ctl(s, dc);
                run_steps(s, dc, n_prompt + n_max - 1);
                const int ns = dc.h_ctl->n_sampled;
                s.n_decoded += n_prompt - 1 + ns;
                if (ns > 0) {
                    CUDA_CHECK(cudaMemcpyAsync(dc.h_tok, dc.mp.tok_out, (size_t)ns * sizeof(TokData), cudaMemcpyDeviceToHost, s.stream));
                    CUDA_CHECK(cudaStreamSynchronize(s.stream));
                }
                dc.seq.tokens.assign(dc.h_tok, dc.h_tok + ns);
                for (int i = 0; i < ns; i++) dc.seq.sum_logprobs_all += dc.h_tok[i].plog;
                dc.seq.result_len = dc.h_ctl->result_len; dc.seek_delta = dc.h_ctl->seek_delta;
                dc.failed = dc.h_ctl->failed; dc.completed = dc.h_ctl->completed; dc.has_ts = dc.h_ctl->has_ts;
                if (c.keep_logits) {
                    const int nk = std::min(dc.h_ctl->n_kept, s.keep_cap);
                    const size_t old = s.h_keep.size();
                    s.h_keep.resize(old + (size_t)nk * hp.n_vocab);
                    CUDA_CHECK(cudaMemcpy(s.h_keep.data() + old, s.keep, (size_t)nk * hp.n_vocab * sizeof(float), cudaMemcpyDeviceToHost));
                    s.n_keep += nk;
                }
            } else if (!beam && batch_sample_enabled() && batch_beam_supported(s) && n_max <= kMaxDraws) {
                // ------- t > 0 (default; SS_BATCH_SAMPLE=0: off): the best_of sampled decoders as sequences of one batched step -------
                decode_sampled_batched(s, P, t_cur, n_cur, prompt, seek, seek_end, n_max, tid0_init);
            } else if (beam && batch_beam_enabled() && batch_beam_supported(s)) {
                // ------- default (SS_BATCH_BEAM=0: off): the live beams as sequences of one batched decoder step (engine_batch.cc) -------
                decode_beam_batched(s, P, t_cur, n_cur, prompt, seek, seek_end, n_max, tid0_init);
            } else {
                // ------- t > 0: best_of sampled decoders; beam search at any temperature: host-side sampling -------
                for (int j = 0; j < n_cur; j++) set_sampling(*s.dec[j], P, tid0_init);
                step_host_sampled(s, *s.dec[0], prompt.data(), n_prompt, 0);
                s.n_decoded += n_prompt - 1;
                process_logits_host(m, P, *s.dec[0], s.h_logits, t_cur);
                for (int j = 1; j < n_cur; j++) {
                    kv_copy(s, *s.dec[0], *s.dec[j], n_prompt);
                    s.dec[j]->probs = s.dec[0]->probs; s.dec[j]->logits = s.dec[0]->logits; s.dec[j]->logprobs = s.dec[0]->logprobs;
                }
                std::vector<BeamCandidate> cands;
                for (int i = 0; i < n_max; i++) {
                    cands.clear();
                    for (int j = 0; j < n_cur; j++) {
                        Decoder &dc = *s.dec[j];
                        if (dc.completed || dc.failed) continue;
                        if (!beam) {
                            dc.seq.tokens.push_back(sample_token_host(m, dc, false));
                            dc.seq.sum_logprobs_all += dc.seq.tokens.back().plog;
                        } else {
                            for (const TokData &t : sample_topk_host(m, dc, P.beam_size)) {
                                cands.push_back({j, dc.seek_delta, dc.has_ts, dc.seq});
                                cands.back().seq.tokens.push_back(t);
                                cands.back().seq.sum_logprobs_all += t.plog;
                            }
                        }
                    }
                    if (beam) beam_advance(s, cands, n_cur, i, n_prompt + i);
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
            CUDA_CHECK(cudaEventRecord(s.ev[3], s.stream));
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
            { float ms; cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]); s.ms_dec += ms; }
            if (it == 0) { float ms; cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.ms_enc += ms; }

            {
                double best_score = -INFINITY;
                for (int j = 0; j < n_cur; j++) {
                    Decoder &dc = *s.dec[j];
                    if (dc.failed) continue;
                    dc.seq.tokens.resize(std::min<size_t>(dc.seq.tokens.size(), (size_t)dc.seq.result_len));
                    sequence_score(P, dc.seq);
                    if (dc.seq.entropy < P.entropy_thold) { dc.failed = true; continue; }
                    if (best_score < dc.seq.score) { best_score = dc.seq.score; best_decoder_id = j; }
                }
            }
            bool success = true;
            if (it != temps.size() - 1) {
                const Decoder &dc = *s.dec[best_decoder_id];
                if (dc.failed || dc.seq.avg_logprobs < P.logprob_thold) success = false;
            }
            if (success) break;
        }
        {
            const Decoder &bd = *s.dec[best_decoder_id];
            const int seek_delta = bd.seek_delta, result_len = bd.seq.result_len;
            const auto &tc = bd.seq.tokens;
            std::vector<int> keep;
            if (prompt.front() == v.prev) keep.assign(prompt.begin() + 1, prompt.end() - prompt_init.size());
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
            seek += seek_delta;
        }
    }
    return postprocess(s, stream_mode);
}

// ------------------------------------------------------------------------------------------------
// Rust-side post-processing (whisper.rs:9-14, 41-43, 84-128, 175-201)
// ------------------------------------------------------------------------------------------------
bool is_valid_utf8(const std::string &t) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(t.data());
    size_t n = t.size(), i = 0;
    while (i < n) {
        unsigned char c = p[i];
        if (c < 0x80) { i++; continue; }
        int len; uint32_t cp;
        if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
        else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
        else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
        else return false;
        if (i + len > n) return false;
        for (int k = 1; k < len; k++) { if ((p[i + k] & 0xC0) != 0x80) return false; cp = (cp << 6) | (p[i + k] & 0x3F); }
        if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && cp < 0x10000)) return false;
        if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += len;
    }
    return true;
}

static const char *kPromo[14] = {
    "请不吝点赞", "請不吝點贊", "點贊", "訂閱", "订阅", "打赏", "打賞", "打賞支持明鏡與點點欄目", "打赏支持明镜与点点栏目",
    "並且按下小鈴鐺才能收到最新消息哦!", "請按讚、訂閱、分享!", "明镜需要您的支持 欢迎收看订阅明镜",
    "請按讚,訂閱,分享,打開小鈴鐺,並且按下小鈴鐺才能收到最新消息謝謝觀看",
    "請按讚,訂閱,分享,打開小鈴鐺,並且按下小鈴鐺才能收到最新消息哦!"};

static bool contains(const std::string &t, const char *needle) { return t.find(needle) != std::string::npos; }
static bool ends_with(const std::string &t, const char *suffix) {
    const size_t n = strlen(suffix);
    return t.size() >= n && memcmp(t.data() + t.size() - n, suffix, n) == 0;
}
std::string add_punctuation(const std::string &text) {
    if (ends_with(text, "。") || ends_with(text, "！") || ends_with(text, "？") || ends_with(text, "，")) return text;
    const bool q = contains(text, "吗") || contains(text, "呢") || contains(text, "什么") || contains(text, "为何") || contains(text, "怎么");
    const bool e = contains(text, "啊") || contains(text, "哇") || contains(text, "太") || contains(text, "真") || contains(text, "好") || contains(text, "真是");
    std::string r = text;
    if (q) r += "？"; else if (e) r += "！"; else r += " ";
    return r;
}

bool is_promotional_text(const std::string &t) {
    for (const char *p : kPromo) if (contains(t, p)) return true;
    return false;
}

int postprocess(State &s, bool stream_mode) {
    s.out.clear(); s.full_text.clear();
    const int n = (int)s.raw.size();
    int speaker = 0;
    for (int i = 0; i < n; i++) {
        const std::string &text = s.raw[i].text;
        if (!is_valid_utf8(text) || text.find('\0') != std::string::npos) {   // full_get_segment_text -> Err -> `?` (whisper.rs:85)
            s.out.clear(); s.full_text.clear();
            SS_THROW(-7, "segment %d text is not valid UTF-8", i);
        }
        if (is_promotional_text(text)) continue;
        if (i > 0 && s.raw[i - 1].speaker_turn_next) speaker++;
        const std::string processed = add_punctuation(text);
        if (stream_mode) {
            if (i == n - 1) { s.out.push_back({processed, speaker, (double)s.raw[i].t0, (double)s.raw[i].t1}); s.full_text = processed; }
        } else {
            s.out.push_back({processed, speaker, (double)s.raw[i].t0, (double)s.raw[i].t1});
            s.full_text += processed;
        }
    }
    return 0;
}

}  // namespace ss
