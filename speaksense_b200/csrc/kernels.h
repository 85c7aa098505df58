// kernels.h - device-side argument blocks and host launchers of the sm_100a kernels.
#pragma once
#include "common.h"
#include "model.h"

namespace ss {

// ---------------------------------------------------------------- log-mel (mel.cu)
// Raw log10-mel for frames [0, n_len) of one clip into mel[n_mels][n_len] plus the clip's running
// maximum (float bits, ordered-int atomicMax), then in-place clamp/scale, then window extraction.
void mel_enqueue(const Model &m, const float *d_pcm, size_t n_samples, float *d_mel, int n_len,
                 int *d_max_bits, cudaStream_t st, int *launches);
// frames [seek, seek+2*n_audio_ctx) -> f16 [2*n_audio_ctx + 2][n_mels] (zero rows at both ends,
// zero past n_len) = the conv1 implicit-GEMM operand.
void mel_window_enqueue(const Model &m, const float *d_mel, int n_len, int seek, __half *d_win,
                        cudaStream_t st, int *launches);
inline int mel_n_len(size_t n_samples) { return (int)((n_samples + (size_t)kSampleRate * kChunkSec) / kHop); }
inline int mel_n_len_org(size_t n_samples) { return 1 + (int)(((long)n_samples + kNFft / 2 - kNFft) / kHop); }

// ---------------------------------------------------------------- audio denoise (denoise.cu; reference src/audio/mod.rs:507-735)
size_t denoise_scratch_floats(size_t n, int frame_size, float overlap);
// denoise_audio(d_in[0..n)) -> d_out on `st`; returns the noise type (0 stationary / 1 non-stationary / 2 mixed)
int denoise_enqueue(const float *d_in, size_t n, int frame_size, float overlap, float strength, float *d_out, float *scratch,
                    cudaStream_t st, int *launches, float *nv_out);

// n_frames independent StreamAudioProcessor frames of frame_size samples each: denoise + noise gate (audio/mod.rs:131-139)
void denoise_frames_enqueue(const float *d_in, int n_frames, int frame_size, float strength, float noise_gate, float *d_out,
                            cudaStream_t st, int *launches);

// ---------------------------------------------------------------- tcgen05 GEMM (gemm_sm100.cu)
struct GemmOperand {        // a K-major (or, for B, MN-major) f16 operand described as up to 4-D strided view
    const __half *ptr = nullptr;
    long rows = 0;          // extent of the row (M or N) dimension per batch
    long ld = 0;            // elements between rows
    long batch0 = 1, stride0 = 0;   // inner batch (e.g. head)
    long batch1 = 1, stride1 = 0;   // outer batch (e.g. clip)
};
enum : int { GEMM_OUT_F32 = 0, GEMM_OUT_F16 = 1 };
struct GemmEpilogue {
    const float *bias = nullptr;     // [N]
    float alpha = 1.0f;              // applied to columns < alpha_cols after bias
    int alpha_cols = 0;
    int gelu = 0;                    // ggml f16-LUT GELU semantics
    const float *pos = nullptr;      // + pos[(m % pos_rows)][n], ld = N
    int pos_rows = 1;
    int residual = 0;                // out (f32) += value
    int out_type = GEMM_OUT_F16;
    void *out = nullptr;
    long out_ld = 0, out_stride0 = 0, out_stride1 = 0;   // elements
    int head_major = 0;              // cross-KV cache layout: row = (n/64) * head_rows + m, 64 halfs per row as eight 16-byte chunks with
                                     // chunk c of row m at position c ^ (m & 7) (+ batch strides)
    long head_rows = 0;
    int out_row_offset = 0;          // rows shift (padded conv layouts)
    int a_broadcast = 0;             // A has no batch dimension (B / bias / out do)
    long bias_stride0 = 0;           // bias elements between inner batches
};
// D[b][M][N] = A[b][M][K] . B[b][N][K]^T  (B K-major) or A . B[b][K][N] (b_mn_major)
void gemm_enqueue(const GemmOperand &A, const GemmOperand &B, int M, int N, int K, bool b_mn_major,
                  const GemmEpilogue &ep, cudaStream_t st, int *launches);
void gemm_init();   // resolves cuTensorMapEncodeTiled, sets kernel attributes
// 4-D TMA descriptor (inner, rows, batch0, batch1) of an f16 operand with 128-byte swizzle; shared with attention_sm100.cu
void make_tensor_map_4d(void *cu_tensor_map /* CUtensorMap* */, const GemmOperand &op, long inner, long rows, int box_inner, int box_rows);

// ---------------------------------------------------------------- fused encoder attention (attention_sm100.cu)
void attention_init();
// qkv: [clips][T][3*H*64] f16 (Q | K | V) -> out: [clips][T][H*64] f16 = softmax(Q K^T * scale) V per head
void attention_enqueue(const __half *qkv, __half *out, int clips, int T, int H, float scale, cudaStream_t st, int *launches);

// ---------------------------------------------------------------- encoder helpers (encoder.cu)
void layernorm_f16_enqueue(const float *x, __half *y, int rows, int d, const LNp &ln, cudaStream_t st, int *launches);
void layernorm_f32_enqueue(const float *x, float *y, int rows, int d, const LNp &ln, cudaStream_t st, int *launches);
void f32_to_f16_enqueue(const float *x, __half *y, size_t n, cudaStream_t st, int *launches);

// ---------------------------------------------------------------- decoder (decoder.cu)
constexpr int kNumLangSuppress = 100;
constexpr int kMaxPrompt = 448;
constexpr int kMaxDraws = 256;     // uniforms a sampled decoder can consume in one window (n_max = n_text_ctx / 2 - 4 <= 252)

struct TokData { int id, tid; float p, plog, pt, ptsum; };

struct DecCtl {   // device resident; written by dec_sample_kernel, read by every kernel of the step
    int pos, token, done, n_sampled;
    int has_ts, seek_delta, result_len, failed, completed;
    int last_id, penult_id;
    int n_prompt, pos0;
    int seek, seek_end, n_max;
    int sample, keep_logits, n_kept;
    int all_logits;   // compute the LM head for every token, not only from the last prompt token on
    float temperature;   // sample == 3 (t > 0 fallback decoders on the batched step): logits are divided by it before the filter
    int prompt[kMaxPrompt];
    // sample == 3: draw i of this window = std::generate_canonical<double, 53>(mt19937), generated by the host from the decoder's
    // own generator (whisper.cpp: one std::mt19937 per decoder) so that host- and device-sampled paths consume the same stream
    double u[kMaxDraws];
};

struct MegaLayer {
    const __half *w[6];          // qkv, o, cq, co, fc1, fc2
    const float *b[6];
    const float *lnw[3], *lnb[3];   // attn_ln, cross_attn_ln, mlp_ln
};
constexpr int kMaxLayers = 32;
typedef unsigned long long ss_u64;
struct MegaParams {   // device-resident descriptor of one decoder sequence (decoder_mega.cu)
    int d, H, L, T, ctx, n_vocab;
    int xsplit;                  // key splits per head of the cross-attention phase
    float s4;                    // head_dim^-1/4
    const __half *tok_emb; const float *d_pos; const float *lnf_w, *lnf_b;
    MegaLayer layer[kMaxLayers];
    DecCtl *ctl;
    // flagged exchange buffers ({epoch, float} words), all inside one arena that is zeroed per launch
    ss_u64 *xA, *xB, *xC, *q1, *kcur, *vcur, *att1, *q2, *att2, *hbuf, *part, *stats;
    float *logits; TokData *tok_out; float *keep; int keep_cap;
    __half *self_k, *self_v;              // [layer][head][n_text_ctx][64]
    const __half *cross_k, *cross_v;      // [layer][K | V][head][n_audio_ctx][64]: layer stride 2*T*d, V = K + T*d
    long long *prof;             // optional [grid][8] cycle counters (SS_MEGA_PROF=1), else null
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, blank;
    int suppress_blank, tdrz, tid0_init;
};
size_t decode_mega_smem_bytes();
void decode_mega_configure();
int decode_mega_grid(int device);
void decode_mega_launch(const MegaParams *d_params, void *d_ll, size_t ll_bytes, int max_steps, int grid, cudaStream_t st);

// ---------------------------------------------------------------- batched decoder (decoder_batch.cu)
// One decode step for up to kMaxBatch independent sequences (the clips of ss_transcribe_batch; BASELINE configs 3/4):
// every weight matrix is streamed from HBM once per step for the whole batch (skinny tensor-core GEMMs, weights as the
// MMA A operand, the sequences on the 8-wide N dimension), attention runs per (sequence, head).  A sequence borrows the
// control block, self-KV cache and token buffer of decoder 0 of its State and the State's cross-KV cache.
constexpr int kMaxBatch = 32;
struct BatchSeq {
    DecCtl *ctl;
    __half *self_k, *self_v;              // [layer][head][n_text_ctx][64]
    const __half *cross_k, *cross_v;      // [layer][K | V][head][n_audio_ctx][64]
    TokData *tok_out;
};
struct BatchParams {
    int B;                                // live sequences (<= kMaxBatch); operand buffers hold 8 * n_tiles rows
    int d, H, L, T, ctx, n_vocab;
    float s4;
    const __half *tok_emb; const float *d_pos; const float *lnf_w, *lnf_b;
    float *x;                             // residual stream f32 [rows][d]
    __half *xn, *q, *att, *hid;           // f16 operands [rows][d] ([rows][4d] for hid)
    float *logits;                        // [rows][n_vocab]
    float *part;                          // cross-attention partial records [B][H][xsplit][66] when the keys are split
    int *n_done;                          // sequences that have finished (every kernel returns at once when == B)
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, blank;
    int suppress_blank, tdrz, tid0_init;
    BatchSeq seq[kMaxBatch];
};
// scratch floats / halfs a batch of `rows` operand rows needs (see decode_batch_bind)
size_t decode_batch_scratch_bytes(int d, int n_vocab, int H, int xsplit_max);
// carve the operand buffers of `P` out of one zero-initialised device allocation
void decode_batch_bind(BatchParams &P, void *scratch);
int decode_batch_xsplit(int B, int H, int sms);
// enqueue one step (all layers; LM head + sampling / prompt feeding) for the live sequences of P.  `w` carries the
// weight pointers (any decoder's MegaParams of the same engine).
struct BeamStep { float temperature; int k; TokData *cand; };   // cand: device [kMaxBatch][8]
// `beam` != nullptr (opt-in beam search on the batched step): the step ends with k candidates per sequence instead of a greedy token
void decode_batch_step_enqueue(const BatchParams &P, const MegaParams &w, bool need_logits, int xsplit, cudaStream_t st, int *launches,
                               const BeamStep *beam = nullptr);
// The same step as a CUDA graph: the ~390 launches (with their programmatic-dependent-launch edges) are captured once per
// (parameters, need_logits) and replayed for every token of the round - every per-token quantity lives in device memory (control
// blocks), so the kernel arguments do not change.  A caller keeps one BatchStepGraph per round; a change of the arguments
// (live beams leaving, another xsplit) re-captures.  SS_BATCH_GRAPH=0: plain launches.
struct BatchStepGraph {
    struct Slot { void *exec = nullptr; int n_launch = 0; bool valid = false; BatchParams P; int xsplit = 0; BeamStep beam{0.f, 0, nullptr}; bool has_beam = false; };
    Slot slot[2];      // [need_logits]
    ~BatchStepGraph();
};
void decode_batch_step_graph(BatchStepGraph &G, const BatchParams &P, const MegaParams &w, bool need_logits, int xsplit, cudaStream_t st,
                             int *launches, const BeamStep *beam = nullptr);

// ---- programmatic dependent launch (chains of short kernels: the batched decoder step, the encoder pass of one clip) ----
// A kernel launched with the attribute may start while its predecessor in the stream is still running; it must execute pdl_wait()
// before it touches anything the predecessor reads or writes.  pdl_trigger() lets the successor's launch begin.  Without the
// attribute both instructions are no-ops, and a successor without the attribute waits for full completion as usual.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
template <typename... KArgs, typename... Args>
inline void launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}
// Whether the encoder's kernels are launched with the attribute: SS_ENC_PDL (default on) AND the calling thread's scope.  A dependent
// grid that starts early holds its SMs (a GEMM CTA: ~200 KB of shared memory) until its predecessor is done - free on a GPU that one
// session owns, but taken from the other sessions' kernels on a shared one (measured: 8 concurrent streams 209 / 218x with, 271 /
// 232x without; one clip alone 4.60 against 4.65 ms).  run_encode therefore narrows the scope to "this replica has one session".
bool encoder_pdl_enabled();
void encoder_pdl_scope(bool on);      // thread-local; true by default

}  // namespace ss
