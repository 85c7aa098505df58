// engine_internal.h - helpers of engine.cc shared with engine_batch.cc (not part of the library's interface).
#pragma once
#include <nvtx3/nvToolsExt.h>

#include "engine.h"

namespace ss {

// NVTX range over a host-side stage (header-only nvtx3: a no-op unless a tool such as Nsight Systems is attached).  Ranges:
// ss.transcribe > ss.log_mel / ss.encode / ss.decode, ss.transcribe_batch > ss.batch.encode / ss.batch.decode, ss.denoise.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

Decoder *new_decoder(State &s, bool with_keep);
void ensure_params(State &s, Decoder &d);
void upload_ctl(State &s, Decoder &d);
void set_sampling(Decoder &d, const FullParams &P, int tid0_init);
// host restatement of whisper_process_logits / whisper_sample_token (t > 0 fallback decoders)
void process_logits_host(const Model &m, const FullParams &P, Decoder &dc, const float *raw, float temperature);
TokData sample_token_host(const Model &m, Decoder &dc, bool best);
void sequence_score(const FullParams &P, Sequence &q);
// one forward step of decoder `d` feeding `tokens` at position n_past; raw logits of the last one land in s.h_logits
void step_host_sampled(State &s, Decoder &d, const int *tokens, int n, int n_past);
void kv_copy(State &s, Decoder &from, Decoder &to, int n_pos);

// ---- beam search (whisper_full's BEAM_SEARCH strategy)
struct BeamCandidate { int decoder_idx, seek_delta; bool has_ts; Sequence seq; };
std::vector<TokData> sample_topk_host(const Model &m, const Decoder &dc, int k);
// whisper_full's candidate assignment of one beam-search step, host only: sorts `cands` by sum_logprobs_all (descending, stable) and
// returns, per decoder, the index of the candidate it continues with (-1: not live).  From the second sampled token on (i > 0) a
// decoder skips the candidates behind its own that carry the same tokens; the cursor wraps.  (ss_debug_beam_assign probes it on CPU.)
std::vector<int> beam_pick(std::vector<BeamCandidate> &cands, const std::vector<char> &live, int i);
void beam_advance(State &s, std::vector<BeamCandidate> &cands, int n_cur, int i, int n_past);
// (SS_BATCH_BEAM, on by default): one window's beam search with the live beams as sequences of one batched decoder step
// (SS_BATCH_SAMPLE, on by default): the best_of sampled decoders of a t > 0 rung as sequences of one batched decoder step, drawn on the
// device (bd_sample_kernel, sample == 3) from uniforms the host generates with each decoder's own std::mt19937
bool batch_sample_enabled();
void decode_sampled_batched(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek, int seek_end,
                            int n_max, int tid0_init);
bool batch_beam_enabled();
bool batch_beam_supported(const State &s);
void decode_beam_batched(State &s, const FullParams &P, float t_cur, int n_cur, const std::vector<int> &prompt, int seek, int seek_end,
                         int n_max, int tid0_init);

}  // namespace ss
