/*
 * whisper_oracle.h - CPU restatement of the transcribe hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for speaksense_b200: a plain-C restatement of what sits behind
 * `state.full(params, &audio)` at /root/reference/src/asr/whisper.rs:75 (whisper-rs 0.11.1 ->
 * whisper-rs-sys 0.9.0 -> bundled whisper.cpp ~v1.5.x; NOT vendored in the reference tree, see
 * SURVEY.md §8c and Appendix A).  PARITY UNPINNED: the reference holds no golden vectors for this
 * path, so the oracle is pinned against HuggingFace `transformers` Whisper with identical weights
 * (tests/golden/, tools/make_golden.py) instead.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (csrc/, libspeaksense_whisper.so) never links or calls it.
 */
#ifndef WHISPER_ORACLE_H
#define WHISPER_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wo_model wo_model;
typedef struct wo_state wo_state;

typedef struct {
    int n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
} wo_hparams;

/* decoding parameters as fixed by reference build_params (src/asr/whisper.rs:131-173) plus the
 * per-request overrides (:60-71).  wo_default_params() fills the reference's values. */
typedef struct {
    const char *language;      /* NULL -> "en" (library default)                  whisper.rs:60-63 */
    int   tdrz_enable;         /* speaker_diarization                              whisper.rs:137-140 */
    int   no_context;          /* false; true in stream mode                       whisper.rs:155,67 */
    int   single_segment;      /* false                                            whisper.rs:148 */
    int   best_of;             /* Greedy{best_of:5}                                whisper.rs:132 */
    int   beam_size;           /* 0/1 = greedy; >1 = beam search (extension) */
    float temperature;         /* 0.0                                              whisper.rs:159 */
    float temperature_inc;     /* 0.2 library default */
    float entropy_thold;       /* 2.4                                              whisper.rs:160 */
    float logprob_thold;       /* -1.0                                             whisper.rs:161 */
    float max_initial_ts;      /* 1.0                                              whisper.rs:153 */
    float length_penalty;      /* -1.0                                             whisper.rs:170 */
    int   suppress_blank;      /* true library default */
    int   n_max_text_ctx;      /* 16384 library default */
    int   max_tokens;          /* 0                                                whisper.rs:165 */
    int   n_threads;           /* 16                                               whisper.rs:143 */
    int   keep_logits;         /* debug: keep per-step raw logits of decoder 0 at temperature[0] */
} wo_params;

void wo_default_params(wo_params *p);

wo_model *wo_load(const char *path);           /* NULL on error, see wo_last_error() */
void      wo_free(wo_model *m);
const wo_hparams *wo_get_hparams(const wo_model *m);
const char *wo_last_error(void);
int  wo_token_id(const wo_model *m, const char *name);   /* eot sot translate transcribe solm prev nosp not beg */
int  wo_lang_id(const char *lang);                        /* -1 unknown */
int  wo_token_bytes(const wo_model *m, int id, const char **bytes); /* returns length */
void wo_set_threads(int n);

/* ---- stage functions (teacher-forced parity) ---- */
/* PCM -> log-mel.  *mel is malloc'd [n_mels][n_len] f32 (caller frees with wo_free_buf). */
int  wo_log_mel(const wo_model *m, const float *pcm, size_t n, float **mel, int *n_len, int *n_len_org);
void wo_free_buf(void *p);

wo_state *wo_state_new(wo_model *m);
void      wo_state_free(wo_state *s);
/* encoder on mel frames [seek, seek+2*n_audio_ctx) (zero-filled past n_len) + cross-KV. */
int  wo_encode(wo_state *s, const float *mel, int n_len, int seek);
const float *wo_encoder_out(const wo_state *s);           /* [n_audio_ctx][n_audio_state] f32 */
/* decoder over `n` tokens at positions n_past.. of KV sequence `seq`; logits of the last token. */
int  wo_decode(wo_state *s, int seq, const int *tokens, int n, int n_past, float *logits_out);

/* ---- whole path == WhisperState::full ---- */
int  wo_full(wo_state *s, const float *pcm, size_t n, const wo_params *p);
int  wo_n_segments(const wo_state *s);
const char *wo_segment_text(const wo_state *s, int i);    /* raw bytes, NUL terminated */
int64_t wo_segment_t0(const wo_state *s, int i);
int64_t wo_segment_t1(const wo_state *s, int i);
int  wo_segment_speaker_turn_next(const wo_state *s, int i);
/* diagnostics of the last wo_full: tokens of all windows concatenated */
int  wo_n_result_tokens(const wo_state *s);
int  wo_result_token(const wo_state *s, int i, float *p, float *plog);
int  wo_n_fallbacks(const wo_state *s);                   /* temperature steps beyond the first, summed */
int  wo_n_decoded(const wo_state *s);                     /* decoder forward passes (tokens) */
int  wo_n_windows(const wo_state *s);
int  wo_n_kept_logits(const wo_state *s);
const float *wo_kept_logits(const wo_state *s, int step); /* raw logits [n_vocab] */

/* diagnostics of the t > 0 draws of the last wo_full: how many uniforms were consumed, and the smallest distance of a uniform to a
 * boundary of the interval it fell into (cumulative probability units).  A second implementation whose probabilities differ by
 * less than that margin must draw the same tokens. */
double wo_min_sample_margin(const wo_state *s);
long   wo_n_draws(const wo_state *s);

/* test probe: the logits filter (whisper_process_logits) on a given history of sampled token ids and decoder timestamp
 * state; logits_out [n_vocab], -inf = masked */
int  wo_probe_process_logits(wo_state *s, const wo_params *p, const int *ids, int n_ids, int has_ts, int seek_delta,
                             const float *raw, float temperature, float *logits_out);

/* test probes: whisper_full's per-token bookkeeping over a list of sampled ids (out = failed, completed, result_len, seek_delta,
 * has_ts, steps consumed) and whisper_sequence_score + the entropy gate's statistic (out = sum_logprobs, avg_logprobs, entropy, score) */
int  wo_probe_bookkeeping(wo_state *s, const wo_params *p, const int *ids, int n, int seek, int seek_end, int n_max, int *out);
int  wo_probe_score(wo_state *s, const wo_params *p, const int *ids, const float *plogs, int n, int result_len, double *out);

/* test probe: draws of the restated std::mt19937 + std::discrete_distribution<> pair */
int  wo_probe_sample(wo_state *s, uint32_t seed, const float *probs, int n, int count, int *out);

#ifdef __cplusplus
}
#endif
#endif
