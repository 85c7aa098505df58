/*
 * speaksense_whisper.h - C ABI of the B200-native Whisper engine (libspeaksense_whisper.so).
 *
 * Drop-in boundary for the hot path behind SpeakSense's `AsrEngine` trait
 * (/root/reference/src/asr/mod.rs:58-73) as implemented by `WhisperAsr`
 * (/root/reference/src/asr/whisper.rs:16-129).  Each entry point names the whisper-rs call it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` stub a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; no exceptions cross the boundary; every int-returning
 * function returns 0 on success and a negative ss_status on error, with a thread-local message in
 * ss_last_error().  There is NO CPU fallback: every compute entry point fails with SS_ERR_NO_DEVICE
 * when no sm_100 device is usable.
 */
#ifndef SPEAKSENSE_WHISPER_H
#define SPEAKSENSE_WHISPER_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 1

typedef enum {
    SS_OK = 0,
    SS_ERR_INVALID = -1,      /* bad argument */
    SS_ERR_IO = -2,           /* model file unreadable / malformed            (whisper.rs:23-24) */
    SS_ERR_NO_DEVICE = -3,    /* no CUDA device / not sm_100: no CPU fallback */
    SS_ERR_CUDA = -4,         /* CUDA runtime / driver error */
    SS_ERR_OOM = -5,
    SS_ERR_LANGUAGE = -6,     /* unknown language code                        (whisper.rs:60-63) */
    SS_ERR_UTF8 = -7,         /* segment text is not valid UTF-8: the reference's
                                 full_get_segment_text fails the whole call   (whisper.rs:85) */
    SS_ERR_NCCL = -8,
    SS_ERR_INTERNAL = -9
} ss_status;

/* == whisper_rs::WhisperContext: immutable after open, thread-safe, shared by all states
 *    (whisper.rs:17,26 `Arc<WhisperContext>`). */
typedef struct ss_engine ss_engine;
/* == whisper_rs::WhisperState: KV caches + scratch of one session, pinned to the engine's device,
 *    NOT thread-safe (the reference guards it with a Mutex, whisper.rs:51-54). */
typedef struct ss_state ss_state;

/* == AsrParams (src/asr/mod.rs:9-15) + what transcribe_with_state derives from it
 *    (whisper.rs:58-71).  Everything else is fixed by build_params (whisper.rs:131-173). */
typedef struct {
    const char *language;        /* NULL => library default "en"; "zh" | "en" | "ja" | ...        */
    int speaker_diarization;     /* set_tdrz_enable(true)                     whisper.rs:137-140 */
    int stream_mode;             /* no_context=true, keep only last segment   whisper.rs:65-69,102-111 */
    int min_segment_length;      /* carried, never read by the reference      mod.rs:14 */
    int beam_size;               /* 0/1 => Greedy{best_of:5} like the reference (whisper.rs:132);
                                    >1 => beam search (extension for BASELINE config 5) */
    int debug_keep_logits;       /* keep raw per-step logits of the t=0 decoder for ss_debug_logits */
} ss_params;

void ss_params_default(ss_params *p);                     /* == AsrParams::new()   mod.rs:18-25 */

int  ss_abi_version(void);
const char *ss_last_error(void);
const char *ss_build_info(void);                          /* arch / compiler string */

/* == WhisperContext::new_with_params(path, default)      whisper.rs:23
 * Loads a legacy ggml .bin (f16) onto CUDA device `device`. */
int  ss_engine_open(const char *ggml_path, int device, ss_engine **out);
/* Data-parallel open: rank 0 reads the file, the packed weight arena is ncclBroadcast
 * to the other ranks over NVLink; `nccl_id` is the 128-byte ncclUniqueId from ss_nccl_unique_id()
 * on rank 0, ferried by the host (torch.distributed / env).  ggml_path may be NULL on ranks > 0. */
int  ss_nccl_unique_id(unsigned char out[128]);
int  ss_engine_open_dist(const char *ggml_path, int device, int rank, int world,
                         const unsigned char nccl_id[128], ss_engine **out);
/* One process that owns several GPUs (the reference is ONE server process whose context is shared by every
 * session: src/main.rs:38-39, src/asr/whisper.rs:17,26).  The file is parsed once; the packed arena is uploaded to devices[0]
 * and broadcast to the other devices with an in-process ncclBroadcast (ncclCommInitAll + one group call) over NVLink.
 * The engine then holds one immutable weight replica per device; every state is pinned to one of them. */
int  ss_engine_open_multi(const char *ggml_path, const int *devices, int n_devices, ss_engine **out);
int  ss_engine_n_devices(const ss_engine *e);
int  ss_engine_device(const ss_engine *e, int i);          /* CUDA ordinal of replica i, -1 if out of range */
int  ss_engine_n_states(const ss_engine *e, int i);        /* live states on replica i */
/* FNV-1a (64 bit) of replica i's weight arena as it sits in HBM: equals ss_model_probe()'s arena_fnv1a when the upload / broadcast
 * was exact (bench.py checks it on every rank) */
int  ss_engine_arena_fnv1a(const ss_engine *e, int i, uint64_t *out);
void ss_engine_close(ss_engine *e);                        /* refcounted: states keep it alive */
int  ss_engine_info(const ss_engine *e, int *n_vocab, int *n_audio_state, int *n_audio_layer,
                    int *n_text_layer, int *n_mels, int64_t *weight_bytes);

/* == WhisperContext::create_state()                      whisper.rs:30-39 */
int  ss_state_new(ss_engine *e, ss_state **out);            /* multi-device engine: == ss_state_new_on(e, -1, out) */
/* state on the replica that lives on CUDA device `device`; -1 = the replica with the fewest live states */
int  ss_state_new_on(ss_engine *e, int device, ss_state **out);
int  ss_state_device(const ss_state *s);
/* outcome of the last transcribe on this state: 0 or the negative ss_status; a batch call fails clip by clip (the read-back of one
 * clip failing - whisper.rs:85 - leaves the other clips' results in place) */
int  ss_state_status(const ss_state *s);
const char *ss_state_error(const ss_state *s);
void ss_state_free(ss_state *s);

/* == transcribe_with_state up to and including the segment read-back and Rust post-processing
 *    (whisper.rs:45-129): build_params + state.full + promo filter + punctuation + stream-mode
 *    last-segment rule.  Blocking.  pcm: mono 16 kHz f32 HOST memory. */
int  ss_transcribe(ss_engine *e, ss_state *s, const float *pcm, size_t n_samples, const ss_params *p);
/* Batched form (data-parallel inside one GPU; BASELINE configs 3/4): states[i] receives the result for pcm[i], exactly
 * what ss_transcribe(states[i], pcm[i]) would leave there.  The clips' temperature-0 greedy decodes share one batched
 * decoder step per token (the weights are streamed once per step for the whole batch); fallbacks run per clip.  The states
 * must be distinct for that; beam_size > 1 / debug_keep_logits / batches below SS_BATCH_MIN (default 4) clips / SS_BATCH_DECODE=0 in
 * the environment decode clip by clip.
 * pcm[i] == NULL: clip i is the resident PCM of states[i] (ss_upload_pcm / ss_denoise_audio), n_samples[i] is ignored. */
int  ss_transcribe_batch(ss_engine *e, ss_state *const *states, const float *const *pcm,
                         const size_t *n_samples, int batch, const ss_params *p);

/* Device-resident form used for kernel-side throughput measurement: ss_upload_pcm stages the clip in
 * HBM once, ss_transcribe_resident runs the same path without the host->device copy. */
int  ss_upload_pcm(ss_engine *e, ss_state *s, const float *pcm, size_t n_samples);
int  ss_transcribe_resident(ss_engine *e, ss_state *s, const ss_params *p);
/* ---- audio denoise in front of the hot path ----
 * == DenoiseConfig (src/audio/mod.rs:41-62; defaults :51-61) */
typedef struct ss_denoise_config {
    int   frame_size;               /* 2048 (power of two, <= 4096) */
    float overlap;                  /* 0.75 */
    float strength;                 /* 0.2 */
    float noise_gate;               /* 0.003 (StreamAudioProcessor only) */
    int   enable_noise_reduction;   /* 1     (StreamAudioProcessor only) */
    float threshold;                /* 0.002 (never read by the reference) */
} ss_denoise_config;
void ss_denoise_config_default(ss_denoise_config *c);
/* == denoise_audio(&samples, &config) (src/audio/mod.rs:507-528), as called per 5 s chunk by the gRPC handler
 *    (grpc/handlers/asr.rs:196) right before transcribe_with_state: noise-type analysis, spectral subtraction and / or
 *    Wiener filter over Hann STFT frames, overlap-add with the reference's scaling (unnormalised inverse FFT, x10).
 *    pcm: HOST f32.  The denoised chunk stays resident in HBM as the state's PCM, so the caller may follow with
 *    ss_transcribe_resident (no second host->device copy); `out` (HOST, n_samples floats) may be NULL.
 *    noise_type: 0 stationary, 1 non-stationary, 2 mixed.  Error if n_samples < frame_size (the reference panics). */
int  ss_denoise_audio(ss_engine *e, ss_state *s, const float *pcm, size_t n_samples, const ss_denoise_config *cfg,
                      float *out, int *noise_type, float *spectral_variance);
/* == steps 4-5 of StreamAudioProcessor::process_frame (src/audio/mod.rs:131-139) for n_frames independent frames of
 *    cfg->frame_size samples (the REST path's 2048-sample frames, already scaled by the VAD gain): denoise_audio on the
 *    single frame + noise gate.  frames / out: HOST, n_frames * frame_size floats.  One launch for all frames. */
int  ss_denoise_frames(ss_engine *e, ss_state *s, const float *frames, int n_frames, const ss_denoise_config *cfg, float *out);
/* Replays the decode-step CUDA graph n_steps times at positions n_past0.. (dummy tokens) and returns
 * the device time per step (CUDA events on the state's stream): the roofline probe of stage 3. */
int  ss_bench_decode_steps(ss_engine *e, ss_state *s, int n_steps, int n_past0, float *ms_per_step);

/* results of the last transcribe on this state; pointers valid until the next call on it.
 * "raw" == what whisper-rs returns (full_n_segments / full_get_segment_*; whisper.rs:77-95);
 * the un-prefixed accessors == TranscribeResult after the Rust post-processing (whisper.rs:84-128). */
int  ss_n_segments_raw(const ss_state *s);
const char *ss_segment_text_raw(const ss_state *s, int i);
int64_t ss_segment_t0_raw(const ss_state *s, int i);       /* 10 ms ticks */
int64_t ss_segment_t1_raw(const ss_state *s, int i);
int  ss_segment_speaker_turn_next_raw(const ss_state *s, int i);

int  ss_n_segments(const ss_state *s);                     /* TranscribeResult.segments.len() */
const char *ss_segment_text(const ss_state *s, int i);     /* UTF-8 */
double ss_segment_start(const ss_state *s, int i);         /* t0 ticks as f64   whisper.rs:107 */
double ss_segment_end(const ss_state *s, int i);
int  ss_segment_speaker_id(const ss_state *s, int i);
const char *ss_full_text(const ss_state *s);               /* TranscribeResult.full_text */

/* diagnostics of the last transcribe */
int  ss_n_result_tokens(const ss_state *s);
int  ss_result_token(const ss_state *s, int i, float *p, float *plog);
int  ss_n_fallbacks(const ss_state *s);
int  ss_n_decoded(const ss_state *s);
int  ss_n_windows(const ss_state *s);
int  ss_n_kernel_launches(const ss_state *s);              /* our kernels launched by the last call */
/* device time of the stages of the last call, ms (CUDA events on the state's stream) */
int  ss_stage_ms(const ss_state *s, float *mel_ms, float *encoder_ms, float *decode_ms);
const float *ss_debug_logits(const ss_state *s, int step, int *n_vocab);  /* parity hook; HOST ptr */

/* ---- host-only helpers (no device needed): the Rust-side text rules of whisper.rs, exposed so that the
 *      CPU test-suite can pin them, and the ggml loader's view of a model file ---- */
int  ss_is_promotional_text(const char *utf8);                       /* whisper.rs:41-43 (list at :9-14) */
int  ss_add_punctuation(const char *utf8, char *out, size_t out_cap);  /* whisper.rs:175-201; returns bytes written or <0 */
int  ss_is_valid_utf8(const char *bytes, size_t n);                  /* what full_get_segment_text enforces, whisper.rs:85 */
/* whisper_process_logits as the engine's host-side decoders run it (process_logits_host: host-stepped fallback decoders, the first
 * step of a beam search), on a given history of sampled token ids and raw logits [n_vocab] of the model file's vocabulary, with
 * build_params' constants; out = the filtered logits (-inf = masked) */
int  ss_debug_process_logits(const char *ggml_path, const int *ids, int n_ids, int has_ts, int seek_delta, const float *raw,
                             float temperature, float *out);
/* whisper_sequence_score + the entropy of the last 32 tokens (the temperature ladder's gates), engine version:
 * out = {sum_logprobs, avg_logprobs, entropy, score} over the first result_len tokens */
int  ss_debug_sequence_score(const int *ids, const float *plogs, int n, int result_len, float length_penalty, double out[4]);
/* beam search's candidate assignment of one step (whisper_full, BEAM_SEARCH branch; the function the beam decoders run, host only):
 * candidate c = len[c] token ids at ids[c * max_len], its sum_logprobs_all and the decoder it came from, in the order the decoders
 * produced them; live[j] = decoder j still running; i = index of the token being sampled.  out[j] = index of the candidate decoder j
 * continues with, -1 for a decoder that is not live. */
int  ss_debug_beam_assign(const int *ids, const int *len, int max_len, const double *sums, const int *decoder_idx, int n_cands,
                          const int *live, int n_cur, int i, int *out);
/* parses a ggml .bin exactly like ss_engine_open (same errors) without touching a GPU */
int  ss_model_probe(const char *ggml_path, int hparams_out[11], int64_t *arena_bytes, uint64_t *arena_fnv1a,
                    int *token_eot, int *token_beg, int *n_vocab_strings);

/* ---- stage-level entry points (parity tests and roofline measurement) ---- */
/* PCM -> log-mel on the device; copies [n_mels][n_len] f32 to host `mel_out` (may be NULL to only
 * time it).  n_len/n_len_org as whisper.cpp defines them. */
int  ss_log_mel(ss_engine *e, ss_state *s, const float *pcm, size_t n_samples, float *mel_out,
                size_t mel_cap, int *n_len, int *n_len_org);
/* encoder + cross-KV for the window starting at mel frame `seek` of the last ss_log_mel on this
 * state; copies [n_audio_ctx][n_audio_state] f32 to enc_out if non-NULL. */
int  ss_encode(ss_engine *e, ss_state *s, int seek, float *enc_out, size_t enc_cap);
/* teacher-forced decoder: feeds n tokens at positions n_past.., returns the last token's raw
 * logits [n_vocab] into logits_out (host). */
int  ss_decode(ss_engine *e, ss_state *s, const int *tokens, int n, int n_past, float *logits_out);
/* D[M][N] = A[M][K] * B[N][K]^T, f16 in, f32 out, through the tcgen05/TMA GEMM (host pointers). */
int  ss_debug_gemm(int device, const uint16_t *a, const uint16_t *b, float *d, int M, int N, int K,
                   int b_mn_major);

#ifdef __cplusplus
}
#endif
#endif
