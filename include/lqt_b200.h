/* lqt_b200.h -- C-ABI of the B200-native Qwen3-TTS hot path.
 *
 * Drop-in boundary: these entry points are exactly what the reference's host code
 * (leaxer-ai/leaxer-qwen3-tts, src/tts_onnx.cpp) would bind in place of its seven (+1)
 * Ort::Session::Run call sites. Each function cites the reference interface it replaces.
 * Plain C types only; caller-owned HOST buffers; int status (0 = ok, non-zero = error, message
 * via lqt_last_error). One handle = one GPU = one host thread at a time (same threading contract
 * as the reference's TTSEngine, which is not re-entrant: src/tts_onnx.h:167-190).
 *
 * There is no CPU fallback: lqt_create fails if no sm_100 device is present.
 */
#ifndef LQT_B200_H
#define LQT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lqt_engine lqt_engine;

/* src/tts_onnx.h:99-105 SamplingParams, plus the seeded-Philox extension (north_star):
 * key = (seed, utterance_id), counter = (frame, codebook 0..15). greedy != 0 -> argmax with
 * lowest-index tie break (== reference `--top-k 1` barring exact ties). */
typedef struct lqt_sampling {
    float    temperature;
    float    top_p;
    int32_t  top_k;
    int32_t  max_new_tokens;
    uint32_t seed;
    uint32_t utterance_id;
    int32_t  greedy;
} lqt_sampling;

typedef struct lqt_info {
    int32_t hidden, layers, heads, kv_heads, head_dim, vocab;      /* src/tts_onnx.h:31-35 */
    int32_t cp_vocab, cp_steps;                                    /* src/tts_onnx.h:36-37 */
    int32_t samples_per_frame, sample_rate;                        /* 1920, 24000 (tts_onnx.h:69) */
    int32_t has_speaker_encoder;                                   /* tts_onnx.h:161 */
    int32_t max_pos, num_sms;
} lqt_info;

typedef struct lqt_stats {
    uint64_t kernel_launches;     /* kernels of this library launched since create/reset */
    uint64_t graph_launches;
    float    last_generate_ms;    /* CUDA-event time of the last device frame loop */
    float    last_vocoder_ms;     /* CUDA-event time of the last vocoder pass */
    float    last_prefill_ms;
    int32_t  last_frames;
    float    last_total_ms;       /* lqt_synthesize_tokens: CUDA-event time prompt build -> last vocoder kernel */
    float    first_audio_ms;      /* lqt_synthesize_tokens: CUDA-event time prompt build -> the first chunk of PCM (320 ms) copied into
                                     the caller's buffer (chunked vocoding on a second stream); = last_total_ms when chunking is off */
    int32_t  frame_impl_active;   /* the frame loop lqt_synthesize_tokens runs on this handle: LQT_FRAME_PERSISTENT, _BATCHED or _GRAPH */
    int32_t  cooperative_launch;  /* 1 = the persistent kernel is launched with the cooperative attribute (co-residency guaranteed);
                                     0 = refused by the driver/profiler: the second-stream first-audio overlap is then switched off */
} lqt_stats;

/* Engine options (lqt_create_ex). kv_dtype: LQT_KV_BF16 = paged bf16 talker KV cache (default,
 * north_star); LQT_KV_F32 = "parity mode": fp32 pages, no rounding point in the talker, for
 * free-running token-exact comparison against the fp32 oracle over long utterances. */
enum { LQT_KV_BF16 = 0, LQT_KV_F32 = 1 };
/* frame_impl: LQT_FRAME_PERSISTENT (default) = loops A+B run inside one persistent cooperative kernel
 * whose producer warp streams the weights with TMA bulk copies; LQT_FRAME_GRAPH = the round-1 v1
 * schedule (a CUDA graph of ~577 per-op kernels per frame), kept only for A/B measurements. */
/* LQT_FRAME_AUTO (what lqt_create uses) = the persistent kernel where the model shape fits it, else LQT_FRAME_BATCHED: the
 * batched path's CUDA graph of TMA-fed tcgen05 GEMM kernels with a single KV slot (any hidden / MLP width that is a multiple
 * of 64 / 128, e.g. the 1.7B talker). The choice is reported in lqt_stats.frame_impl_active. An explicit
 * LQT_FRAME_PERSISTENT is strict: lqt_create_ex FAILS with the reason instead of falling back. LQT_FRAME_GRAPH is the
 * round-1 graph of per-op GEMV kernels, kept for A/B only; on a LQT_FRAME_BATCHED handle it still backs the per-graph
 * parity entry points (lqt_talker_prefill/decode, lqt_code_predictor, lqt_generate). */
enum { LQT_FRAME_PERSISTENT = 0, LQT_FRAME_GRAPH = 1, LQT_FRAME_AUTO = 2, LQT_FRAME_BATCHED = 3 };
typedef struct lqt_options {
    int32_t kv_dtype;
    int32_t n_slots;      /* KV slots (utterances resident at once); 0 = default (2) */
    int32_t frame_impl;
} lqt_options;

/* src/tts_onnx.cpp:84-130 (TTSEngine ctor) + :134-232 (load_model): loads the 7 required graph
 * files (+ optional speaker_encoder) from model_dir. On failure returns non-zero, *out = NULL and
 * lqt_create_error() describes why (the reference sets error_msg_ and leaves ready_ = false). The message is per calling thread:
 * engines may be created from several threads (one per device) at once. */
int lqt_create(const char* model_dir, int device_id, lqt_engine** out);
int lqt_create_ex(const char* model_dir, int device_id, const lqt_options* opt, lqt_engine** out);
const char* lqt_create_error(void);
/* Header-only validation of one .lqw weight file (no GPU needed): 0 = well formed; otherwise non-zero and `err` (capacity
 * err_cap, always NUL-terminated) says why. lqt_create runs the same checks on every file before any device allocation, then
 * verifies the dtype and shape of every tensor against the model spec. */
int lqt_check_model_file(const char* path, char* err, int32_t err_cap);
void lqt_destroy(lqt_engine* h);
const char* lqt_last_error(lqt_engine* h);
int lqt_get_info(lqt_engine* h, lqt_info* out);
int lqt_get_stats(lqt_engine* h, lqt_stats* out);
int lqt_reset_stats(lqt_engine* h);

/* ---- per-graph entry points: one per Ort::Session (parity-test surface) ---------------------- */

/* text_project.onnx  src/tts_onnx.cpp:545-559: input_ids i64 [1,S] -> embeds f32 [1,S,H] */
int lqt_text_project(lqt_engine* h, const int64_t* ids, int32_t S, float* out);
/* codec_embed.onnx  :561-590: input_ids i64 [1,N] -> embeds f32 [1,N,H] */
int lqt_codec_embed(lqt_engine* h, const int64_t* ids, int32_t N, float* out);
/* code_predictor_embed.onnx  :592-613: (input_ids [1,1], generation_step [1]) -> embeds [1,1,H] */
int lqt_code_predictor_embed(lqt_engine* h, int64_t id, int64_t generation_step, float* out);
/* talker_prefill.onnx  :615-665: inputs_embeds [1,P,H] (attention_mask is all ones, :791) ->
 * logits of the LAST position [V] (the only row the host reads, :797-798), last_hidden [H].
 * present_key/value stay on the device in the paged bf16 KV cache of `slot` (reset first). */
int lqt_talker_prefill(lqt_engine* h, int32_t slot, const float* embeds, int32_t P,
                       float* logits_last, float* last_hidden);
/* talker_decode.onnx  :667-732: inputs_embeds [1,1,H] + past KV of `slot` -> logits [V],
 * last_hidden [H]; KV of `slot` grows by one position. */
int lqt_talker_decode(lqt_engine* h, int32_t slot, const float* embed,
                      float* logits, float* last_hidden);
int lqt_kv_reset(lqt_engine* h, int32_t slot);
int lqt_kv_len(lqt_engine* h, int32_t slot);
/* code_predictor.onnx  :734-757: inputs_embeds [1,L,H] (L = 2..16), generation_step [1] ->
 * logits [cp_vocab] of head `generation_step` at the last position. */
int lqt_code_predictor(lqt_engine* h, const float* embeds, int32_t L, int64_t generation_step,
                       float* logits);
/* tokenizer12hz_decode.onnx  :759-776: audio_codes i64 [1,T,16] -> audio_values f32, lengths[0].
 * `audio` must hold T * samples_per_frame floats. */
int lqt_vocoder_decode(lqt_engine* h, const int64_t* codes, int32_t T, float* audio,
                       int64_t* length);
/* Streaming form of the same graph (the reference vocodes once at the end, :430): the codes of one utterance are
 * handed over in chunks of any size, in order; every layer with left context (pre_conv, the 72-position attention window, the
 * depthwise and k = 7 dilated convolutions, the transposed convolutions) keeps its tail between calls, so the concatenated
 * chunks are BIT-IDENTICAL to lqt_vocoder_decode of the whole sequence and nothing is decoded twice. reset starts a new
 * utterance. */
int lqt_vocoder_stream_reset(lqt_engine* h);
int lqt_vocoder_stream_chunk(lqt_engine* h, const int64_t* codes, int32_t T, float* audio, int64_t* length);
/* speaker_encoder.onnx  :367-403: log-mel f32 [1,frames,128] -> embedding [H] */
int lqt_speaker_encoder(lqt_engine* h, const float* mel_t, int32_t frames, float* out);
/* Clone front end on the device: extract_speaker_embedding's log-mel (src/tts_onnx.cpp:331-359 with
 * src/io/mel.cpp:132-236: 1024-point Hann STFT, hop 256, no centre padding, 128 HTK-mel triangles, log(e + 1e-10)) as one
 * kernel. `audio` is 24 kHz mono f32 (host). lqt_log_mel returns the reference layout [128][frames]
 * (mel_out nullable, capacity 128 * ((n-1024)/256+1)); lqt_speaker_embed_audio chains it with speaker_encoder.onnx's stand-in
 * (:367-403) without the mel leaving the device. */
int lqt_log_mel(lqt_engine* h, const float* audio, int64_t n_samples, float* mel_out, int32_t* frames);
int lqt_speaker_embed_audio(lqt_engine* h, const float* audio, int64_t n_samples, float* out);
/* sample_token  :878-905 (+ special-token mask :803-807 when mask_codec_specials != 0) on the
 * device sampler, Philox counter (frame, codebook). */
int lqt_sample(lqt_engine* h, const float* logits, int32_t V, const lqt_sampling* sp,
               uint32_t frame, uint32_t codebook, int32_t mask_codec_specials, int64_t* id);

/* ---- fast path: the two nested loops on the device ------------------------------------------- */

/* generate_codes + predict_subcodes  src/tts_onnx.cpp:782-872.
 * prompt [P,H], trailing_text_hidden [trailing_len,H], tts_pad_embed [H] as built by
 * build_prompt_embeddings (:442-539). codes_out [max_new_tokens,16] i64, *n_frames = frames
 * produced (stops at CODEC_EOS). forced_codes (nullable, [n_forced,16]) = teacher forcing: the
 * sampled tokens are replaced by these after sampling (parity triage); logits_trace (nullable,
 * [max_new_tokens,16,trace_stride] f32, trace_stride >= vocab) receives every logits vector. */
int lqt_generate(lqt_engine* h, int32_t slot, const float* prompt, int32_t P,
                 const float* trailing, int32_t trailing_len, const float* tts_pad,
                 const lqt_sampling* sp, const int64_t* forced_codes, int32_t n_forced,
                 int64_t* codes_out, int32_t* n_frames, float* logits_trace, int32_t trace_stride);

/* synthesize_tokens  src/tts_onnx.cpp:405-436 (speaker_embed != NULL: the clone variant
 * :299-313): prompt assembly (:442-539) + generate + vocoder, everything on the device.
 * lang_codec_id = language_to_codec_id(lang) (0 = Auto, tts_onnx.h:230-238).
 * audio_out capacity in floats; *n_samples = samples written (0 = empty result, which the
 * reference also returns when the first code is EOS, :418). codes_out nullable [max_new,16]. */
int lqt_synthesize_tokens(lqt_engine* h, const int64_t* token_ids, int32_t n_ids,
                          int32_t lang_codec_id, const float* speaker_embed,
                          const lqt_sampling* sp, float* audio_out, int64_t audio_capacity,
                          int64_t* n_samples, int64_t* codes_out, int32_t* n_frames);

/* ---- batched path: many utterances per GPU (BASELINE configs[3], [4]) --------------------------- *
 * The reference synthesises exactly one utterance per call (src/tts_onnx.cpp:405-436, batch dim 1 at :547, 618, 672-674);
 * a serving host would loop over requests. lqt_synthesize_batch takes the whole list: up to `max_concurrent` utterances
 * run in lockstep KV slots (continuous batching: a finished utterance's slot and KV pages go to the next request), every
 * matrix product is one TMA-fed tcgen05 GEMM over all slots, each utterance keeps its own Philox key
 * (seed, utterance_id) so its codes do not depend on the batch composition. Per request the semantics are those of
 * lqt_synthesize_tokens. planes: 3 = fp32-exact activations (three bf16 planes, token-exact against the oracle),
 * 2 = 16-bit mantissa, 1 = plain bf16 activations (logits within the 2e-2 bound). */
typedef struct lqt_batch_request {
    const int64_t* token_ids; int32_t n_ids;     /* [IM_START, ASSISTANT, TTS_BOS, text..., TTS_EOS, IM_END] */
    int32_t  lang_codec_id;                      /* 0 = auto */
    const float* speaker_embed;                  /* nullable [hidden] */
    uint32_t utterance_id;                       /* Philox key word 1 */
    int32_t  max_new_tokens;
    const int64_t* forced_codes; int32_t n_forced;   /* nullable teacher forcing, [n_forced,16] */
    float*   audio_out; int64_t audio_capacity; int64_t* n_samples;   /* nullable: skip the vocoder */
    int64_t* codes_out; int32_t* n_frames;       /* [max_new_tokens,16] */
    float*   logits_trace;                       /* nullable [max_new_tokens,16,max(vocab,cp_vocab)] */
} lqt_batch_request;
typedef struct lqt_batch_options {
    int32_t max_concurrent;   /* KV slots = utterances in flight; 0 = all requests at once */
    int32_t planes;           /* 0 = 3 */
    int32_t poll_frames;      /* frames between completion checks; 0 = 8 */
} lqt_batch_options;
int lqt_synthesize_batch(lqt_engine* h, const lqt_batch_request* reqs, int32_t n_reqs, const lqt_sampling* sp,
                         const lqt_batch_options* opt);
/* Parity surface of the batched path's GEMM kernel alone (tc_gemm.cuh): out[b][n] = sum_k bf16(W[n][k]) * x[b][k];
 * W [N,K] and x [B,K] are fp32 host arrays (W is rounded to bf16 on the device, x split into `planes` planes);
 * splits = 0 picks the split-K factor like the engine does. */
int lqt_debug_tc_gemm(lqt_engine* h, const float* W, const float* x, int32_t N, int32_t K, int32_t B, int32_t planes,
                      int32_t splits, float* out);

/* Streaming form of lqt_synthesize_tokens (the reference returns the whole waveform at the end, :430-436): same
 * arguments and result, plus a callback that is invoked on the calling thread, in order, for every chunk of PCM (the
 * first one 4 frames = 320 ms, then 25 frames = 2 s: $LQT_FIRST_CHUNK / $LQT_STREAM_CHUNK) as soon as it has been vocoded and copied into audio_out -- while the frame kernel is still
 * generating the rest of the utterance. pcm points into audio_out at first_sample. The chunks are bit-identical to the one-shot
 * result (streaming vocoder with carried state). */
typedef void (*lqt_audio_callback)(void* user, const float* pcm, int64_t first_sample, int64_t n_samples);
int lqt_synthesize_stream(lqt_engine* h, const int64_t* token_ids, int32_t n_ids, int32_t lang_codec_id, const float* speaker_embed,
                          const lqt_sampling* sp, float* audio_out, int64_t audio_capacity, int64_t* n_samples,
                          int64_t* codes_out, int32_t* n_frames, lqt_audio_callback on_audio, void* user);

/* build_prompt_embeddings  :442-539 on the device; outputs to host for parity tests.
 * prompt_out [10,H] capacity, *P rows used; trailing_out [n_ids,H] capacity, *trailing_len. */
int lqt_build_prompt(lqt_engine* h, const int64_t* token_ids, int32_t n_ids,
                     int32_t lang_codec_id, const float* speaker_embed,
                     float* prompt_out, int32_t* P, float* trailing_out, int32_t* trailing_len,
                     float* tts_pad_out);

/* Profiling aid (no reference counterpart): phase timeline of one CTA of the persistent frame kernel.
 * enable_entries > 0: arm recording of CTA `cta` for the following launches (buffer of that many
 * entries, overwritten by every launch); enable_entries == 0: copy the last launch's entries to
 * `out` (each = SM clock << 8 | tag, see csrc/frame_kernel.cuh) and return their count. */
int lqt_debug_timeline(lqt_engine* h, int32_t enable_entries, int32_t cta, uint64_t* out, int32_t out_cap);

/* Profiling/debug aid: values of one activation exchange buffer of the persistent frame kernel after the last launch. */
int lqt_debug_exchange(lqt_engine* h, int32_t which, float* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* LQT_B200_H */
