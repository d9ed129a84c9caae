/*
 * libsupertonic_cuda — C ABI of the B200-native (sm_100a) Supertonic synthesis forward pass.
 *
 * Drop-in boundary for the reference's four ONNX Runtime sessions
 * (zhoubin-me/supertonic cpp/helper.cpp: Ort::Session::Run at :519-523 duration_predictor,
 * :552-556 text_encoder, :643-647 vector_estimator, :668-672 vocoder; sessions created at
 * :784-795). Plain pointers and sizes only; every pointer is HOST memory owned and kept alive
 * by the caller for the duration of the call (same ownership as the Ort::Value inputs the
 * reference builds at cpp/helper.cpp:495-509, 574-580, 604-618); outputs are written into
 * caller-provided host buffers. All tensors are dense row-major with the reference's layouts.
 *
 * Model files: stc_create reads the four .onnx graphs of an asset directory. It takes the layer plan from the `stc_arch` metadata the
 * surrogate generator writes, or derives it from the graph's NODES (ConvNeXt / attention / time-conditioning patterns, see
 * stc_derive_arch) — so a released export loads if it is built from those patterns and is rejected with the list of unexplained
 * nodes otherwise. The released Supertonic assets were never available to this project: everything measured and tested so far ran on
 * labelled SURROGATE graphs of the hypothesised architecture.
 *
 * Errors: every function returns STC_OK (0) or a negative status; the message is available via
 * stc_last_error(). There is NO CPU fallback: without a CUDA device stc_create fails.
 *
 * Threading: one handle per GPU; calls on one handle are serialised by the caller (the reference
 * is single-caller too: globals at cpp/helper.cpp:18-19, statics at :931-934). Distinct handles
 * are fully concurrent.
 */
#ifndef SUPERTONIC_CUDA_H_
#define SUPERTONIC_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stc_handle stc_handle;

enum {
    STC_OK = 0,
    STC_ERR_INVALID = -1,   /* bad argument / shape mismatch (reference: std::runtime_error) */
    STC_ERR_IO = -2,        /* asset file missing or malformed (reference: cpp/helper.cpp:805, 1057) */
    STC_ERR_CUDA = -3,      /* CUDA runtime failure or no usable device */
    STC_ERR_UNSUPPORTED = -4,
    STC_ERR_CAPACITY = -5   /* caller buffer too small */
};

/* GEMM arithmetic for the dense contractions. */
enum {
    STC_PREC_DEFAULT = 0,   /* = STC_PREC_BF16X3 unless env STC_PRECISION says otherwise */
    STC_PREC_BF16X3 = 1,    /* tcgen05 kind::f16, 3-term split-bf16 products, fp32 accumulate in TMEM */
    STC_PREC_FP32_SIMT = 2  /* CUDA-core fp32 FMA GEMM (debug cross-check; not the product path) */
};

/* Model geometry, read from onnx_dir/tts.json (reference loadCfgs, cpp/helper.cpp:801-818) and
 * from the graphs' I/O signatures. */
typedef struct stc_config {
    int32_t sample_rate;            /* ae.sample_rate */
    int32_t base_chunk_size;        /* ae.base_chunk_size */
    int32_t chunk_compress_factor;  /* ttl.chunk_compress_factor */
    int32_t latent_dim;             /* ttl.latent_dim */
    int32_t latent_channels;        /* D = latent_dim * chunk_compress_factor (cpp/helper.cpp:439) */
    int32_t chunk_size;             /* cs = base_chunk_size * chunk_compress_factor (:437) */
    int32_t text_emb_channels;      /* C of text_emb[B,C,T] */
    int32_t style_ttl_tokens, style_ttl_dim;   /* style_ttl[B,S,Cs] */
    int32_t style_dp_tokens, style_dp_dim;     /* style_dp[B,e1,e2] */
    int32_t vocab_size;
} stc_config;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Replaces loadOnnxAll (cpp/helper.cpp:784-795) + the use_gpu branch of loadTextToSpeech
 * (:903-937, which throws today): parses the four .onnx files in `onnx_dir`, uploads their
 * initializers to `device`, builds the kernel plan. */
int stc_create(const char* onnx_dir, int device, int precision, stc_handle** out);
void stc_destroy(stc_handle* h);
/* Thread-local message of the last failing call (valid with h == NULL for stc_create failures). */
const char* stc_last_error(const stc_handle* h);
int stc_get_config(const stc_handle* h, stc_config* out);

/* Shape check of the style tensors a caller is about to pass for a batch of B utterances: style_ttl must be [B, S, Cs] and
 * style_dp [B, e1, e2] with the dims of stc_config (ONNX Runtime raises "Got invalid dimensions for input" at the same
 * boundary, cpp/helper.cpp:519, 552, 643). The entry points below take bare pointers and trust the caller, so every
 * wrapper that builds these tensors from client-selectable voice-style files (cpp/helper.cpp:829-897) must call this first.
 * Either shape pointer may be NULL (not checked). Returns STC_ERR_INVALID with a message naming the offending dims. */
int stc_validate_style(const stc_handle* h, int B, const int64_t style_ttl_shape[3], const int64_t style_dp_shape[3]);

/* ---- parity layer: 1:1 with the four Session::Run calls, host pointers in/out ---------------- */

/* duration_predictor.onnx — cpp/helper.cpp:512-526. text_ids[B,T] int64, style_dp[B,e1,e2],
 * text_mask[B,1,T] -> duration[B] seconds (before the /speed of :529-531). */
int stc_duration(stc_handle* h, const int64_t* text_ids, const float* style_dp, const float* text_mask,
                 int B, int T, float* duration_out);

/* text_encoder.onnx — cpp/helper.cpp:545-556, 583-587. -> text_emb[B,C,T]; shape_out = {B,C,T}. */
int stc_text_encode(stc_handle* h, const int64_t* text_ids, const float* style_ttl, const float* text_mask,
                    int B, int T, float* text_emb_out, int64_t shape_out[3]);

/* vector_estimator.onnx — cpp/helper.cpp:620-658: ONE Euler step; the output is the updated latent.
 * noisy_latent[B,D,L], text_emb[B,C,T], style_ttl[B,S,Cs], text_mask[B,1,T], latent_mask[B,1,L],
 * total_step[B], current_step[B] (float32, as the reference passes them :573-618) -> denoised[B,D,L]. */
int stc_vector_step(stc_handle* h, const float* noisy_latent, const float* text_emb, const float* style_ttl,
                    const float* text_mask, const float* latent_mask, const float* total_step,
                    const float* current_step, int B, int L, int T, float* denoised_out);

/* vocoder.onnx — cpp/helper.cpp:662-682. latent[B,D,L] -> wav[B, L*cs]. */
int stc_vocode(stc_handle* h, const float* latent, int B, int L, float* wav_out);

/* ---- fast layer: the whole _infer body (cpp/helper.cpp:488-682) device-resident --------------- */

/* DP -> /speed -> float32 length math of sampleNoisyLatent (:424-467) -> TE -> noise*mask ->
 * total_step x VE -> vocoder, one device->host sync for the data-dependent latent length.
 *
 *   noise: NULL -> device Philox N(0,1) keyed by (seed, utterance index, channel, frame);
 *          else host float32 [B][D][noise_ld] with noise_ld >= the L the call will compute
 *          (the deterministic-noise hook the reference lacks: its RNG is unseedable, :442-443).
 *   wav_out: [B][wav_ld] floats, wav_ld >= L*cs else STC_ERR_CAPACITY (L_out is still written, so
 *          the caller can retry). Row b holds the full untrimmed L*cs samples like result.wav (:679).
 *   duration_out[B]: seconds after /speed (:529-531). wav_lengths_out[B] (optional): (int64)(d*sr) (:434).
 *   latent_out (optional): final latent [B][D][L] (for parity tests).
 */
int stc_synthesize(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                   const float* style_dp, int B, int T, int total_step, float speed,
                   const float* noise, int64_t noise_ld, uint64_t seed,
                   float* wav_out, int64_t wav_ld, float* duration_out, int64_t* wav_lengths_out,
                   int64_t* L_out, float* latent_out);

/* Same work, but inputs already resident in device memory and outputs left there (bench `value`
 * leg: no host<->device traffic in the timed region). Pointers are DEVICE pointers except L_out.
 * wav_dev must hold B*wav_ld floats. */
int stc_synthesize_device(stc_handle* h, const int64_t* text_ids_dev, const float* text_mask_dev,
                          const float* style_ttl_dev, const float* style_dp_dev, int B, int T,
                          int total_step, float speed, uint64_t seed, float* wav_dev, int64_t wav_ld,
                          float* duration_dev, int64_t* L_out);

/* Throughput variant: the latent side runs on PACKED rows (utterance b owns ceil(wav_len_b/cs) frames, the integer
 * formula of getLatentMask, cpp/helper.cpp:767) and the text side on packed tokens (the mask rows must be the prefix masks
 * of lengthToMask, cpp/helper.cpp:740-757) — no padded frames or tokens are computed. Results on every utterance's valid
 * region equal the rectangle variant (utterances are independent; tests: batch-composition invariance).
 *   wav_out: packed floats, utterance b at [wav_offsets_out[b], wav_offsets_out[b+1]) = frames_b*cs samples, of which the
 *            first wav_lengths_out[b] are the utterance (what cpp/example_onnx.cpp:104-109 keeps); wav_cap in floats.
 *   noise (optional) as in stc_synthesize, indexed [b][d][frame]; latent_out (optional): packed [sum frames][D] rows. */
int stc_synthesize_packed(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                          const float* style_dp, int B, int T, int total_step, float speed,
                          const float* noise, int64_t noise_ld, uint64_t seed,
                          float* wav_out, int64_t wav_cap, int64_t* wav_offsets_out, float* duration_out,
                          int64_t* wav_lengths_out, float* latent_out);
/* Asynchronous flavour of stc_synthesize_packed for request streams: returns as soon as the result SIZES are known
 * (duration_out / wav_lengths_out / wav_offsets_out are valid on return) and all device work including the device->host copy of
 * the waveform has been enqueued. wav_out_pinned must be page-locked (stc_pinned_alloc) and is valid after stc_wait(h). The next
 * stc_synthesize_packed_async on the same handle may be issued BEFORE stc_wait: its computation overlaps this call's copy (two
 * alternating device result buffers; a third call blocks until the first has landed). Any synchronous entry point drains
 * outstanding asynchronous calls first. Noise: device Philox keyed by `seed`. */
int stc_synthesize_packed_async(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                                const float* style_dp, int B, int T, int total_step, float speed, uint64_t seed,
                                float* wav_out_pinned, int64_t wav_cap, int64_t* wav_offsets_out, float* duration_out,
                                int64_t* wav_lengths_out);
/* Output options of the packed entry point below (long-form pipeline):
 *   pcm16       1: samples are int16, quantised on the device exactly like writeWavFile — (int16_t)(max(-1, min(1, x)) * 32767),
 *               the cast truncating toward zero (cpp/helper.cpp:985-988) — which halves the device->host traffic; 0: float32.
 *   gap_samples zeros written after every utterance but the last: TextToSpeech::call joins the chunks of a long text with
 *               (int)(silence_duration * sample_rate) zeros between the UNTRIMMED chunk waveforms (cpp/helper.cpp:706-714); with the
 *               chunks of one text as the batch, `out` is that joined waveform.
 *   noise_index optional HOST int64[B] (NULL: b): utterance b draws the device noise stream (seed, noise_index[b]) instead of (seed, b),
 *               so a request batch split over several calls or GPUs gets, utterance by utterance, the noise one call would give it. */
typedef struct stc_out_opts {
    int32_t pcm16;
    int64_t gap_samples;
    const int64_t* noise_index;
} stc_out_opts;

/* stc_synthesize_packed with output options. `out` holds out_cap ELEMENTS (float or int16_t per opts->pcm16; opts == NULL: plain
 * float32). Utterance b's frames_b * cs samples start at wav_offsets_out[b]; wav_offsets_out[B] is the total element count (gaps
 * included, none after the last utterance). async_copy != 0: as stc_synthesize_packed_async (`out` page-locked, valid after
 * stc_wait, no injected noise). */
int stc_synthesize_packed_ex(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                             const float* style_dp, int B, int T, int total_step, float speed,
                             const float* noise, int64_t noise_ld, uint64_t seed, const stc_out_opts* opts,
                             void* out, int64_t out_cap, int64_t* wav_offsets_out, float* duration_out,
                             int64_t* wav_lengths_out, int async_copy);

/* Waits for every outstanding asynchronous call of the handle; reports their errors. */
int stc_wait(stc_handle* h);

/* text_lens (HOST int32[B], optional): token count of every utterance (= sum of its mask row); lets the text side run on
 * packed rows too. NULL -> the text side is computed on the padded [B,T] rectangle. */
int stc_synthesize_packed_device(stc_handle* h, const int64_t* text_ids_dev, const float* text_mask_dev,
                                 const float* style_ttl_dev, const float* style_dp_dev, const int32_t* text_lens,
                                 int B, int T, int total_step, float speed, uint64_t seed, float* wav_dev,
                                 int64_t wav_cap, int64_t* wav_offsets_out, float* duration_dev);

/* ---- page-locked host buffers (optional) ------------------------------------------------------ */

/* Any host pointer is accepted by the entry points above; buffers obtained here are page-locked, so the device->host copy
 * of the waveform (the bulk of the traffic: 4 bytes x 44100 per audio-second) runs at PCIe speed instead of being staged
 * through the driver's bounce buffers. Needs a CUDA device. */
int stc_pinned_alloc(size_t bytes, void** out);
void stc_pinned_free(void* p);

/* ---- host front-end (kept on the host, semantics of the C++ reference) ----------------------- */

/* UnicodeProcessor::call (cpp/helper.cpp:355-390) for n texts. Two-pass: call with text_ids == NULL to
 * get T_out (max token count), then with buffers text_ids[n*T], text_mask[n*T]. langs: "en" ... */
int stc_text_to_ids(stc_handle* h, const char* const* texts, const char* const* langs, int n,
                    int64_t* text_ids, float* text_mask, int64_t T_cap, int64_t* T_out);

/* The same front-end without a GPU handle (host-only: usable on machines with no CUDA device). */
typedef struct stc_frontend stc_frontend;
int stc_frontend_open(const char* unicode_indexer_json, stc_frontend** out);   /* loadTextProcessor, cpp/helper.cpp:820-823 */
void stc_frontend_close(stc_frontend* fe);
int stc_frontend_text_to_ids(const stc_frontend* fe, const char* const* texts, const char* const* langs, int n,
                             int64_t* text_ids, float* text_mask, int64_t T_cap, int64_t* T_out);

/* chunkText (cpp/helper.cpp:1117-1186). Writes chunk byte offsets/lengths into `text` order copies:
 * out_buf receives the chunks back to back, NUL-separated; returns count in n_out. */
int stc_chunk_text(const char* text, int max_len, char* out_buf, size_t out_cap, size_t* out_len, int* n_out);

/* ---- introspection -------------------------------------------------------------------------- */

/* Kernels launched by this handle since creation (bench `gpu_launches`). */
uint64_t stc_launch_count(const stc_handle* h);
/* Which variant each size-dependent kernel dispatcher chose, as "name=count\n" lines (launches issued or captured since the handle
 * was created): gemm_bn64 / gemm_bn128 / gemm_bn256 / gemm2_bf16x3 / gemm2_f16 / gemm_f16_bn*, mlp_*, dwconv_ln_slide[_ring] / _tile /
 * _vec / _generic, attention_tc[_small] / attention_simt. Lets a parity test assert that the kernels a benchmark-size input reaches
 * are the ones it compared with the oracle. *need = bytes required (incl. NUL); STC_ERR_CAPACITY if cap is smaller. */
int stc_kernel_variants(const stc_handle* h, char* buf, size_t cap, size_t* need);
/* Use CUDA graphs for the fast layer (default 1). */
int stc_set_graphs(stc_handle* h, int enabled);
/* cudaStream_t the handle launches on (as void*), for callers that time with CUDA events. */
void* stc_stream(stc_handle* h);
/* Device-side timing of the most recent stc_synthesize* call: milliseconds per stage
 * {dp, te, ve_total, vocoder, whole}. Requires stc_set_profile(h, >=1) (adds event records). */
int stc_set_profile(stc_handle* h, int level);   /* 0 off, 1 stage events, 2 + per-kernel events (bench roofline leg) */
/* Per-kernel-class totals of the most recent stc_synthesize* call at profile level 2.
 * cls: 0 = tcgen05 GEMM (split-bf16 operands), 1 = depthwise-conv+LayerNorm, 2 = attention core, 3 = fused ConvNeXt MLP (with its
 * reduce kernel), 4 = tcgen05 GEMM with single-pass fp16 operands (the vocoder projections), 5 = depthwise-conv+LayerNorm launches
 * of >= 32 MB (the vocoder's, HBM-resident; class 1 then holds the L2-resident ones).
 * out = {milliseconds (CUDA events around each launch), algorithmic FLOPs, algorithmic bytes, launches}. */
int stc_kernel_profile(const stc_handle* h, int cls, double out[4]);
int stc_last_stage_ms(const stc_handle* h, float out[5]);

/* Tuning / self-check of the tcgen05 GEMM: out[M,N] = epi(A[M,K] W[K,N]) on seeded random data with a forced tile width
 * `bn` (64/128/256; 0 = the library's own choice) and cluster shape cm x cn; epilogue 0 = bias, 1 = bias+GELU -> split
 * bf16 operand, 2 = bias, layer-scale, residual (in place), row mask. Returns the mean device time of `iters` back-to-back
 * launches (L2-warm) and the max-abs difference to the CUDA-core fp32 GEMM of the same operands. */
int stc_debug_gemm(stc_handle* h, int M, int N, int K, int bn, int cm, int cn, int epilogue, int iters,
                   float* ms_per_iter, float* max_abs_err);

/* Same for the fused ConvNeXt MLP (C = 256, H = 1024) against the two-GEMM form on M rows of seeded data: mean device time of
 * each (CUDA-graph replay of `iters` launches) and the max-abs difference of the updated residual stream. */
int stc_debug_mlp(stc_handle* h, int M, int iters, float* ms_fused, float* ms_unfused, float* max_abs_err);

/* Same for depthwise conv1d + LayerNorm (the `Conv(group=C)` -> `LayerNormalization` pair of every ConvNeXt block): C in
 * {128, 256, 512}, K in {5, 7}, `B` ragged sequences packed into `rows` rows (the last rows/16 are bucket padding, one sequence
 * is empty). Register sliding-window kernel (chains of `rt` rows; 0 = the library's choice) against the shared-memory tiled
 * kernel: mean device time of each writing split-bf16 operands, and the max-abs difference of their fp32 outputs. */
int stc_debug_dwconv(stc_handle* h, int rows, int C, int K, int dil, int causal, int B, int rt, int iters, float* ms_slide,
                     float* ms_tile, float* max_abs_diff);

/* The layer plan derived from the NODES of one graph file (kind: "duration_predictor" | "text_encoder" | "vector_estimator" |
 * "vocoder") as a JSON string — what stc_create uses for graphs that carry no `stc_arch` metadata (a released export; the reference
 * loads whatever is on disk, cpp/helper.cpp:776-795). Host only. STC_ERR_UNSUPPORTED with the list of unexplained nodes when a
 * pattern is not recognised; STC_ERR_CAPACITY (and *need) when buf is too small. */
int stc_derive_arch(const char* onnx_path, const char* kind, char* buf, size_t cap, size_t* need);

/* The device PCM16 quantiser of stc_out_opts.pcm16 on n caller-provided float samples (known-answer tests against the reference's
 * writeWavFile, cpp/helper.cpp:985-988). */
int stc_debug_pcm16(stc_handle* h, const float* samples, int64_t n, int16_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SUPERTONIC_CUDA_H_ */
