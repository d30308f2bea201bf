/*
 * llamagen_b200 — C-ABI of the B200-native autoregressive image-sampling engine.
 *
 * Drop-in boundary for ONE hot path of FoundationVision/LlamaGen (reference paths are
 * relative to the reference checkout):
 *   - autoregressive/models/generate.py:126-176   generate()  (prefill, KV-cached decode, CFG, sampling)
 *   - autoregressive/models/gpt.py:332-382        Transformer.forward (inference branches)
 *   - tokenizer/tokenizer_image/vq_model.py:52-55 VQModel.decode_code (lookup + conv/attn decoder)
 *   - tokenizer/tokenizer_image/vq_model.py:215-233 VectorQuantizer.forward (argmin-L2, encode side)
 *
 * The reference has no FFI layer (it is pure PyTorch); the Python shim in llamagen_b200/ mirrors its
 * Python API (GPT_models / VQ_models / generate / decode_code) and calls these entry points through
 * ctypes.  Conventions:
 *   - every pointer marked "dev" is a CUDA device pointer owned by the CALLER (PyTorch); the library
 *     never frees caller memory.  Engine-owned memory is only small repacked weights / tables.
 *   - every call returns 0 on success, <0 on error; lg_last_error() returns a thread-local message.
 *     Nothing throws or aborts.
 *   - all work is enqueued asynchronously on the given cudaStream_t (passed as void*); no hidden
 *     device synchronisation except where stated.
 *   - one engine per device per process; an engine is not re-entrant.
 */
#ifndef LLAMAGEN_B200_H
#define LLAMAGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_ABI_VERSION 1

enum { LG_DTYPE_F32 = 0, LG_DTYPE_BF16 = 1 };
enum { LG_MODEL_C2I = 0, LG_MODEL_T2I = 1 };

/* Mirrors autoregressive/models/gpt.py:23-50 (ModelArgs) — only the fields the inference path reads. */
typedef struct lg_model_cfg {
    int32_t n_layer;
    int32_t n_head;
    int32_t dim;
    int32_t ffn_dim;       /* FeedForward hidden size, gpt.py:154-159 */
    int32_t vocab_size;
    int32_t cls_token_num; /* 1 (c2i) or 120 (t2i) */
    int32_t block_size;    /* image tokens = grid*grid */
    int32_t num_classes;   /* c2i null class index == num_classes (generate.py:130) */
    int32_t caption_dim;   /* t2i feature width (2048) */
    int32_t model_type;    /* LG_MODEL_* */
    int32_t dtype;         /* LG_DTYPE_* of weights, activations and KV cache */
    float   norm_eps;      /* RMSNorm eps, gpt.py:31 */
} lg_model_cfg;

/* Sampling parameters; mirrors generate()'s **sampling_kwargs + cfg args (generate.py:57,126). */
typedef struct lg_sample_cfg {
    float    cfg_scale;     /* >1 enables classifier-free guidance (rows doubled) */
    int32_t  cfg_interval;  /* generate.py:113: after decode iteration i > cfg_interval use cond logits only; -1 = never */
    float    temperature;
    int32_t  top_k;         /* 0 = off */
    float    top_p;         /* 1.0 = off */
    int32_t  greedy;        /* 1 == sample_logits=False (argmax) */
    uint64_t seed;          /* counter-based RNG seed for the multinomial draw */
} lg_sample_cfg;

typedef struct lg_engine lg_engine;
typedef struct lg_vq lg_vq;

int         lg_version(void);
const char* lg_last_error(void);
/* Number of kernels this library has launched since load (or since lg_reset_launch_count). */
uint64_t    lg_launch_count(void);
void        lg_reset_launch_count(void);

/* ---- AR transformer engine: replaces Transformer.setup_caches/forward (gpt.py:316-382) ---------- */
int  lg_engine_create(const lg_model_cfg* cfg, int device, lg_engine** out);
void lg_engine_destroy(lg_engine* e);
/* Borrow a checkpoint tensor by its reference state_dict name, plus "freqs_cis"
 * (fp32 [cls_token_num+block_size, head_dim/2, 2], gpt.py:404-417).  dtype must equal cfg.dtype
 * except freqs_cis (always f32).  The caller keeps the tensor alive. */
int  lg_engine_bind_weight(lg_engine* e, const char* name, const void* dev_ptr,
                           const int64_t* shape, int ndim, int dtype);
/* Validate that every tensor the configured model needs is bound. */
int  lg_engine_finalize(lg_engine* e);
/* Bytes of caller-provided scratch (KV cache + activations) for `rows` sequences (rows = 2*B under CFG)
 * of up to max_seq positions (cls_token_num + new tokens). Replaces setup_caches (gpt.py:316-330). */
int  lg_engine_workspace_bytes(lg_engine* e, int rows, int max_seq, size_t* bytes);
int  lg_engine_set_workspace(lg_engine* e, void* dev_ws, size_t bytes, int rows, int max_seq);

/* Prefill over the condition (generate.py:77-86 / gpt.py:348-349).
 *  c2i: cond = dev int32 [B] class labels.      t2i: cond = dev [B, T, caption_dim] features (cfg.dtype),
 *  emb_mask = dev f32 [B, T] or NULL.  With use_cfg the engine appends the null-condition rows itself
 *  (generate.py:129-137).  Writes last-position logits f32 [rows, V] to logits_out (dev), rows = B or 2B. */
int  lg_prefill(lg_engine* e, const void* cond, const float* emb_mask, int B, int T, int use_cfg,
                float* logits_out, void* stream);
/* One decode step at absolute position `pos` (generate.py:89-102 model call): tokens dev int32 [B];
 * under CFG both halves consume the same token (torch.cat([x, x])). logits_out f32 [rows, V]. */
int  lg_decode_step(lg_engine* e, const int32_t* tokens, int B, int pos, int use_cfg,
                    float* logits_out, void* stream);
/* Iteration-level scheduling for class-conditional serving (autoregressive/serve/llm_engine.py:511 `step()`, serve/llm.py:238-266):
 * ONE decode step for sequences at DIFFERENT depths. pos_rows dev int32 [R]: position of every row (R = 2B under CFG, cond rows
 * first); a row at position 0 has just joined and takes its class embedding (cond rows: tokens[b] is the label, uncond rows: the
 * null class), any other row the embedding of tokens[b]. Writes row r's K/V at pos_rows[r] and attends over [0, pos_rows[r]].
 * logits_out dev f32 [R, V]. */
int  lg_decode_rows(lg_engine* e, const int32_t* tokens, const int32_t* pos_rows, int B, int use_cfg, float* logits_out, void* stream);
/* lg_sample with per-request RNG streams: image b draws with seed_rows[b] at token index step_rows[b] (exactly the draw a
 * batch-of-one lg_generate with that seed makes), result -> out_idx[b] and/or out_seq[b*seq_stride + step_rows[b]]. */
int  lg_sample_rows(const float* logits, int B, int V, int mix_cfg, int round_dtype, const lg_sample_cfg* sc, const uint64_t* seed_rows,
                    const int32_t* step_rows, int32_t* out_idx, int32_t* out_seq, int seq_stride, void* stream);
/* Fused CFG-mix + temperature + top-k + top-p + softmax + (argmax | multinomial): generate.py:57-66,95-97.
 * logits f32 [rows, V] (rows = 2B when mix_cfg, cond rows first); writes out_idx int32 [B] and, when
 * non-NULL, out_probs f32 [B, V] (the post-filter softmax the reference returns as `probs`).
 * round_dtype: LG_DTYPE_BF16 rounds raw logits to bf16 first (the reference's `.float()` of a bf16 head). */
int  lg_sample(const float* logits, int B, int V, int mix_cfg, int round_dtype, const lg_sample_cfg* sc,
               uint64_t step, int32_t* out_idx, float* out_probs, void* stream);
/* Whole generate(): prefill + (S-1) decode steps, tokens never leave the device.
 * out_tokens dev int32 [B, S]. dbg_logits: NULL or dev f32 [S, B, V] receiving the (CFG-mixed, unfiltered)
 * logits of every step. teacher: NULL or dev int32 [B, S] tokens to feed instead of the sampled ones. */
int  lg_generate(lg_engine* e, const void* cond, const float* emb_mask, int B, int T, int S,
                 const lg_sample_cfg* sc, int32_t* out_tokens, float* dbg_logits,
                 const int32_t* teacher, void* stream);

/* ---- VQ tokenizer: replaces VQModel.decode_code / VectorQuantizer (vq_model.py) ------------------- */
typedef struct lg_vq_cfg {
    int32_t codebook_size;
    int32_t codebook_embed_dim;
    int32_t z_channels;     /* 256 */
    int32_t ch;             /* 128 */
    int32_t num_res_blocks; /* 2 */
    int32_t n_mult;         /* len(decoder_ch_mult) */
    int32_t ch_mult[8];     /* decoder_ch_mult, vq_model.py:421-424 */
    int32_t l2_norm;        /* codebook_l2_norm */
} lg_vq_cfg;

int  lg_vq_create(const lg_vq_cfg* cfg, int device, lg_vq** out);
void lg_vq_destroy(lg_vq* v);
/* Bind by reference state_dict name ("quantize.embedding.weight", "post_quant_conv.weight",
 * "decoder.conv_in.weight", ...). All tensors are f32, NCHW-ordered conv weights as in the checkpoint. */
int  lg_vq_bind_weight(lg_vq* v, const char* name, const void* dev_ptr, const int64_t* shape, int ndim);
/* Repack conv weights to bf16 [Cout][ky][kx][Cin], L2-normalise the codebook once (vq_model.py:264).
 * Synchronises the device. */
int  lg_vq_finalize(lg_vq* v, void* stream);
int  lg_vq_workspace_bytes(lg_vq* v, int B, int grid, size_t* bytes);
/* decode_code (vq_model.py:52-55): codes dev int32 [B, grid*grid] -> out dev f32 NCHW [B,3,H,W]. */
int  lg_vq_decode(lg_vq* v, const int32_t* codes, int B, int grid, void* dev_ws, size_t ws_bytes,
                  float* out_nchw, void* stream);
/* decode_code followed by the samplers' pixel finishing (sample_c2i_ddp.py:141-143 without the optional resize):
 * clamp(127.5*x + 128, 0, 255) -> uint8, NHWC [B,H,W,3], written straight from conv_out's accumulator drain (no fp32 image in
 * HBM). Needs the tcgen05 conv path (every registry VQ model); bytes equal lg_vq_decode + lg_pixels_to_u8. */
int  lg_vq_decode_u8(lg_vq* v, const int32_t* codes, int B, int grid, void* dev_ws, size_t ws_bytes,
                     uint8_t* out_nhwc, void* stream);
/* VectorQuantizer.forward index path (vq_model.py:215-233): z dev f32 NCHW [B, e_dim, g, g] -> idx int64 [B*g*g]. */
int  lg_vq_argmin(lg_vq* v, const float* z_nchw, int B, int grid, int64_t* out_idx, void* stream);
/* VQModel.encode (vq_model.py:41-45): Encoder.forward :100-124 (Downsample :389-397) -> quant_conv -> VectorQuantizer.forward
 * :215-255 in eval mode. x dev f32 NCHW [B,3,H,W] (square, H a multiple of 2^(n_mult-1)) -> out_idx dev int64 [B*g*g];
 * optional out_quant dev f32 NCHW [B,e_dim,g,g] (the straight-through tensor z + (e[idx] - z)) and out_z (pre-quantisation
 * quant_conv output), either may be NULL. Needs the encoder.* and quant_conv.* tensors bound before lg_vq_finalize; the
 * encoder uses ch_mult of the cfg (the reference's encoder_ch_mult == decoder_ch_mult in both registry entries,
 * vq_model.py:418-422). Workspace: lg_vq_workspace_bytes(B, H / 2^(n_mult-1)). */
int  lg_vq_encode(lg_vq* v, const float* x_nchw, int B, int H, int W, void* dev_ws, size_t ws_bytes, int64_t* out_idx,
                  float* out_quant_nchw, float* out_z_nchw, void* stream);
/* Pixel finishing of the samplers (sample_c2i_ddp.py:141-143): optional F.interpolate(mode='bicubic') to out_h x out_w,
 * then clamp(127.5*x + 128, 0, 255) -> uint8, NCHW f32 in -> NHWC u8 out, one pass. */
int  lg_pixels_to_u8(const float* in_nchw, int B, int C, int H, int W, int out_h, int out_w, uint8_t* out_nhwc, void* stream);

/* ---- per-kernel-class device timing for bench.py's roofline leg --------------------------------------
 * While enabled, launches are bracketed by CUDA events on the launching stream (CUDA-graph replay is bypassed).
 * lg_profile_read synchronises the device and returns the summed duration / launch count of one class
 * (class ids: see `lg_profile_class_name`). */
/* Toggle programmatic dependent launch for subsequent launches (default: on, or LG_PDL=0/1). With it off, CUPTI
 * kernel durations are additive (no early-started kernels waiting on their dependency), which is what the
 * roofline leg of bench.py traces. */
int         lg_set_pdl(int on);
int         lg_profile_enable(int on);
int         lg_profile_reset(void);
int         lg_profile_read(int cls, double* total_ms, uint64_t* launches);
const char* lg_profile_class_name(int cls);   /* NULL past the last class */

/* Cap the number of CTAs (hence SMs) the VQ decoder's tensor-core convolutions occupy: ctas > 0 runs them as that many persistent
 * CTAs, 0 restores one CTA per tile, -1 defers to the LG_CONV_CTAS environment variable (default). Process-wide; used by
 * SamplePipeline so that decoding batch i does not evict the latency-bound AR sampling of batch i+1 from the SMs
 * (the reference runs the two back to back, sample_c2i_ddp.py:128-143). */
int         lg_vq_set_cta_budget(int ctas);

/* ---- stand-alone kernels exported for unit parity tests ------------------------------------------- */
/* y[M,N] (f32) = x[M,K] * w[N,K]^T, operands in `dtype`; the same dispatch the engine uses. */
int  lg_test_gemm(const void* x, const void* w, int M, int N, int K, int dtype, float* y,
                  void* dev_scratch, size_t scratch_bytes, void* stream);

/* The decode step's direct-epilogue GEMM (csrc/gemm_dx.cu) on its own. mode 0: y (f32 [M,N]) = norm(x) * wa^T; mode 1: h (bf16
 * [M,N], in place) = bf16(h + bf16(norm(x) * wa^T)) — the residual add of gpt.py:255-256; mode 2: ff (bf16 [M,N]) =
 * silu(norm(x) * wa^T) * (norm(x) * wb^T) — gpt.py:167. norm(x) = RMSNorm(x) * normw (gpt.py:143-148) when normw != NULL, else x.
 * x [M,K], wa/wb [N,K], normw [K]: bf16 device pointers. */
int  lg_test_gemm_dx(const void* x, const void* wa, const void* wb, int M, int N, int K, int mode, const void* normw,
                     float eps, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LLAMAGEN_B200_H */
