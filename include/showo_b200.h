/* showo_b200.h -- C ABI of libshowo_b200.so, the sm_100a engine behind the Show-o hot path.
 *
 * The reference (showlab/Show-o) is pure Python and has no FFI: its seam is the method surface of
 * `models.Showo` / `models.MAGVITv2`.  Every entry point below names the reference method
 * it stands behind; the Python shim in show-o_b200/ (classes Showo, MAGVITv2 with the reference signatures) is
 * the binding a maintainer would add -- see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; all `*_dev` pointers are device pointers owned by the caller (torch tensors);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises unless stated;
 *   - every function returns 0 on success, non-zero on failure; showo_last_error() describes the failure
 *     (thread-local, valid until the next call on the same thread).  Nothing throws across the ABI;
 *   - an engine handle is re-entrant across handles but not thread-safe per handle;
 *   - there is no CPU fallback: on a box without an sm_100 device every compute entry point fails loudly.
 */
#ifndef SHOWO_B200_H_
#define SHOWO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHOWO_B200_ABI_VERSION 1
#if defined(__GNUC__)
#define SHOWO_API __attribute__((visibility("default")))
#else
#define SHOWO_API
#endif

/* Per-sequence closed form of the omni attention mask (training/omni_attention.py:48-96 mask_mods; equals the dense
 * builders training/prompting_utils.py:466-511,591-624 on every non-pad query row).  Query q may attend key k iff
 *   ( k <= q  ||  full_begin <= q < full_end  ||  win_begin <= k < win_end )  &&  !( k < pad_end && q >= pad_end )
 * t2i row : pad_end = #left pads, full = [soi, eoi], win = empty.   lm row : all zero (pure causal).
 * mmu row : win = [0, eoi_pos+1).   mmu_vit row : win = [1+sys+1, 1+sys+1+576). */
typedef struct {
    int32_t pad_end;
    int32_t full_begin, full_end;
    int32_t win_begin, win_end;
} showo_seq_mask_t;

/* Backbone geometry (PhiConfig as instantiated by models/modeling_showo.py:38-47) + Show-o vocabulary layout
 * (configs/showo_demo.yaml:19-24). head_dim = hidden / n_heads must be 64, rotary_dim 32. */
typedef struct {
    int32_t vocab_size;              /* 58498 */
    int32_t hidden;                  /* 2048  */
    int32_t n_layers;                /* 24    */
    int32_t n_heads;                 /* 32    */
    int32_t ffn;                     /* 8192  */
    int32_t rotary_dim;              /* 32    */
    int32_t max_pos;                 /* 2048  */
    float   ln_eps;                  /* 1e-5  */
    float   rope_theta;              /* 10000 */
    int32_t llm_vocab_size;          /* 50295 */
    int32_t num_new_special_tokens;  /* 10    */
    int32_t codebook_size;           /* 8192  */
} showo_config_t;

typedef struct showo_engine showo_engine_t;

SHOWO_API const char* showo_last_error(void);
SHOWO_API int showo_abi_version(void);
/* number of CUDA devices with compute capability 10.x visible to the library (0 on a CPU box) */
SHOWO_API int showo_device_count(void);

/* models/modeling_showo.py:27-54  Showo.__init__ */
SHOWO_API int showo_engine_create(const showo_config_t* cfg, int device, showo_engine_t** out);
SHOWO_API int showo_engine_destroy(showo_engine_t* e);

/* Showo.from_pretrained / load_state_dict: hand one fp32 tensor of the reference state_dict (key names of
 * SURVEY.md section 8b, e.g. "showo.model.layers.3.mlp.fc1.weight") to the engine.  `data` may be a host or a device
 * pointer (is_device).  The engine converts to bf16 and copies into its fused layout (q|k|v|fc1 rows, dense|fc2
 * columns); it does not keep a reference to `data`.  Synchronous. */
SHOWO_API int showo_load_weight(showo_engine_t* e, const char* name, const float* data, int64_t numel, int is_device);
/* returns 0 when every tensor of the model has been loaded, else sets last_error to the first missing key */
SHOWO_API int showo_weights_complete(showo_engine_t* e);

/* The caller's dense attention mask (training/prompting_utils.py:466-511 create_attention_mask_predict_next, :591-624
 * create_attention_mask_for_mmu[_vit]; inference_t2i.py:300, inference_mmu.py:146,166) -> closed form: one kernel derives the
 * five integers per sequence from mask_dev [B, L, L] (elem_bytes 4: additive fp32, 0 = attend; 1: bool / uint8, non-zero =
 * attend; batch stride in elements) and checks them against every non-pad query row.  mismatches_host[b] != 0 means the
 * tensor is not an omni mask and must not be used with the descriptor.  Synchronises the stream. */
SHOWO_API int showo_mask_descriptors(const void* mask_dev, int elem_bytes, int B, int L, int64_t batch_stride_elems,
                           showo_seq_mask_t* out_host, int32_t* mismatches_host, void* stream);

/* models/modeling_showo.py:59-79  Showo.forward (labels=None): logits fp32 [B, L, vocab].
 * Exactly one of ids_dev [B,L] (int64) / embeds_dev [B,L,hidden] (fp32) is non-NULL.  masks_host: B descriptors. */
SHOWO_API int showo_forward(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                  const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream);

/* One denoise-step forward of t2i_generate (modeling_showo.py:136-147) WITHOUT sampling, for parity tests:
 * runs the text prefix [0, prefix_len) once and the image rows, returns the sliced logits
 * [n_branch*B, N, codebook] fp32 (cond rows first, then uncond rows if uncond_ids_dev != NULL).
 * ids rows are [prefix (prefix_len) | soi | N image ids | eoi], L = prefix_len + N + 2. */
SHOWO_API int showo_t2i_logits(showo_engine_t* e, const int64_t* ids_dev, const int64_t* uncond_ids_dev, int B, int L, int N,
                     int prefix_len, const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream);

/* models/modeling_showo.py:104-181  Showo.t2i_generate.  ids_dev [B,L] int64 is updated IN PLACE like the reference.
 * The host evaluates the schedule (the reference takes an arbitrary Python callable): mask_len_floor[s] =
 * floor(N * noise_schedule((s+1)/T)), temperature[s] = the compounded temperature handed to mask_by_random_topk.
 * noise_expo_dev [T, B*N, C] / noise_unif_dev [T, B, N]: optional host-drawn noise for parity runs (the order
 * torch.multinomial / gumbel_noise consume it); NULL -> counter-based Philox keyed by (seed, step, row, col).
 * sampled_out_dev [B,N] int64 receives the last step's sampled codes (the function's return value).
 * masks_host: B descriptors (guidance == 0 or uncond NULL) or 2B (cond rows then uncond rows). */
SHOWO_API int showo_t2i_generate(showo_engine_t* e, int64_t* ids_dev, const int64_t* uncond_ids_dev, int B, int L, int N,
                       int prefix_len, const showo_seq_mask_t* masks_host, int timesteps, float guidance_scale,
                       const int32_t* mask_len_floor, const float* temperature, const float* noise_expo_dev,
                       const float* noise_unif_dev, uint64_t seed, int64_t* sampled_out_dev, void* stream);

/* models/modeling_showo.py:149-179 + models/sampling.py:31-36: the fused sampler step on caller-supplied logits
 * [B, N, C] fp32 (cond, optional uncond).  ids_dev [B, ids_stride]: image ids at [ids_pos0, ids_pos0+N), updated in
 * place; sampled_out_dev [B,N] int64; masking_out_dev optional [B,N] uint8. */
SHOWO_API int showo_sampler_step(const float* logits_cond_dev, const float* logits_uncond_dev, int B, int N, int C,
                       float guidance_scale, int64_t* ids_dev, int64_t ids_stride, int ids_pos0, int image_offset,
                       int mask_token_id, int mask_len_floor, float temperature, const float* noise_expo_dev,
                       const float* noise_unif_dev, uint64_t seed, uint32_t step, int64_t* sampled_out_dev,
                       uint8_t* masking_out_dev, void* stream);

/* models/modeling_showo.py:183-240  Showo.mmu_generate, batched with a KV cache: row b is exactly what the
 * reference returns for a B=1 call on row b (greedy when top_k == 1).  ids_dev [B, L0] int64 (or embeds_dev
 * [B, L0, hidden] fp32).  out_tokens_dev [B, max_new_tokens] int64, out_lengths_dev [B] int32 (tokens produced
 * up to and including eot; eot_token < 0 disables early stop).  top_k == 1 (the reference script's setting,
 * inference_mmu.py:81) is greedy; top_k <= 0 means None (no filter); otherwise logits / temperature, top-k filter,
 * softmax and a categorical draw per token (:219-228) with Exp(1) noise from noise_expo_dev
 * [max_new_tokens, B, vocab] (parity mode) or, when NULL, from the library's Philox stream keyed by `seed`. */
SHOWO_API int showo_mmu_generate(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L0,
                       const showo_seq_mask_t* masks_host, int max_new_tokens, int top_k, float temperature,
                       int64_t eot_token, uint64_t seed, const float* noise_expo_dev, int64_t* out_tokens_dev,
                       int32_t* out_lengths_dev, void* stream);

/* One F.cross_entropy(ignore_index) term of Showo.forward (models/modeling_showo.py:81-100) on fp32 logits [n_seq, L, V]:
 * rows (b, t), b < nb, t < nt pair logits[b0 + b, t0 + t, :] with labels[b0 + b, t0 + t + shift] (labels int64
 * [n_seq, L]).  t2i: (0, B_t2i, max_seq_length + 1, L - max_seq_length - 1, 0); lm / mmu: (first row, count, 0, L - 1, 1).
 * out2_dev: {mean over the counted rows (NaN when none, like torch), number of counted rows}. */
SHOWO_API int showo_cross_entropy(const float* logits_dev, const int64_t* labels_dev, int64_t L, int V, int b0, int nb, int t0,
                        int nt, int shift, int64_t ignore_index, float* out2_dev, void* stream);

/* The next-token draw of mmu_generate on its own (models/modeling_showo.py:219-228), for parity tests:
 * logits_dev [B, ld >= V] fp32 -> out_tokens_dev [B] int64.  noise_expo_dev [B, V] or NULL (Philox(seed, step)). */
SHOWO_API int showo_mmu_sample(const float* logits_dev, int64_t ld, int B, int V, float temperature, int top_k,
                     const float* noise_expo_dev, uint64_t seed, uint32_t step, int64_t* out_tokens_dev, void* stream);

/* fp32 verification forward ("an fp32 / TF32-accumulate verification mode is needed for any stricter claim"): the same
 * call as showo_forward on a second, plain implementation -- fp32 activations, the fp32 master weights the engine keeps once
 * showo_optimizer_enable has been called, CUDA cores only, six separate Linear layers per block as the reference writes them.  Used by
 * the parity tests to pin the engine to the oracle at fp32 re-association level (token decisions bit-identical) and to measure the
 * fast path's bf16 error.  Not a fallback: no product call routes here. */
SHOWO_API int showo_forward_fp32(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                       const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream);

/* Showo.mm_projector (models/modeling_showo.py:49-54: Linear(1024, 2048) -> nn.GELU() -> Linear(2048, 2048)) on the CLIP-ViT features, as
 * inference_mmu.py:128-131 calls it: feats_dev fp32 [n, 1024] -> out_dev fp32 [n, 2048] (bf16 operands, fp32 accumulation, exact erf GELU).
 * Weights arrive through showo_load_weight under "mm_projector.0.weight" / ".0.bias" / ".2.weight" / ".2.bias" (optional set). */
SHOWO_API int showo_mm_projector(showo_engine_t* e, const float* feats_dev, int64_t n, float* out_dev, void* stream);
/* Its backward, for the rows of the LAST showo_mm_projector call (training/train_w_clip_vit.py:599-601 trains the projector through
 * `input_embeddings`; the CLIP features themselves are frozen, :199-203): dy_dev fp32 [n, 2048] = the rows of showo_backward's
 * dembeds_out that the projector's output occupied.  Gradients (fp32) go to an engine-owned buffer [0.weight | 0.bias | 2.weight |
 * 2.bias] = 6,295,552 elements: read per tensor with showo_read_grad("mm_projector.*"), all-reduce in place through
 * showo_mm_projector_grad_buffer; showo_adamw_step updates the four tensors (weights decayed, biases not) whenever a projector
 * backward has run since the previous step. */
SHOWO_API int showo_mm_projector_backward(showo_engine_t* e, const float* dy_dev, int64_t n, void* stream);
SHOWO_API int showo_mm_projector_grad_buffer(showo_engine_t* e, float** base_dev, int64_t* numel);

/* model.showo.model.embed_tokens(ids) as called from outside (inference_mmu.py:134-136): out fp32 [n, hidden] */
SHOWO_API int showo_embed_tokens(showo_engine_t* e, const int64_t* ids_dev, int64_t n, float* out_dev, void* stream);

/* ---------------------------------------------------------------- training step (training/train.py:589-612)
 * Showo.forward with labels under autograd (models/modeling_showo.py:59-102): the same logits and cross-entropy terms as
 * showo_forward + showo_cross_entropy, with every layer's activations kept in engine-owned buffers for showo_backward.
 * terms_host: 3 x {b0, nb, t0, nt, shift} (t2i, lm, mmu) exactly as in showo_cross_entropy.  logits_out_dev [B, L, vocab]
 * fp32 may be NULL (the logits then stay in an engine buffer).  losses_out_dev: float[3][2] = {mean, counted rows}.
 * Inputs: ids_dev alone, embeds_dev alone ([B, L, hidden] fp32), or BOTH = the mixed rows of training/train_w_clip_vit.py:532-537:
 * positions with ids >= 0 are looked up in the embedding table, positions with ids < 0 take embeds_dev[b, t, :] (the mm_projector
 * output); the backward then scatters the embedding gradient of the former and hands the gradient of the latter back in
 * showo_backward's dembeds_out_dev. */
SHOWO_API int showo_train_forward(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                        const showo_seq_mask_t* masks_host, const int64_t* labels_dev, const int32_t* terms_host,
                        int64_t ignore_index, float* logits_out_dev, float* losses_out_dev, void* stream);
/* loss.backward() (training/train.py:612) for loss = sum_i loss_grads[i] * loss_i of the last showo_train_forward:
 * gradients of every parameter go to an engine-owned fp32 buffer (read with showo_read_grad); dembeds_out_dev, optional
 * [B, L, hidden] fp32, receives the gradient wrt the input embeddings (the mm_projector / embed_tokens side of
 * train_w_clip_vit.py).  loss_grads_dev: float[3] on the device. */
SHOWO_API int showo_backward(showo_engine_t* e, const float* loss_grads_dev, float* dembeds_out_dev, void* stream);
/* The same backward in phases, so that the caller can overlap the gradient all-reduce of finished parameters with the backward of
 * the earlier layers (what DDP's buckets do under accelerate, training/train.py:612): phase -1 = loss, head and final LayerNorm,
 * then phase l = decoder layer l for l = n_layers-1 .. 0, then phase -2 = embedding (+ dembeds_out_dev).  showo_backward runs all
 * of them in this order. */
SHOWO_API int showo_backward_phase(showo_engine_t* e, int phase, const float* loss_grads_dev, float* dembeds_out_dev, void* stream);
/* The engine's gradient buffer (fp32, one allocation, the packed layout of the weights) and the [begin, end) element range a phase
 * of showo_backward_phase writes: the buckets of the data-parallel gradient all-reduce. */
SHOWO_API int showo_grad_buffer(showo_engine_t* e, float** base_dev, int64_t* numel);
SHOWO_API int showo_grad_range(showo_engine_t* e, int phase, int64_t* begin, int64_t* end);
/* The optimizer step of the training loop in the engine (training/train.py:211-236 AdamW construction, :617 optimizer.step()):
 * torch.optim.AdamW's update with decoupled weight decay on every parameter whose name has no "bias" (the reference's other no_decay
 * patterns match no Phi parameter name), on engine-owned fp32 master weights and moments, one pass per tensor over gradient + master +
 * moments that also rewrites the engine's bf16 / fp32 working copy.  showo_optimizer_enable allocates masters and moments (3 x 5.8 GB
 * for the full model) and must be called BEFORE the weights are handed over with showo_load_weight (it forgets the loaded set).
 * showo_adamw_step consumes the gradients of the last showo_backward (all-reduced in place by the caller when data parallel);
 * the bias-correction step counter lives in the engine.  showo_read_param returns the fp32 master of one reference parameter. */
SHOWO_API int showo_optimizer_enable(showo_engine_t* e);
SHOWO_API int showo_adamw_step(showo_engine_t* e, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);
SHOWO_API int showo_read_param(showo_engine_t* e, const char* name, float* out_dev, int64_t numel, void* stream);
/* gradient of one parameter of the reference state_dict (same names as showo_load_weight) -> out_dev (fp32, contiguous) */
SHOWO_API int showo_read_grad(showo_engine_t* e, const char* name, float* out_dev, int64_t numel, void* stream);

/* The producer of the training step's t2i rows, on the device (training/train.py:468-488): training/utils.py:77-154
 * mask_or_random_replace_tokens (noise_type "mask", no contiguous-region masking, predict_all_tokens off -- every shipped yaml)
 * and training/prompting_utils.py:39-90 UniversalPrompting.t2i_prompt.  mode bit 1 = mask step, bit 2 = prompt step; 3 = both in
 * one launch.
 *   image_tokens_dev [B, N] int64: codes + text-vocabulary offset (train.py:476-477).  text_ids_dev [B, text_stride] int64 /
 *   text_len_dev [B] int32: the tokenised captions without specials (tokenisation is host work).  special_ids_host: int64[8] =
 *   {pad, bos, eos, task (<|t2i|>), soi, eoi, mask_id, ignore_id}.  schedule: 0 cosine, 1 linear, 2 pow (schedule_param), 3 =
 *   timesteps_dev already holds mask_prob (any Python schedule evaluated by the caller).
 *   Noise, in the order the reference draws it: timesteps_dev [B] = torch.rand(batch_size), rand_dev [B, N] =
 *   torch.rand(batch_size, seq_len) (position j is masked iff argsort(rand)[j] < round(N * mask_prob), stable ties),
 *   drop_probs_dev [B] = torch.rand(len(text_ids)) of t2i_prompt; any of them NULL -> the library's Philox stream keyed by seed.
 *   Outputs: mode 3 / 2: input_ids_out_dev, labels_out_dev [B, L] int64 with L = max_text_len + 1 + N + 2; attn_ones_out_dev
 *   optional [B, L + 1] int64 (t2i_prompt's attention_masks: the reference computes the pad count after padding, so the row is
 *   all ones and one element longer than the sequence); descs_out_dev optional DEVICE array of B mask descriptors (what
 *   create_attention_mask_predict_next gives for these rows); mask_prob_out_dev optional [B].  mode 1: input_ids_out_dev /
 *   labels_out_dev are the [B, N] masked ids / labels of mask_or_random_replace_tokens; mode 2 reads them back through
 *   masked_in_dev / labels_in_dev. */
SHOWO_API int showo_t2i_train_prep(const int64_t* image_tokens_dev, int B, int N, const int64_t* text_ids_dev,
                         const int32_t* text_len_dev, int64_t text_stride, int max_text_len, const int64_t* special_ids_host,
                         float min_masking_rate, float cond_dropout_prob, int schedule, float schedule_param,
                         const float* timesteps_dev, const float* rand_dev, const float* drop_probs_dev, uint64_t seed,
                         int mode, const int64_t* masked_in_dev, const int64_t* labels_in_dev, int64_t* input_ids_out_dev,
                         int64_t* labels_out_dev, int64_t* attn_ones_out_dev, showo_seq_mask_t* descs_out_dev,
                         float* mask_prob_out_dev, void* stream);

/* Data-parallel generation: global index of the first batch row of this engine's calls (rank * rows per rank).  The library's
 * Philox noise (showo_t2i_generate / showo_mmu_generate without host noise) is keyed by (seed, GLOBAL row, token, step), so with the same
 * seed on every rank the images / tokens of a row do not depend on how the batch is split over GPUs.  Default 0. */
SHOWO_API int showo_set_rng_row_base(showo_engine_t* e, int64_t first_row);

/* seconds spent / kernels launched by the last generate call (for bench.py's gpu_launches) */
SHOWO_API int64_t showo_kernel_launches(showo_engine_t* e);

/* ---------------------------------------------------------------- MAGVIT-v2 (models/modeling_magvitv2.py:402-433) */
typedef struct magvit_engine magvit_engine_t;
SHOWO_API int magvit_engine_create(int device, magvit_engine_t** out);
SHOWO_API int magvit_engine_destroy(magvit_engine_t* m);
/* one fp32 tensor of MAGVITv2.state_dict() ("decoder.up.2.block.1.conv1.weight", ...) */
SHOWO_API int magvit_load_weight(magvit_engine_t* m, const char* name, const float* data, int64_t numel, int is_device);
SHOWO_API int magvit_weights_complete(magvit_engine_t* m);
/* MAGVITv2.decode_code(ids, shape=(h,w)): ids_dev [B, h*w] int64 -> pixels_out_dev [B,3,16h,16w] fp32 (NCHW) */
SHOWO_API int magvit_decode_code(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, float* pixels_out_dev,
                       void* stream);
/* decode + the caller's post-processing (inference_t2i.py:338-341): clamp((x+1)/2,0,1)*255 -> uint8 NHWC [B,16h,16w,3] */
SHOWO_API int magvit_decode_code_u8(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, uint8_t* images_out_dev,
                          void* stream);
/* MAGVITv2.get_code(pixel_values): pixels_dev [B,3,R,R] fp32 NCHW -> ids_out_dev [B,(R/16)^2] int64 */
SHOWO_API int magvit_get_code(magvit_engine_t* m, const float* pixels_dev, int B, int R, int64_t* ids_out_dev, void* stream);
/* fp32 verification path ("an fp32-accumulate verification mode is needed for any stricter claim"): the same two
 * entry points on fp32 NCHW activations and fp32 weights, CUDA cores only -- a plain second implementation used by the parity tests to
 * show that the fast path differs from the reference only by bf16 rounding.  z_out_dev (optional) receives the quantizer's pre-sign
 * values [B, 13, R/16, R/16].  Not a fallback: no product call routes here. */
SHOWO_API int magvit_decode_code_fp32(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, float* pixels_out_dev, void* stream);
SHOWO_API int magvit_get_code_fp32(magvit_engine_t* m, const float* pixels_dev, int B, int R, int64_t* ids_out_dev, float* z_out_dev,
                         void* stream);
SHOWO_API int64_t magvit_kernel_launches(magvit_engine_t* m);

/* ---------------------------------------------------------------- CLIP ViT vision tower (models/clip_encoder.py:6-51)
 * The frozen encoder in front of the w_clip_vit MMU path: CLIPVisionTower.forward = transformers' CLIPVisionModel(images,
 * output_hidden_states=True).hidden_states[-2][:, 1:] (inference_mmu.py:100-131, training/train_w_clip_vit.py:532-537).  The network
 * is third-party code (transformers, pinned 4.41.1 in requirements.txt; openai/clip-vit-large-patch14-336); oracle/clip_oracle.py
 * restates it and is pinned to the live library.  head_dim must be 64, hidden a multiple of 128 (<= 2048). */
typedef struct clip_engine clip_engine_t;
typedef struct {
    int32_t image_size, patch_size;   /* 336, 14 */
    int32_t hidden, n_layers, n_heads, ffn;   /* 1024, 24, 16, 4096 */
    float ln_eps;                     /* 1e-5 */
} clip_config_t;
SHOWO_API int clip_engine_create(const clip_config_t* cfg, int device, clip_engine_t** out);
SHOWO_API int clip_engine_destroy(clip_engine_t* e);
/* one fp32 tensor of CLIPVisionModel.state_dict() ("vision_model.encoder.layers.3.self_attn.q_proj.weight", ...) */
SHOWO_API int clip_load_weight(clip_engine_t* e, const char* name, const float* data, int64_t numel, int is_device);
SHOWO_API int clip_weights_complete(clip_engine_t* e);
/* pixels_dev [B, 3, S, S] fp32 (image-processor output) -> out_dev fp32 [B, T - 1, hidden] (drop_cls != 0, the reference's
 * select_feature 'patch') or [B, T, hidden] ('cls_patch'); select_layer indexes hidden_states like the reference (-2). */
SHOWO_API int clip_forward(clip_engine_t* e, const float* pixels_dev, int B, int select_layer, int drop_cls, float* out_dev, void* stream);
SHOWO_API int64_t clip_kernel_launches(clip_engine_t* e);

/* ---------------------------------------------------------------- raw kernels, exported for the parity tests */
/* C[M,N] (+epilogue) = A[M,K] bf16 * B[N,K]^T bf16 on tcgen05; epi 0: bf16 out = acc+bias (gelu_new on cols >=
 * gelu_from), 1: f32 out = resid + acc + bias, 2: f32 out = acc + bias.  block_n 0 = auto.
 * epi 3: the weight-gradient form C[M,N] f32 = A^T B (+ bias[n]) with A = [K, lda >= M] and B = [K, ldb >= N] row-major (the
 * contraction index is the row of both; MN-major tcgen05 operands, no transposed copies). */
SHOWO_API int showo_gemm_bf16(const void* A_dev, int64_t lda, const void* B_dev, int64_t ldb, int M, int N, int K, void* out_dev,
                    int64_t ldc, const float* bias_dev, const float* resid_dev, int64_t ldr, int gelu_from, int epi,
                    int block_n, void* stream);
/* omni-mask attention over a [n_seq*rows, ld] bf16 buffer holding k|v|q column blocks (test entry: runs q/k layernorm
 * + rotary + cache scatter, then attention; output overwrites the q block). */
SHOWO_API int showo_attention_test(void* qkv_dev, int64_t ld, int n_seq, int rows_per_seq, int pos0, int H,
                         const float* qg, const float* qb, const float* kg, const float* kb, float eps,
                         float rope_theta, int rotary_dim, void* kcache_dev, void* vtcache_dev, int Lmax, int n_keys,
                         const showo_seq_mask_t* masks_host, void* stream);
/* backward of the omni-mask attention on its own: q/k (rotated), v, o, d_o bf16 [n_seq*L, H*64], lse fp32 [n_seq*L, H] in the
 * exp2 domain (max*scale*log2e + log2(sum)); outputs dq, dk, dv bf16 [n_seq*L, H*64].  scale = 1/8. */
SHOWO_API int showo_attention_bwd_test(const void* q_dev, const void* k_dev, const void* v_dev, const void* o_dev, const void* do_dev,
                             const float* lse_dev, void* dq_dev, void* dk_dev, void* dv_dev, int n_seq, int L, int H,
                             const showo_seq_mask_t* masks_host, void* stream);
/* the prefill / denoise-step attention kernels alone on prepared operands (benchmarks, profiling): q bf16 [n_seq*rows, ld]
 * (head h at columns 64h..64h+63, already normalised + rotated), K cache [n_seq][H][Lmax][64], V^T cache [n_seq][H][64][Lmax],
 * descriptors ON THE DEVICE; output to out_dev (row stride out_ld) or, when NULL, in place over q. */
SHOWO_API int showo_attention_run(void* q_dev, int64_t ld, int n_seq, int rows_per_seq, int pos0, int H, const void* kcache_dev,
                        const void* vtcache_dev, int Lmax, int n_keys, const showo_seq_mask_t* masks_dev, void* out_dev,
                        int64_t out_ld, void* stream);
SHOWO_API int showo_layernorm_test(const float* x_dev, const float* gamma_dev, const float* beta_dev, float eps, void* out_bf16_dev,
                         int rows, int D, void* stream);
/* NHWC bf16 3x3 (taps=9) or 1x1 (taps=1) convolution, stride 1, same padding; w [cout_pad, taps*cin] bf16 */
SHOWO_API int showo_conv_test(const void* x_dev, const void* w_dev, const float* bias_dev, const void* resid_dev, void* out_dev,
                    int NB, int H, int W, int cin, int cout, int taps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHOWO_B200_H_ */
