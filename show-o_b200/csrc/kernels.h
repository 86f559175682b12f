// Internal launcher interface between the engine (host orchestration) and the .cu kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/showo_b200.h"

namespace showo {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ GEMM (gemm.cu)
enum GemmEpi { GEMM_BIAS_BF16 = 0, GEMM_RESID_F32 = 1, GEMM_BIAS_F32 = 2 };
struct GemmArgs {
    const bf16* A; int64_t lda;       // [M,K] bf16, row stride lda (elements, multiple of 8)
    const bf16* B; int64_t ldb;       // [N,K] bf16 (weights, K-contiguous)
    int M, N, K;
    void* out; int64_t ldc;
    const float* bias;
    const float* resid; int64_t ldr;  // GEMM_RESID_F32
    int gelu_from;                    // GEMM_BIAS_BF16: gelu_new on columns >= gelu_from (N = none)
    // GEMM_BIAS_BF16, training step: gelu_mode 1 = columns >= gelu_from are written raw to out AND as gelu_new(bf16 value) to
    // gelu_out[m, n - gelu_from] (the forward keeps the pre-activation for the backward); gelu_mode 2 = columns >= gelu_from are
    // multiplied by gelu_new'(gelu_pre[m, n - gelu_from]) and written to gelu_out[m, n - gelu_from] instead of out (the dgrad GEMM of
    // fc2 producing d fc1 directly)
    int gelu_mode = 0; bf16* gelu_out = nullptr; int64_t gelu_out_ld = 0; const bf16* gelu_pre = nullptr; int64_t gelu_pre_ld = 0;
    int block_n;                      // 0 = auto, else 64 / 128 / 256
    // decode-path extras (gemm_skinny only): fused input LayerNorm of fp32 rows, greedy argmax epilogue
    unsigned long long* argmax_keys = nullptr;
    // decode path: bytes the NEXT kernel of the chain will stream (its weights), pulled into L2 while this kernel runs
    const void* l2_prefetch = nullptr; size_t l2_prefetch_bytes = 0;
    // GEMM_RESID_F32 feeding a LayerNorm-folded GEMM (QkvFuse::ln_part): also write the new residual rows as bf16 to ln_xb and their
    // per-row statistics over every 64-column slot to ln_part ([N / 64][M rows][mean, M2])
    bf16* ln_xb = nullptr; int64_t ln_xb_ld = 0; float* ln_part = nullptr;
    // decode path, GEMM_RESID_F32 with M <= 16: the CTA finishing the last tile also writes LayerNorm(out rows) as bf16 to ln_out
    bf16* ln_out = nullptr; const float* ln_gamma = nullptr; const float* ln_beta = nullptr; float ln_eps = 0.f;
};
int gemm_bf16(const GemmArgs& a, GemmEpi epi, cudaStream_t st);

// The layer's fused projection GEMM (W1 = [Wk; Wv; Wq; Wfc1]) with the q/k LayerNorm + partial rotary + KV-cache scatter
// fused into the epilogue (EPI_QKV_BF16 in gemm_tcgen05.cuh).  out receives q (cols [2D,3D)) and gelu(fc1) (cols [3D,N)).
struct QkvFuse {
    int D, H, rows_per_seq, pos0, Lmax;
    const float* q_gamma; const float* q_beta; const float* k_gamma; const float* k_beta; float eps;
    const float* cos_tab; const float* sin_tab;
    bf16* kcache; bf16* vtcache;
    // LayerNorm folded into this GEMM (inference path, layers >= 1): A is the RAW residual stream in bf16, the weight rows are
    // pre-multiplied by the LayerNorm gamma, and the epilogue applies   y = rstd (acc - mu c_n) + d_n   with the row statistics
    // the previous layer's residual GEMM left in ln_part ([K / 64][M rows][mean, M2] over 64-column slots), c_n = sum_j W'[n][j],
    // d_n = sum_j beta_j W[n][j] + bias_n (passed as the GEMM's bias).  nullptr: the classic LayerNorm'ed A operand.
    const float* ln_part = nullptr; const float* ln_c = nullptr; float ln_eps = 0.f;
};
int gemm_qkv_bf16(const GemmArgs& a, const QkvFuse& f, cudaStream_t st);
// weight-streaming path for M <= 16 rows (decode); epi: 0 bias->bf16(+gelu), 1 resid+bias->f32, 2 bias->f32, 3 fused qkv (gemv.cu)
// C[M, N] (fp32) = A^T B with A = [K][lda >= M] and B = [K][ldb >= N] row-major bf16 (both "token-major": the contraction index is the
// row): the weight-gradient GEMM dW = dY^T X reading dY and X as they lie (MN-major UMMA operands).  Rows past K, features past M / N
// are zero-filled by the TMA unit.  Uses out / ldc of GemmArgs.
int gemm_bf16_tn(const GemmArgs& a, cudaStream_t st);
bool skinny_ln_fold_ok(int K);        // the skinny GEMM in use can run the folded-LayerNorm epilogues for this K
int gemm_skinny(const GemmArgs& a, int epi, const QkvFuse* qf, cudaStream_t st);
// stream-K fix-up workspace shared by the streamed skinny GEMM and the decode megakernel: partial tiles [grid][2][16x64]
// fp32 and self-resetting per-tile tickets
int skinny_workspace(int grid, int tiles, float** partials, int** tickets, cudaStream_t st);

// One decode step (one new token per sequence, M <= 16 rows) of ALL layers in one persistent kernel (decode_mega.cu):
// LayerNorm -> GEMM1 (+ q/k-LN, rotary, KV scatter, gelu) -> attention -> GEMM2 (+ residual) per layer, then the
// final LayerNorm into xh.  Buffers as in run_layers (engine.cu).
struct DecodeMegaLayer {
    const bf16* w1; const bf16* w2;
    const float *b1, *b2, *ln_g, *ln_b, *qg, *qb, *kg, *kb;
    bf16 *kc, *vc;
};
struct DecodeMegaDesc {
    const DecodeMegaLayer* layers;       // host array [NL]
    const bf16* w1_slab; const bf16* w2_slab;   // layers' W1 [NL * W1N, D] and W2 [NL * D, D + F] back to back
    int NL, M, D, F, H, W1N;
    float* x; bf16* xh; bf16* buf;
    float ln_eps; const float* fln_g; const float* fln_b;
    const float* cos_tab; const float* sin_tab;
    int pos0, n_keys, Lmax, max_keys;    // max_keys: longest context of this generation
    int cache_seqs;                      // sequences per layer slab of the KV cache ([NL][cache_seqs][H][Lmax][64])
    const showo_seq_mask_t* masks; float scale;
};
bool decode_mega_supported(const DecodeMegaDesc& d);
int decode_mega_step(const DecodeMegaDesc& d, cudaStream_t st);

// implicit-GEMM convolution on NHWC bf16 (3x3 pad 1, or 1x1), stride 1.  cin multiple of 64, weights [Cout_pad, taps*cin].
struct ConvArgs {
    const bf16* x;          // [NB, H, W, cin]
    const bf16* w;          // [cout_pad(>=cout, multiple of 64), taps*cin]
    const float* bias;      // [cout]
    const bf16* resid;      // optional [NB*H*W, ldr]
    int64_t ldr;
    bf16* out;              // [NB*H*W, ldc]
    int64_t ldc;
    int NB, H, W, cin, cout, taps;
};
int conv_nhwc_bf16(const ConvArgs& a, cudaStream_t st);
int gemm_num_sms();
int make_tmap_2d(void* out_cutensormap, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows);

// ------------------------------------------------------------------ elementwise / norm (elementwise.cu)
// out[r, :] = bf16(LN(x[map(r), :]))   with map(r) = (r / rows_out) * rows_in + row_off + r % rows_out
int layernorm_bf16(const float* x, const float* gamma, const float* beta, float eps, bf16* out, int n_rows_out,
                   int D, int rows_out_per_seq, int rows_in_per_seq, int row_off, cudaStream_t st);
// x[r, :] = float(table[ids[seq(r) * ids_stride + pos0 + r % rows_per_seq], :])
int embed_gather(const int64_t* ids, int64_t ids_stride, int pos0, const bf16* table, float* x, int n_rows,
                 int rows_per_seq, int D, int vocab, cudaStream_t st);
int f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t st);
int gelu_erf_bf16(const bf16* src, bf16* dst, int64_t n, cudaStream_t st);     // nn.GELU() exact (erf) form, out of place
// d_io[i] *= gelu_erf'(pre[i]): the backward of nn.GELU() on bf16, in place on the incoming gradient
int gelu_erf_bwd_bf16(bf16* d_io, const bf16* pre, int64_t n, cudaStream_t st);
// x[r, :] = embeds[r, :] for every row r whose ids[r] < 0 (rows of a mixed ids / embeddings input that carry a caller-supplied vector)
int embed_override(const int64_t* ids, const float* embeds, float* x, int n_rows, int D, cudaStream_t st);
// mean cross-entropy (ignore_index rows skipped) of logits[b0 + b, t0 + t, :] vs labels[b0 + b, t0 + t + shift], b < nb, t < nt;
// ws: 2 * nb * nt floats of scratch; out2: {mean loss, number of counted rows}
int cross_entropy_mean(const float* logits, const int64_t* labels, int64_t L, int V, int b0, int nb, int t0, int nt, int shift,
                       int64_t ignore_index, float* ws, float* out2, cudaStream_t st);
int copy_f32_to_f32_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int rows, int cols, cudaStream_t st);
// dst[r, c0 + c] = bf16(src[r, c])  -- packs weight blocks into the fused weight matrices
int pack_block_bf16(const float* src, int64_t src_ld, bf16* dst, int64_t dst_ld, int rows, int cols, cudaStream_t st);

// q/k LayerNorm(head_dim) + partial rotary + KV-cache scatter.   (phi.py:665-694)
//   qkv: [n_rows, ld] bf16 with k at col 0, v at col D, q at col 2D (per head h: cols h*64 .. h*64+63)
//   row r belongs to sequence r / rows_per_seq at position pos0 + r % rows_per_seq
//   K cache  [seq][H][Lmax][64], V^T cache [seq][H][64][Lmax]
struct QkRopeArgs {
    bf16* qkv; int64_t ld; int n_rows, rows_per_seq, pos0;
    int H, D;
    const float* q_gamma; const float* q_beta; const float* k_gamma; const float* k_beta; float eps;
    const float* cos_tab; const float* sin_tab;   // [max_pos, 32]  (emb = cat(freqs, freqs); we store the 16 freqs twice)
    bf16* kcache; bf16* vtcache; int Lmax;
};
int qk_norm_rope_scatter(const QkRopeArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ attention (attention.cu)
struct AttnArgs {
    bf16* q;            // points at the q column block of the qkv buffer; output overwrites q in place
    int64_t ld;         // row stride (elements)
    int n_seq, H, rows_per_seq, pos0;    // query row r of seq s sits at position pos0 + r
    const bf16* kcache; const bf16* vtcache; int Lmax;
    int n_keys;                          // keys [0, n_keys) are valid in the cache
    const showo_seq_mask_t* masks;       // device array [n_seq]
    float scale;                         // 1/sqrt(head_dim)
    // training step (train.cu): output rows go to `out` (row stride out_ld) instead of overwriting q, and the row's
    // log-sum-exp in the exp2 domain (max * scale * log2e + log2(sum)) is saved at lse[row * H + head]
    bf16* out = nullptr; int64_t out_ld = 0;
    float* lse = nullptr;
    int row_begin = 0;                   // mma.sync kernel: first query row (of every sequence) it processes
    int* work_ctr = nullptr;             // tcgen05 kernel: self-resetting work counter of its tail phase (engine-owned; NULL: per-device default)
    // decode kernel: bytes the next kernel of the chain will stream (its weights), pulled into L2 while this kernel runs
    const void* l2_prefetch = nullptr; size_t l2_prefetch_bytes = 0;
};
int omni_attention(const AttnArgs& a, cudaStream_t st);
// tcgen05/TMEM/TMA kernel (attention_tc.cu) for the leading full 128-row tiles of every sequence: attention_tc_rows() says how
// many rows it takes; omni_attention() sends the remaining rows (row_begin ..) to the mma.sync kernel
int attention_tc_rows(const AttnArgs& a);
int omni_attention_tc(const AttnArgs& a, cudaStream_t st);
int attention_tc_tail_rows(const AttnArgs& a);   // trailing rows the tcgen05 kernel takes itself (its tail phase), else 0
int attention_tc_tail_rows(const AttnArgs& a);   // trailing rows the tcgen05 kernel handles itself (its tail phase), else 0
// single-query (decode) variant: one query row per sequence at position n_keys-1
int omni_attention_decode(const AttnArgs& a, cudaStream_t st);

// backward of the omni-mask attention (attention_bwd.cu), flash-style: the scores are recomputed per 64 x 64 tile from the
// rotated q / k rows, P = exp2(S * scale * log2e - lse), dS = P o (dP - delta); dK / dV by key block, dQ by query block
// (two passes, no atomics, deterministic).  All operands row-major bf16 [n_seq * L, ld] with head h at columns [64h, 64h+64).
struct AttnBwdArgs {
    const bf16* q; int64_t q_ld;
    const bf16* k; int64_t k_ld;
    const bf16* v; int64_t v_ld;
    const bf16* o; int64_t o_ld;         // forward output
    const bf16* d_o; int64_t do_ld;      // gradient of the loss wrt the forward output
    const float* lse;                    // [n_seq * L, H] from the forward (exp2 domain)
    float* delta;                        // [n_seq * L, H] workspace: rowsum(dO o O)
    bf16* dq; int64_t dq_ld;
    bf16* dk; int64_t dk_ld;
    bf16* dv; int64_t dv_ld;
    int n_seq, H, L;
    const showo_seq_mask_t* masks;       // device array [n_seq]
    float scale;
};
int omni_attention_backward(const AttnBwdArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ sampler (sampler.cu)
struct SamplerArgs {
    const float* logits_cond;    // [B * rows_per_seq, ld] fp32, image row n of seq b at row b*rows_per_seq + n
    const float* logits_uncond;  // same layout or nullptr
    int64_t ld; int rows_per_seq;
    int B, N, C;                 // C = codebook size (8192)
    float guidance;              // w; logits = (1+w) cond - w uncond when uncond != nullptr
    int64_t* ids; int64_t ids_stride; int ids_pos0;   // input_ids [B, L]: image part at [pos0, pos0+N)
    int64_t* ids2; int64_t ids2_stride;                // optional mirror (uncond rows) or nullptr
    int64_t* sampled_out;        // [B, N] int64 codes (the function's return value)
    int image_offset; int mask_token_id;
    int mask_len_floor;          // floor(N * mask_ratio) for this step (host-evaluated schedule)
    float temperature;           // compounded temperature for this step
    const float* noise_expo;     // [B*N, C] Exp(1) or nullptr -> Philox
    const float* noise_unif;     // [B, N] U(0,1) or nullptr -> Philox
    uint64_t seed; uint32_t step;
    int row_base = 0;            // Philox mode: global index of this call's first batch row (data parallel: rank * rows per rank), so that
                                 // the noise of a row does not depend on how the batch is split over GPUs
    float* conf_ws;              // [B, N] workspace
    int* sampled_ws;             // [B, N] workspace
    uint8_t* masking_out;        // optional [B, N] (debug / parity), or nullptr
};
int t2i_sampler_step(const SamplerArgs& a, cudaStream_t st);
// greedy / top-k=1 next-token pick for MMU decode: out[b] = argmax_v logits[b, v]; appended to ids at position pos
int argmax_rows(const float* logits, int64_t ld, int B, int V, int64_t* out, cudaStream_t st);
// mmu_generate's next-token draw (modeling_showo.py:219-228): temperature, top-k filter (top_k <= 0: none), softmax,
// categorical draw; Exp(1) noise [B, V] host-supplied (parity mode) or Philox(seed, step)
struct MmuSampleArgs {
    const float* logits; int64_t ld; int B, V; float temperature; int top_k;
    const float* noise_expo; uint64_t seed; uint32_t step;
    int row_base = 0;                    // Philox mode: global index of the first row (see SamplerArgs::row_base)
    int64_t* out; int64_t out_stride;    // token of row b -> out[b * out_stride]
    int64_t* out_next;                   // optional dense [B] copy (the next step's input ids)
};
int mmu_sample(const MmuSampleArgs& a, cudaStream_t st);

}  // namespace showo
