// Host orchestration of the Show-o backbone behind the C ABI (include/showo_b200.h): weight packing, workspaces,
// the per-layer kernel sequence, the t2i denoise loop with text-prefix reuse, KV-cached batched MMU decoding.
//
// Per layer (PhiDecoderLayer, phi.py:774-790: y = Attn(LN x) + MLP(LN x) + x with ONE shared pre-LN) the engine runs
//   1. layernorm_bf16        x(f32) -> xh(bf16)
//   2. gemm  [M,D] x W1^T    W1 = [Wk; Wv; Wq; Wfc1]  ([3D+F, D])  -> buf [M, 3D+F] bf16 = k | v | q | gelu_new(fc1)
//   3. qk_norm_rope_scatter  q/k LayerNorm(64) + partial rotary; K -> cache, V -> cache (transposed); q in place
//   4. omni_attention        softmax(q K^T / 8 + mask) V  -> overwrites the q block (so buf = k | v | attn | act)
//   5. gemm  [M,D+F] x W2^T  W2 = [Wdense | Wfc2] ([D, D+F]), A = buf[:, 2D:], epilogue x += acc + (b_dense + b_fc2)
// i.e. two GEMMs per layer instead of six, and the attention/MLP outputs never round-trip HBM separately.
#include <math.h>

#include <atomic>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "engine_state.h"

namespace showo {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<int64_t> g_launches{0};
void note_launch(int n) { g_launches += n; }
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}
int64_t launches_total() { return g_launches.load(); }
bool fused_decode_ln() {
    static int v = -1;
    // opt-in: measured 1.89 vs 1.45 ms per decode step -- one CTA normalising 16 rows behind the last tile (L2 round trips for x, cold
    // gamma / beta) is a longer serial tail than the stand-alone launch it removes
    if (v < 0) { const char* e = getenv("SHOWO_DECODE_LN_FUSED"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v == 1;
}
bool l2_prefetch_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_L2_PREFETCH"); v = (e && atoi(e) == 1) ? 1 : 0; }    // opt-in: measured 1.61 vs 1.45 ms per decode step
    return v == 1;
}
// SHOWO_LN_FOLD: 1 (default) = the decode path (M <= 16 rows) runs without LayerNorm launches for layers >= 1, 2 = every path, 0 = off.
// Measured on one box: decode 1.345 vs 1.460 ms per 16-token step; the 6192-row t2i step is unchanged within noise by it (the
// projection GEMM's epilogue grows by what the LayerNorm launch cost), so the big GEMMs keep the classic operand.
int ln_fold_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_LN_FOLD"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
    return v;
}
bool ln_fold_enabled() { return ln_fold_mode() != 0; }
const char* last_error_cstr() { return g_last_error.c_str(); }

// LayerNorm folded into the projection that consumes it:  LN(x) W^T + b = rstd (x W'^T - mu c) + d  with W' = W * gamma (rounded to
// bf16 -- c sums the ROUNDED values, so a constant row cancels exactly), c_n = sum_j W'[n][j], d_n = b_n + sum_j beta_j W[n][j].
// One warp per output feature.
__global__ void ln_fold_weights_kernel(const bf16* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int N, int K, bf16* __restrict__ wf, float* __restrict__ c,
                                       float* __restrict__ d) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float sc = 0.f, sd = 0.f;
    for (int j = lane; j < K; j += 32) {
        const float wv = __bfloat162float(w[(size_t)n * K + j]);
        const bf16 r = __float2bfloat16(wv * gamma[j]);
        wf[(size_t)n * K + j] = r;
        sc += __bfloat162float(r);
        sd = fmaf(beta[j], wv, sd);
    }
    sc = warp_sum(sc); sd = warp_sum(sd);
    if (lane == 0) { c[n] = sc; d[n] = bias[n] + sd; }
}

__global__ void vec_add_kernel(const float* a, const float* b, float* o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}
// unpack the fused greedy head's per-row key -> token id; publish it (tok for the next embedding gather, out[b*stride]) and
// reset the key for the next step
__global__ void mmu_finish_token_kernel(unsigned long long* keys, int B, int64_t* tok, int64_t* out, int out_stride, int64_t eot,
                                        int* finished) {
    const int b = threadIdx.x;
    if (b >= B) return;
    const unsigned long long k = keys[b];
    const int64_t id = (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    tok[b] = id;
    out[(int64_t)b * out_stride] = id;
    keys[b] = 0ull;
    if (id == eot) finished[b] = 1;
}
// rows that have produced eot_token so far (modeling_showo.py:236-237 stops a B = 1 call there)
__global__ void mmu_mark_finished_kernel(const int64_t* tok, int B, int64_t eot, int* finished) {
    const int b = threadIdx.x;
    if (b < B && tok[b] == eot) finished[b] = 1;
}
__global__ void mmu_lengths_kernel(const int64_t* toks, int max_new, int64_t eot, int32_t* lens) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int n = max_new;
    if (eot >= 0)
        for (int i = 0; i < max_new; ++i)
            if (toks[(int64_t)b * max_new + i] == eot) { n = i + 1; break; }
    lens[b] = n;
}

}  // namespace showo

using namespace showo;

static int engine_set_device(showo_engine* e) {
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    return 0;
}

static int ensure_ws(showo_engine* e, int rows, int n_seq, int L, int64_t logit_elems, cudaStream_t st) {
    const int Lr = cdiv(L, 64) * 64;
    if (rows > e->cap_rows) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        dev_free(e->x); dev_free(e->xh); dev_free(e->buf);
        SHOWO_TRY(dev_alloc(&e->x, (size_t)rows * e->D));
        SHOWO_TRY(dev_alloc(&e->xh, (size_t)rows * e->D));
        SHOWO_TRY(dev_alloc(&e->buf, (size_t)rows * e->W1N));
        dev_free(e->ln_part);
        SHOWO_TRY(dev_alloc(&e->ln_part, (size_t)rows * (e->D / 64) * 2));
        e->cap_rows = rows;
    }
    if (n_seq > e->cap_seq || Lr > e->cap_L) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        const int ns = n_seq > e->cap_seq ? n_seq : e->cap_seq;
        const int nl = Lr > e->cap_L ? Lr : e->cap_L;
        dev_free(e->kcache); dev_free(e->vtcache); dev_free(e->d_masks);
        const size_t n = (size_t)e->NL * ns * e->H * nl * 64;
        SHOWO_TRY(dev_alloc(&e->kcache, n));
        SHOWO_TRY(dev_alloc(&e->vtcache, n));
        SHOWO_CUDA_OK(cudaMemset(e->kcache, 0, n * sizeof(bf16)));
        SHOWO_CUDA_OK(cudaMemset(e->vtcache, 0, n * sizeof(bf16)));
        SHOWO_TRY(dev_alloc(&e->d_masks, (size_t)ns));
        e->cap_seq = ns; e->cap_L = nl;
    }
    if (logit_elems > e->cap_logit_elems) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        dev_free(e->logits_ws);
        SHOWO_TRY(dev_alloc(&e->logits_ws, (size_t)logit_elems));
        e->cap_logit_elems = logit_elems;
    }
    if (!e->conf_ws) {
        SHOWO_TRY(dev_alloc(&e->conf_ws, (size_t)1 << 20));
        SHOWO_TRY(dev_alloc(&e->sampled_ws, (size_t)1 << 20));
    }
    return 0;
}

static size_t layer_cache_stride(const showo_engine* e) { return (size_t)e->cap_seq * e->H * e->cap_L * 64; }

// One pass of all layers over rows laid out as [n_seq][rows_per_seq] in e->x (fp32 residual stream).
static int run_layers(showo_engine* e, int n_seq, int rows_per_seq, int pos0, int n_keys, bool decode, cudaStream_t st) {
    const int M = n_seq * rows_per_seq;
    const int D = e->D, F = e->F;
    for (int l = 0; l < e->NL; ++l) {
        const LayerW& w = e->layers[l];
        // decode: only the first layer's LayerNorm is a launch of its own -- every later one (and the final LayerNorm) is done by the
        // CTA that finishes the last tile of the previous layer's second GEMM (skinny.cuh: sk2_tile_done)
        // LayerNorm folding (SHOWO_LN_FOLD=1): layers >= 1 have no LayerNorm launch at all -- the previous layer's residual GEMM leaves
        // bf16(x) in xh and the rows' slot statistics in ln_part, and the projection GEMM (gamma-scaled weights) corrects in its epilogue
        const bool fold = e->w1f_slab != nullptr && !fused_decode_ln() &&       // the explicit opt-in of the older fusion wins
                          (ln_fold_mode() == 2 ? (M > 16 || skinny_ln_fold_ok(D)) : (ln_fold_mode() == 1 && decode && M <= 16 && skinny_ln_fold_ok(D)));
        const bool ln_fused = !fold && decode && fused_decode_ln() && M <= 16 && D % 128 == 0 && D <= 2048;
        if (!((ln_fused || fold) && l > 0)) SHOWO_TRY(layernorm_bf16(e->x, w.ln_g, w.ln_b, e->cfg.ln_eps, e->xh, M, D, M, M, 0, st));
        bf16* kc = e->kcache + (size_t)l * layer_cache_stride(e);
        bf16* vc = e->vtcache + (size_t)l * layer_cache_stride(e);
        // GEMM1 + (q/k LayerNorm, partial rotary, K / V^T cache scatter, gelu_new) in one kernel
        GemmArgs g1{};
        g1.A = e->xh; g1.lda = D; g1.B = w.w1; g1.ldb = D; g1.M = M; g1.N = e->W1N; g1.K = D;
        g1.out = e->buf; g1.ldc = e->W1N; g1.bias = w.b1;
        QkvFuse qf{};
        qf.D = D; qf.H = e->H; qf.rows_per_seq = rows_per_seq; qf.pos0 = pos0; qf.Lmax = e->cap_L;
        qf.q_gamma = w.qg; qf.q_beta = w.qb; qf.k_gamma = w.kg; qf.k_beta = w.kb; qf.eps = e->cfg.ln_eps;
        qf.cos_tab = e->cos_tab; qf.sin_tab = e->sin_tab; qf.kcache = kc; qf.vtcache = vc;
        if (fold && l > 0) { g1.B = w.w1f; g1.bias = w.ln_d; qf.ln_part = e->ln_part; qf.ln_c = w.ln_c; qf.ln_eps = e->cfg.ln_eps; }
        SHOWO_TRY(gemm_qkv_bf16(g1, qf, st));
        AttnArgs a{};
        a.q = e->buf + 2 * D; a.ld = e->W1N; a.n_seq = n_seq; a.H = e->H; a.rows_per_seq = rows_per_seq; a.pos0 = pos0;
        a.kcache = kc; a.vtcache = vc; a.Lmax = e->cap_L; a.n_keys = n_keys; a.masks = e->d_masks; a.scale = 0.125f;
        a.work_ctr = e->attn_ctr;
        if (decode && l2_prefetch_enabled()) { a.l2_prefetch = w.w2; a.l2_prefetch_bytes = (size_t)D * e->W2K * sizeof(bf16); }
        if (decode) SHOWO_TRY(omni_attention_decode(a, st));
        else SHOWO_TRY(omni_attention(a, st));
        GemmArgs g2{};
        g2.A = e->buf + 2 * D; g2.lda = e->W1N; g2.B = w.w2; g2.ldb = e->W2K; g2.M = M; g2.N = D; g2.K = D + F;
        g2.out = e->x; g2.ldc = D; g2.bias = w.b2; g2.resid = e->x; g2.ldr = D;
        if (fold && l + 1 < e->NL) { g2.ln_xb = e->xh; g2.ln_xb_ld = D; g2.ln_part = e->ln_part; }
        if (ln_fused) {
            g2.ln_out = e->xh; g2.ln_eps = e->cfg.ln_eps;
            g2.ln_gamma = l + 1 < e->NL ? e->layers[l + 1].ln_g : e->fln_g;
            g2.ln_beta = l + 1 < e->NL ? e->layers[l + 1].ln_b : e->fln_b;
        }
        if (decode && l2_prefetch_enabled()) {      // the next layer's fused projection (or the start of the head) streams next
            if (l + 1 < e->NL) { g2.l2_prefetch = e->layers[l + 1].w1; g2.l2_prefetch_bytes = (size_t)e->W1N * D * sizeof(bf16); }
            else { g2.l2_prefetch = e->head_w; g2.l2_prefetch_bytes = std::min((size_t)e->V * D * sizeof(bf16), (size_t)64 << 20); }
        }
        SHOWO_TRY(gemm_bf16(g2, GEMM_RESID_F32, st));
    }
    return 0;
}

// One decode step (one token per sequence at position pos0) through all layers + the final LayerNorm into e->xh:
// the persistent megakernel when the geometry allows it, else the per-kernel path.
static int decode_step_layers(showo_engine* e, int B, int pos0, int max_keys, cudaStream_t st) {
    DecodeMegaDesc d{};
    std::vector<DecodeMegaLayer> hl((size_t)e->NL);
    for (int l = 0; l < e->NL; ++l) {
        const LayerW& w = e->layers[l];
        hl[l] = DecodeMegaLayer{w.w1, w.w2, w.b1, w.b2, w.ln_g, w.ln_b, w.qg, w.qb, w.kg, w.kb,
                                e->kcache + (size_t)l * layer_cache_stride(e), e->vtcache + (size_t)l * layer_cache_stride(e)};
    }
    d.layers = hl.data(); d.w1_slab = e->w1_slab; d.w2_slab = e->w2_slab; d.NL = e->NL; d.M = B; d.D = e->D; d.F = e->F; d.H = e->H; d.W1N = e->W1N;
    d.x = e->x; d.xh = e->xh; d.buf = e->buf; d.ln_eps = e->cfg.ln_eps; d.fln_g = e->fln_g; d.fln_b = e->fln_b;
    d.cos_tab = e->cos_tab; d.sin_tab = e->sin_tab;
    d.pos0 = pos0; d.n_keys = pos0 + 1; d.Lmax = e->cap_L; d.max_keys = max_keys; d.cache_seqs = e->cap_seq;
    d.masks = e->d_masks; d.scale = 0.125f;
    static int check = -1, nl_lim = 0;
    if (check < 0) {
        const char* c = getenv("SHOWO_MEGA_CHECK"); check = c ? atoi(c) : 0;
        const char* n = getenv("SHOWO_MEGA_NL"); nl_lim = n ? atoi(n) : 0;
    }
    if (check && decode_mega_supported(d)) {
        // debug: run the per-kernel path and the megakernel from the same state and report the differences
        const int nl = nl_lim > 0 && nl_lim < e->NL ? nl_lim : e->NL;
        const size_t nx = (size_t)B * e->D, nb = (size_t)B * e->W1N;
        std::vector<float> x0(nx), xr(nx), xm(nx);
        std::vector<bf16> hr(nx), hm(nx), br(nb), bm(nb);
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        SHOWO_CUDA_OK(cudaMemcpy(x0.data(), e->x, nx * 4, cudaMemcpyDeviceToHost));
        const int keep = e->NL; e->NL = nl;
        int rc = run_layers(e, B, 1, pos0, pos0 + 1, true, st);
        e->NL = keep;
        SHOWO_TRY(rc);
        SHOWO_TRY(layernorm_bf16(e->x, e->fln_g, e->fln_b, e->cfg.ln_eps, e->xh, B, e->D, B, B, 0, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        SHOWO_CUDA_OK(cudaMemcpy(xr.data(), e->x, nx * 4, cudaMemcpyDeviceToHost));
        SHOWO_CUDA_OK(cudaMemcpy(hr.data(), e->xh, nx * 2, cudaMemcpyDeviceToHost));
        SHOWO_CUDA_OK(cudaMemcpy(br.data(), e->buf, nb * 2, cudaMemcpyDeviceToHost));
        SHOWO_CUDA_OK(cudaMemcpy(e->x, x0.data(), nx * 4, cudaMemcpyHostToDevice));
        d.NL = nl;
        SHOWO_TRY(decode_mega_step(d, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        SHOWO_CUDA_OK(cudaMemcpy(xm.data(), e->x, nx * 4, cudaMemcpyDeviceToHost));
        SHOWO_CUDA_OK(cudaMemcpy(hm.data(), e->xh, nx * 2, cudaMemcpyDeviceToHost));
        SHOWO_CUDA_OK(cudaMemcpy(bm.data(), e->buf, nb * 2, cudaMemcpyDeviceToHost));
        double dx = 0, dh = 0, dq = 0, da = 0, ax = 0;
        for (size_t i = 0; i < nx; ++i) {
            dx = std::max(dx, (double)fabsf(xr[i] - xm[i])); ax = std::max(ax, (double)fabsf(xr[i]));
            dh = std::max(dh, (double)fabsf(__bfloat162float(hr[i]) - __bfloat162float(hm[i])));
        }
        for (int r = 0; r < B; ++r)
            for (int c = 2 * e->D; c < e->W1N; ++c) {
                const double v = fabsf(__bfloat162float(br[(size_t)r * e->W1N + c]) - __bfloat162float(bm[(size_t)r * e->W1N + c]));
                if (c < 3 * e->D) dq = std::max(dq, v); else da = std::max(da, v);
            }
        fprintf(stderr, "[mega check] pos %d layers %d: max|dx| %.3e (|x| %.3e)  |dxh| %.3e  |d attn_out| %.3e  |d fc1_act| %.3e\n",
                pos0, nl, dx, ax, dh, dq, da);
        return 0;
    }
    if (decode_mega_supported(d)) return decode_mega_step(d, st);
    SHOWO_TRY(run_layers(e, B, 1, pos0, pos0 + 1, true, st));
    if (fused_decode_ln() && B <= 16 && e->D % 128 == 0 && e->D <= 2048) return 0;      // the last layer's GEMM wrote LN_f(x) to xh
    return layernorm_bf16(e->x, e->fln_g, e->fln_b, e->cfg.ln_eps, e->xh, B, e->D, B, B, 0, st);
}

// blocking look at the rows' eot flags (a 256-byte read-back every 16 decode steps)
static bool all_rows_finished(showo_engine* e, int B, cudaStream_t st) {
    int h[64];
    if (cudaMemcpyAsync(h, e->finished_ws, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    for (int b = 0; b < B; ++b) if (!h[b]) return false;
    return true;
}

static int upload_masks(showo_engine* e, const showo_seq_mask_t* masks_host, int n, cudaStream_t st) {
    SHOWO_CUDA_OK(cudaMemcpyAsync(e->d_masks, masks_host, (size_t)n * sizeof(showo_seq_mask_t), cudaMemcpyHostToDevice, st));
    return 0;
}

static int check_ready(showo_engine* e) {
    SHOWO_CHECK(e != nullptr, "null engine");
    SHOWO_TRY(engine_set_device(e));
    if (!e->finalized) SHOWO_TRY(showo_weights_complete(e));
    return 0;
}

int engine_check_ready(showo_engine* e) { return check_ready(e); }
static int derive_ln_fold(showo_engine* e, cudaStream_t st) {
    if (!ln_fold_enabled()) return 0;
    if (!e->w1f_slab) {
        SHOWO_TRY(dev_alloc(&e->w1f_slab, (size_t)e->NL * e->W1N * e->D));
        SHOWO_TRY(dev_alloc(&e->ln_cd, (size_t)e->NL * 2 * e->W1N));
        for (size_t l = 0; l < e->layers.size(); ++l) {
            e->layers[l].w1f = e->w1f_slab + l * (size_t)e->W1N * e->D;
            e->layers[l].ln_c = e->ln_cd + l * 2 * (size_t)e->W1N;
            e->layers[l].ln_d = e->layers[l].ln_c + e->W1N;
        }
    }
    for (size_t l = 1; l < e->layers.size(); ++l) {
        LayerW& w = e->layers[l];
        ln_fold_weights_kernel<<<cdiv(e->W1N, 8), 256, 0, st>>>(w.w1, w.b1, w.ln_g, w.ln_b, e->W1N, e->D, w.w1f, w.ln_c, w.ln_d);
        note_launch();
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

int engine_refresh_derived(showo_engine* e, cudaStream_t st) {
    for (auto& w : e->layers) {
        vec_add_kernel<<<cdiv(e->D, 256), 256, 0, st>>>(w.b_dense, w.b_fc2, w.b2, e->D);
        note_launch();
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    SHOWO_TRY(derive_ln_fold(e, st));
    const int off = e->cfg.llm_vocab_size + e->cfg.num_new_special_tokens, C = e->cfg.codebook_size;
    if (e->head_b_img) SHOWO_CUDA_OK(cudaMemcpyAsync(e->head_b_img, e->head_b + off, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    ++e->weights_version;                         // the training step re-derives its transposed weight copies
    return 0;
}
int engine_upload_masks(showo_engine* e, const showo_seq_mask_t* masks_host, int n, cudaStream_t st) { return upload_masks(e, masks_host, n, st); }
int engine_ensure_ws(showo_engine* e, int rows, int n_seq, int L, int64_t logit_elems, cudaStream_t st) { return ensure_ws(e, rows, n_seq, L, logit_elems, st); }

extern "C" {

const char* showo_last_error(void) { return last_error_cstr(); }
int showo_abi_version(void) { return SHOWO_B200_ABI_VERSION; }
int showo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int showo_engine_create(const showo_config_t* cfg, int device, showo_engine_t** out) {
    SHOWO_CHECK(cfg && out, "null argument");
    SHOWO_CHECK(cfg->hidden % 128 == 0 && cfg->hidden <= 2048, "hidden must be a multiple of 128 and <= 2048");
    SHOWO_CHECK(cfg->hidden / cfg->n_heads == 64 && cfg->hidden % cfg->n_heads == 0, "head_dim must be 64");
    SHOWO_CHECK(cfg->rotary_dim == 32, "rotary_dim must be 32");
    SHOWO_CHECK(cfg->ffn % 64 == 0, "ffn must be a multiple of 64");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CUDA_OK(cudaSetDevice(device));
    showo_engine* e = new showo_engine();
    e->cfg = *cfg; e->device = device;
    e->D = cfg->hidden; e->H = cfg->n_heads; e->F = cfg->ffn; e->NL = cfg->n_layers; e->V = cfg->vocab_size;
    e->W1N = 3 * e->D + e->F; e->W2K = e->D + e->F;
    const int D = e->D, V = e->V;
    SHOWO_TRY(dev_alloc(&e->embed, (size_t)V * D));
    SHOWO_TRY(dev_alloc(&e->head_w, (size_t)V * D));
    SHOWO_TRY(dev_alloc(&e->head_b, (size_t)V));
    SHOWO_TRY(dev_alloc(&e->fln_g, (size_t)D));
    SHOWO_TRY(dev_alloc(&e->fln_b, (size_t)D));
    e->layers.resize(e->NL);
    SHOWO_TRY(dev_alloc(&e->w1_slab, (size_t)e->NL * e->W1N * D));
    SHOWO_TRY(dev_alloc(&e->w2_slab, (size_t)e->NL * D * e->W2K));
    for (size_t li = 0; li < e->layers.size(); ++li) {
        LayerW& w = e->layers[li];
        w.w1 = e->w1_slab + li * (size_t)e->W1N * D;
        w.w2 = e->w2_slab + li * (size_t)D * e->W2K;
        SHOWO_TRY(dev_alloc(&w.b1, (size_t)e->W1N));
        SHOWO_TRY(dev_alloc(&w.b2, (size_t)D));
        SHOWO_TRY(dev_alloc(&w.b_dense, (size_t)D));
        SHOWO_TRY(dev_alloc(&w.b_fc2, (size_t)D));
        SHOWO_TRY(dev_alloc(&w.ln_g, (size_t)D));
        SHOWO_TRY(dev_alloc(&w.ln_b, (size_t)D));
        SHOWO_TRY(dev_alloc(&w.qg, 64)); SHOWO_TRY(dev_alloc(&w.qb, 64));
        SHOWO_TRY(dev_alloc(&w.kg, 64)); SHOWO_TRY(dev_alloc(&w.kb, 64));
    }
    // rotary tables (phi.py:79-112): inv_freq_i = theta^(-2i/32), emb = cat(freqs, freqs)
    {
        const int P = cfg->max_pos;
        std::vector<float> c((size_t)P * 32), s((size_t)P * 32);
        for (int p = 0; p < P; ++p)
            for (int i = 0; i < 16; ++i) {
                const float inv = 1.0f / powf(cfg->rope_theta, (float)(2 * i) / 32.0f);
                const float f = (float)p * inv;
                c[(size_t)p * 32 + i] = c[(size_t)p * 32 + 16 + i] = cosf(f);
                s[(size_t)p * 32 + i] = s[(size_t)p * 32 + 16 + i] = sinf(f);
            }
        SHOWO_TRY(dev_alloc(&e->cos_tab, c.size()));
        SHOWO_TRY(dev_alloc(&e->sin_tab, s.size()));
        SHOWO_CUDA_OK(cudaMemcpy(e->cos_tab, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
        SHOWO_CUDA_OK(cudaMemcpy(e->sin_tab, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
    }
    SHOWO_TRY(dev_alloc(&e->attn_ctr, (size_t)16));
    SHOWO_CUDA_OK(cudaMemset(e->attn_ctr, 0, 16 * sizeof(int)));
    *out = e;
    return 0;
}

int showo_engine_destroy(showo_engine_t* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    dev_free(e->embed); dev_free(e->head_w); dev_free(e->head_b); dev_free(e->head_b_img); dev_free(e->fln_g); dev_free(e->fln_b);
    for (auto& w : e->layers) {
        dev_free(w.b1); dev_free(w.b2); dev_free(w.b_dense); dev_free(w.b_fc2);
        dev_free(w.ln_g); dev_free(w.ln_b); dev_free(w.qg); dev_free(w.qb); dev_free(w.kg); dev_free(w.kb);
    }
    dev_free(e->w1_slab); dev_free(e->w2_slab);
    dev_free(e->cos_tab); dev_free(e->sin_tab); dev_free(e->stage);
    dev_free(e->w1f_slab); dev_free(e->ln_cd); dev_free(e->ln_part);
    dev_free(e->x); dev_free(e->xh); dev_free(e->buf); dev_free(e->kcache); dev_free(e->vtcache); dev_free(e->d_masks);
    dev_free(e->logits_ws); dev_free(e->conf_ws); dev_free(e->sampled_ws); dev_free(e->tok_ws); dev_free(e->argmax_keys); dev_free(e->finished_ws); dev_free(e->attn_ctr);
    dev_free(e->mmp_w0); dev_free(e->mmp_b0); dev_free(e->mmp_w2); dev_free(e->mmp_b2); dev_free(e->mmp_in); dev_free(e->mmp_mid);
    dev_free(e->mmp_pre); dev_free(e->mmp_grads); dev_free(e->mmp_w2t); dev_free(e->mmp_dy); dev_free(e->mmp_dmid);
    if (e->train) train_state_destroy(e->train);
    if (e->opt) opt_state_destroy(e->opt);
    delete e;
    return 0;
}

int showo_load_weight(showo_engine_t* e, const char* name_c, const float* data, int64_t numel, int is_device) {
    SHOWO_CHECK(e && name_c && data, "null argument");
    SHOWO_TRY(engine_set_device(e));
    const std::string name(name_c);
    const int D = e->D, F = e->F, V = e->V;
    // stage on the device as fp32
    const float* src = data;
    if (!is_device) {
        if ((size_t)numel > e->stage_cap) {
            dev_free(e->stage);
            SHOWO_TRY(dev_alloc(&e->stage, (size_t)numel));
            e->stage_cap = (size_t)numel;
        }
        SHOWO_CUDA_OK(cudaMemcpy(e->stage, data, (size_t)numel * 4, cudaMemcpyHostToDevice));
        src = e->stage;
    }
    cudaStream_t st = 0;
    auto expect = [&](int64_t n) -> int {
        SHOWO_CHECK(numel == n, "weight " + name + ": expected " + std::to_string(n) + " elements, got " + std::to_string(numel));
        return 0;
    };
    auto copy_f32 = [&](float* dst, int64_t n) -> int {
        SHOWO_TRY(expect(n));
        SHOWO_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    };
    int rc = -2;
    const std::string lp = "showo.model.layers.";
    if (name == "showo.model.embed_tokens.weight") {
        rc = expect((int64_t)V * D); if (!rc) rc = f32_to_bf16(src, e->embed, numel, st);
    } else if (name == "showo.lm_head.weight") {
        rc = expect((int64_t)V * D); if (!rc) rc = f32_to_bf16(src, e->head_w, numel, st);
    } else if (name == "showo.lm_head.bias") {
        rc = copy_f32(e->head_b, V);
    } else if (name == "showo.model.final_layernorm.weight") {
        rc = copy_f32(e->fln_g, D);
    } else if (name == "showo.model.final_layernorm.bias") {
        rc = copy_f32(e->fln_b, D);
    } else if (name.compare(0, lp.size(), lp) == 0) {
        const size_t dot = name.find('.', lp.size());
        SHOWO_CHECK(dot != std::string::npos, "bad weight name " + name);
        const int l = atoi(name.substr(lp.size(), dot - lp.size()).c_str());
        SHOWO_CHECK(l >= 0 && l < e->NL, "layer index out of range in " + name);
        LayerW& w = e->layers[l];
        const std::string k = name.substr(dot + 1);
        // W1 rows: k | v | q | fc1 ; W2 columns: dense | fc2
        if (k == "self_attn.k_proj.weight") { rc = expect((int64_t)D * D); if (!rc) rc = pack_block_bf16(src, D, w.w1, D, D, D, st); }
        else if (k == "self_attn.v_proj.weight") { rc = expect((int64_t)D * D); if (!rc) rc = pack_block_bf16(src, D, w.w1 + (size_t)D * D, D, D, D, st); }
        else if (k == "self_attn.q_proj.weight") { rc = expect((int64_t)D * D); if (!rc) rc = pack_block_bf16(src, D, w.w1 + (size_t)2 * D * D, D, D, D, st); }
        else if (k == "mlp.fc1.weight") { rc = expect((int64_t)F * D); if (!rc) rc = pack_block_bf16(src, D, w.w1 + (size_t)3 * D * D, D, F, D, st); }
        else if (k == "self_attn.k_proj.bias") rc = copy_f32(w.b1, D);
        else if (k == "self_attn.v_proj.bias") rc = copy_f32(w.b1 + D, D);
        else if (k == "self_attn.q_proj.bias") rc = copy_f32(w.b1 + 2 * D, D);
        else if (k == "mlp.fc1.bias") rc = copy_f32(w.b1 + 3 * D, F);
        else if (k == "self_attn.dense.weight") { rc = expect((int64_t)D * D); if (!rc) rc = pack_block_bf16(src, D, w.w2, e->W2K, D, D, st); }
        else if (k == "mlp.fc2.weight") { rc = expect((int64_t)D * F); if (!rc) rc = pack_block_bf16(src, F, w.w2 + D, e->W2K, D, F, st); }
        else if (k == "self_attn.dense.bias") rc = copy_f32(w.b_dense, D);
        else if (k == "mlp.fc2.bias") rc = copy_f32(w.b_fc2, D);
        else if (k == "input_layernorm.weight") rc = copy_f32(w.ln_g, D);
        else if (k == "input_layernorm.bias") rc = copy_f32(w.ln_b, D);
        else if (k == "self_attn.q_layernorm.weight") rc = copy_f32(w.qg, 64);
        else if (k == "self_attn.q_layernorm.bias") rc = copy_f32(w.qb, 64);
        else if (k == "self_attn.k_layernorm.weight") rc = copy_f32(w.kg, 64);
        else if (k == "self_attn.k_layernorm.bias") rc = copy_f32(w.kb, 64);
        else { set_last_error("unknown weight name " + name); rc = -2; }
    } else if (name.compare(0, 13, "mm_projector.") == 0) {
        // optional (w_clip_vit): not part of the backbone's required set; with the optimizer enabled its fp32 values are kept as
        // masters too (train.cu: the four tensors are updated by showo_adamw_step whenever showo_mm_projector_backward has run)
        constexpr int64_t kIn = 1024, kMid = 2048, kOut = 2048;
        if (!e->mmp_w0) {
            SHOWO_TRY(dev_alloc(&e->mmp_w0, (size_t)(kMid * kIn))); SHOWO_TRY(dev_alloc(&e->mmp_b0, (size_t)kMid));
            SHOWO_TRY(dev_alloc(&e->mmp_w2, (size_t)(kOut * kMid))); SHOWO_TRY(dev_alloc(&e->mmp_b2, (size_t)kOut));
        }
        if (name == "mm_projector.0.weight") { rc = expect(kMid * kIn); if (!rc) rc = f32_to_bf16(src, e->mmp_w0, numel, st); }
        else if (name == "mm_projector.0.bias") rc = copy_f32(e->mmp_b0, kMid);
        else if (name == "mm_projector.2.weight") { rc = expect(kOut * kMid); if (!rc) rc = f32_to_bf16(src, e->mmp_w2, numel, st); }
        else if (name == "mm_projector.2.bias") rc = copy_f32(e->mmp_b2, kOut);
        else { set_last_error("unknown weight name " + name); rc = -2; }
        if (rc) return rc;
        if (e->opt) SHOWO_TRY(opt_store_master(e, name, src, numel, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        e->loaded.insert(name);
        ++e->mmp_version;
        return 0;
    } else {
        set_last_error("unknown weight name " + name);
        rc = -2;
    }
    if (rc) return rc;
    if (e->opt) SHOWO_TRY(opt_store_master(e, name, src, numel, st));
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    e->loaded.insert(name);
    e->finalized = false;
    ++e->weights_version;
    return 0;
}

int showo_weights_complete(showo_engine_t* e) {
    SHOWO_CHECK(e, "null engine");
    SHOWO_TRY(engine_set_device(e));
    std::vector<std::string> need = {"showo.model.embed_tokens.weight", "showo.lm_head.weight", "showo.lm_head.bias",
                                     "showo.model.final_layernorm.weight", "showo.model.final_layernorm.bias"};
    const char* per_layer[] = {"self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight",
                               "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias",
                               "self_attn.dense.weight", "self_attn.dense.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                               "mlp.fc2.weight", "mlp.fc2.bias", "input_layernorm.weight", "input_layernorm.bias",
                               "self_attn.q_layernorm.weight", "self_attn.q_layernorm.bias",
                               "self_attn.k_layernorm.weight", "self_attn.k_layernorm.bias"};
    for (int l = 0; l < e->NL; ++l)
        for (const char* k : per_layer) need.push_back("showo.model.layers." + std::to_string(l) + "." + k);
    for (const auto& n : need) SHOWO_CHECK(e->loaded.count(n) == 1, "weight not loaded: " + n);
    for (auto& w : e->layers) {
        vec_add_kernel<<<cdiv(e->D, 256), 256>>>(w.b_dense, w.b_fc2, w.b2, e->D);
        SHOWO_CUDA_OK(cudaGetLastError());
    }
    {
        const int off = e->cfg.llm_vocab_size + e->cfg.num_new_special_tokens, C = e->cfg.codebook_size;
        SHOWO_CHECK(off + C <= e->V, "vocabulary layout: image codes exceed vocab_size");
        if (!e->head_b_img) SHOWO_TRY(dev_alloc(&e->head_b_img, (size_t)C));
        SHOWO_CUDA_OK(cudaMemcpy(e->head_b_img, e->head_b + off, (size_t)C * 4, cudaMemcpyDeviceToDevice));
    }
    SHOWO_TRY(derive_ln_fold(e, nullptr));
    SHOWO_CUDA_OK(cudaDeviceSynchronize());
    e->finalized = true;
    return 0;
}

int showo_embed_tokens(showo_engine_t* e, const int64_t* ids_dev, int64_t n, float* out_dev, void* stream) {
    SHOWO_TRY(check_ready(e));
    return embed_gather(ids_dev, 0, 0, e->embed, out_dev, (int)n, (int)n, e->D, e->V, (cudaStream_t)stream);
}

int showo_forward(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                  const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream) {
    SHOWO_TRY(check_ready(e));
    SHOWO_CHECK((ids_dev != nullptr) != (embeds_dev != nullptr), "forward: exactly one of ids / embeds");
    SHOWO_CHECK(B > 0 && L > 0 && L <= e->cfg.max_pos && masks_host && logits_out_dev, "forward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    const int M = B * L;
    SHOWO_TRY(ensure_ws(e, M, B, L, 0, st));
    SHOWO_TRY(upload_masks(e, masks_host, B, st));
    if (ids_dev) SHOWO_TRY(embed_gather(ids_dev, L, 0, e->embed, e->x, M, L, e->D, e->V, st));
    else SHOWO_CUDA_OK(cudaMemcpyAsync(e->x, embeds_dev, (size_t)M * e->D * 4, cudaMemcpyDeviceToDevice, st));
    SHOWO_TRY(run_layers(e, B, L, 0, L, false, st));
    SHOWO_TRY(layernorm_bf16(e->x, e->fln_g, e->fln_b, e->cfg.ln_eps, e->xh, M, e->D, M, M, 0, st));
    GemmArgs g{};
    g.A = e->xh; g.lda = e->D; g.B = e->head_w; g.ldb = e->D; g.M = M; g.N = e->V; g.K = e->D;
    g.out = logits_out_dev; g.ldc = e->V; g.bias = e->head_b;
    SHOWO_TRY(gemm_bf16(g, GEMM_BIAS_F32, st));
    e->launches_last = launches_total() - l0;
    return 0;
}

// text prefix [0,P): computed once per generate call; K/V land in the cache, hidden states are dropped.
// Left-pad rows are never read by any other row (their columns are masked for every row past the pads), so only the
// positions [p0, P) with p0 = the smallest pad_end of the batch are computed: typically 20-70 of the 129 prefix rows.
static int t2i_prefix(showo_engine* e, const int64_t* ids, const int64_t* uncond, int B, int L, int P, int nb,
                      const showo_seq_mask_t* masks_host, cudaStream_t st) {
    if (P == 0) return 0;
    int p0 = P;
    for (int i = 0; i < nb * B; ++i) p0 = masks_host[i].pad_end < p0 ? masks_host[i].pad_end : p0;
    if (p0 < 0) p0 = 0;
    if (p0 >= P) p0 = P - 1;
    const int W = P - p0;
    SHOWO_TRY(embed_gather(ids, L, p0, e->embed, e->x, B * W, W, e->D, e->V, st));
    if (nb == 2) SHOWO_TRY(embed_gather(uncond, L, p0, e->embed, e->x + (size_t)B * W * e->D, B * W, W, e->D, e->V, st));
    return run_layers(e, nb * B, W, p0, P, false, st);
}
// image rows [P, L): every step.  Leaves sliced logits [nb*B*N, C] in e->logits_ws.
static int t2i_step_logits(showo_engine* e, const int64_t* ids, int B, int L, int N, int P, int nb, cudaStream_t st) {
    const int R = L - P;
    const int C = e->cfg.codebook_size;
    const int off = e->cfg.llm_vocab_size + e->cfg.num_new_special_tokens;
    // both branches share the image part of the cond ids (modeling_showo.py:137-138)
    SHOWO_TRY(embed_gather(ids, L, P, e->embed, e->x, B * R, R, e->D, e->V, st));
    if (nb == 2) SHOWO_TRY(embed_gather(ids, L, P, e->embed, e->x + (size_t)B * R * e->D, B * R, R, e->D, e->V, st));
    SHOWO_TRY(run_layers(e, nb * B, R, P, L, false, st));
    // final LN on the N image positions only, then the image-vocab slice of the head (modeling_showo.py:144)
    SHOWO_TRY(layernorm_bf16(e->x, e->fln_g, e->fln_b, e->cfg.ln_eps, e->xh, nb * B * N, e->D, N, R, R - N - 1, st));
    GemmArgs g{};
    g.A = e->xh; g.lda = e->D; g.B = e->head_w + (size_t)off * e->D; g.ldb = e->D; g.M = nb * B * N; g.N = C; g.K = e->D;
    g.out = e->logits_ws; g.ldc = C; g.bias = e->head_b_img;
    return gemm_bf16(g, GEMM_BIAS_F32, st);
}

static int t2i_check(showo_engine* e, int B, int L, int N, int P) {
    SHOWO_CHECK(B > 0 && N > 0 && P >= 0, "t2i: bad sizes");
    SHOWO_CHECK((P > 0 && L == P + N + 2) || (P == 0 && L >= N + 2), "t2i: need L == prefix_len + N + 2");
    SHOWO_CHECK(L <= e->cfg.max_pos, "t2i: sequence longer than max_pos");
    SHOWO_CHECK(e->cfg.codebook_size % 4 == 0, "t2i: codebook size must be a multiple of 4");
    return 0;
}

int showo_t2i_logits(showo_engine_t* e, const int64_t* ids_dev, const int64_t* uncond_ids_dev, int B, int L, int N,
                     int prefix_len, const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream) {
    SHOWO_TRY(check_ready(e));
    SHOWO_CHECK(ids_dev && masks_host && logits_out_dev, "t2i_logits: null argument");
    SHOWO_TRY(t2i_check(e, B, L, N, prefix_len));
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = uncond_ids_dev ? 2 : 1;
    const int C = e->cfg.codebook_size;
    const int P = prefix_len, R = L - P;
    const int rows = nb * B * (R > P ? R : P);
    SHOWO_TRY(ensure_ws(e, rows, nb * B, L, (int64_t)nb * B * N * C, st));
    SHOWO_TRY(upload_masks(e, masks_host, nb * B, st));
    SHOWO_TRY(t2i_prefix(e, ids_dev, uncond_ids_dev, B, L, P, nb, masks_host, st));
    SHOWO_TRY(t2i_step_logits(e, ids_dev, B, L, N, P, nb, st));
    SHOWO_CUDA_OK(cudaMemcpyAsync(logits_out_dev, e->logits_ws, (size_t)nb * B * N * C * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int showo_t2i_generate(showo_engine_t* e, int64_t* ids_dev, const int64_t* uncond_ids_dev, int B, int L, int N,
                       int prefix_len, const showo_seq_mask_t* masks_host, int timesteps, float guidance_scale,
                       const int32_t* mask_len_floor, const float* temperature, const float* noise_expo_dev,
                       const float* noise_unif_dev, uint64_t seed, int64_t* sampled_out_dev, void* stream) {
    SHOWO_TRY(check_ready(e));
    SHOWO_CHECK(ids_dev && masks_host && mask_len_floor && temperature && sampled_out_dev, "t2i_generate: null argument");
    SHOWO_CHECK(timesteps > 0, "t2i_generate: timesteps must be positive");
    SHOWO_TRY(t2i_check(e, B, L, N, prefix_len));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    const bool cfg_on = uncond_ids_dev != nullptr && guidance_scale > 0.f;     // modeling_showo.py:136
    const int nb = cfg_on ? 2 : 1;
    const int C = e->cfg.codebook_size;
    const int P = prefix_len, R = L - P;
    const int rows = nb * B * (R > P ? R : P);
    SHOWO_CHECK((int64_t)B * N <= (1 << 20), "t2i_generate: B*N too large");
    SHOWO_TRY(ensure_ws(e, rows, nb * B, L, (int64_t)nb * B * N * C, st));
    SHOWO_TRY(upload_masks(e, masks_host, nb * B, st));
    SHOWO_TRY(t2i_prefix(e, ids_dev, uncond_ids_dev, B, L, P, nb, masks_host, st));
    for (int s = 0; s < timesteps; ++s) {
        SHOWO_TRY(t2i_step_logits(e, ids_dev, B, L, N, P, nb, st));
        SamplerArgs sa{};
        sa.logits_cond = e->logits_ws;
        sa.logits_uncond = cfg_on ? e->logits_ws + (size_t)B * N * C : nullptr;
        sa.ld = C; sa.rows_per_seq = N; sa.B = B; sa.N = N; sa.C = C; sa.guidance = guidance_scale;
        sa.ids = ids_dev; sa.ids_stride = L; sa.ids_pos0 = L - N - 1; sa.ids2 = nullptr; sa.ids2_stride = 0;
        sa.sampled_out = sampled_out_dev;
        sa.image_offset = e->cfg.llm_vocab_size + e->cfg.num_new_special_tokens;
        sa.mask_token_id = e->V - 1;
        sa.mask_len_floor = mask_len_floor[s]; sa.temperature = temperature[s];
        sa.noise_expo = noise_expo_dev ? noise_expo_dev + (size_t)s * B * N * C : nullptr;
        sa.noise_unif = noise_unif_dev ? noise_unif_dev + (size_t)s * B * N : nullptr;
        sa.seed = seed; sa.step = (uint32_t)s; sa.row_base = e->rng_row_base;
        sa.conf_ws = e->conf_ws; sa.sampled_ws = e->sampled_ws; sa.masking_out = nullptr;
        SHOWO_TRY(t2i_sampler_step(sa, st));
    }
    e->launches_last = launches_total() - l0;
    return 0;
}

int showo_sampler_step(const float* logits_cond_dev, const float* logits_uncond_dev, int B, int N, int C,
                       float guidance_scale, int64_t* ids_dev, int64_t ids_stride, int ids_pos0, int image_offset,
                       int mask_token_id, int mask_len_floor, float temperature, const float* noise_expo_dev,
                       const float* noise_unif_dev, uint64_t seed, uint32_t step, int64_t* sampled_out_dev,
                       uint8_t* masking_out_dev, void* stream) {
    SHOWO_CHECK(logits_cond_dev && ids_dev && sampled_out_dev, "sampler_step: null argument");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    float* conf = nullptr; int* samp = nullptr;
    SHOWO_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&conf), (size_t)B * N * 4, st));
    SHOWO_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&samp), (size_t)B * N * 4, st));
    SamplerArgs sa{};
    sa.logits_cond = logits_cond_dev; sa.logits_uncond = logits_uncond_dev; sa.ld = C; sa.rows_per_seq = N;
    sa.B = B; sa.N = N; sa.C = C; sa.guidance = guidance_scale;
    sa.ids = ids_dev; sa.ids_stride = ids_stride; sa.ids_pos0 = ids_pos0; sa.sampled_out = sampled_out_dev;
    sa.image_offset = image_offset; sa.mask_token_id = mask_token_id; sa.mask_len_floor = mask_len_floor;
    sa.temperature = temperature; sa.noise_expo = noise_expo_dev; sa.noise_unif = noise_unif_dev;
    sa.seed = seed; sa.step = step; sa.conf_ws = conf; sa.sampled_ws = samp; sa.masking_out = masking_out_dev;
    int rc = t2i_sampler_step(sa, st);
    cudaFreeAsync(conf, st);
    cudaFreeAsync(samp, st);
    return rc;
}

int showo_mmu_generate(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L0,
                       const showo_seq_mask_t* masks_host, int max_new_tokens, int top_k, float temperature,
                       int64_t eot_token, uint64_t seed, const float* noise_expo_dev, int64_t* out_tokens_dev,
                       int32_t* out_lengths_dev, void* stream) {
    SHOWO_TRY(check_ready(e));
    SHOWO_CHECK((ids_dev != nullptr) != (embeds_dev != nullptr), "mmu_generate: exactly one of ids / embeds");
    SHOWO_CHECK(B > 0 && L0 > 0 && max_new_tokens > 0 && masks_host && out_tokens_dev, "mmu_generate: bad arguments");
    SHOWO_CHECK(temperature > 0.f, "mmu_generate: temperature must be positive");
    SHOWO_CHECK(L0 + max_new_tokens <= e->cfg.max_pos, "mmu_generate: sequence would exceed max_pos");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    const int D = e->D, V = e->V;
    const int Ltot = L0 + max_new_tokens;
    SHOWO_TRY(ensure_ws(e, B * L0, B, Ltot, (int64_t)B * V, st));
    SHOWO_TRY(upload_masks(e, masks_host, B, st));
    // ---- prefill
    if (ids_dev) SHOWO_TRY(embed_gather(ids_dev, L0, 0, e->embed, e->x, B * L0, L0, D, V, st));
    else SHOWO_CUDA_OK(cudaMemcpyAsync(e->x, embeds_dev, (size_t)B * L0 * D * 4, cudaMemcpyDeviceToDevice, st));
    SHOWO_TRY(run_layers(e, B, L0, 0, L0, false, st));
    SHOWO_TRY(layernorm_bf16(e->x, e->fln_g, e->fln_b, e->cfg.ln_eps, e->xh, B, D, 1, L0, L0 - 1, st));
    GemmArgs g{};
    g.A = e->xh; g.lda = D; g.B = e->head_w; g.ldb = D; g.M = B; g.N = V; g.K = D;
    g.out = e->logits_ws; g.ldc = V; g.bias = e->head_b;
    // token t of row b lives at out_tokens[b*max_new + t]; argmax writes a [B] vector -> strided copy via tok_ws
    if (e->tok_ws_cap < B) {
        dev_free(e->tok_ws);
        SHOWO_TRY(dev_alloc(&e->tok_ws, (size_t)B));
        e->tok_ws_cap = B;
    }
    // early stop: every 16 tokens the host looks at the rows' eot flags and leaves the loop when all have finished
    if (!e->finished_ws) SHOWO_TRY(dev_alloc(&e->finished_ws, (size_t)64));
    SHOWO_CHECK(B <= 64 || eot_token < 0, "mmu_generate: early stop supports up to 64 rows");
    SHOWO_CUDA_OK(cudaMemsetAsync(e->finished_ws, 0, 64 * sizeof(int), st));
    const bool greedy = top_k == 1;              // a one-entry distribution: the draw is the argmax whatever the noise
    if (greedy && B <= 16 && D % 64 == 0) {
        // decode fast path: greedy pick fused into the head GEMM's epilogue (no logits tensor, no argmax pass)
        if (!e->argmax_keys) SHOWO_TRY(dev_alloc(&e->argmax_keys, (size_t)16));
        SHOWO_CUDA_OK(cudaMemsetAsync(e->argmax_keys, 0, 16 * 8, st));      // a call that failed half-way must not leak keys into this one
        GemmArgs ga = g;
        ga.out = nullptr; ga.argmax_keys = e->argmax_keys;
        SHOWO_TRY(gemm_skinny(ga, 4 /*SK_ARGMAX*/, nullptr, st));                      // prefill: xh already holds LN(last row)
        for (int t = 0; t < max_new_tokens; ++t) {
            mmu_finish_token_kernel<<<1, 32, 0, st>>>(e->argmax_keys, B, e->tok_ws, out_tokens_dev + t, max_new_tokens, eot_token, e->finished_ws);
            SHOWO_CUDA_OK(cudaGetLastError());
            note_launch();
            if (t == max_new_tokens - 1) break;
            if (eot_token >= 0 && (t & 15) == 15 && all_rows_finished(e, B, st)) break;
            SHOWO_TRY(embed_gather(e->tok_ws, 1, 0, e->embed, e->x, B, 1, D, V, st));
            SHOWO_TRY(decode_step_layers(e, B, L0 + t, Ltot, st));
            SHOWO_TRY(gemm_skinny(ga, 4 /*SK_ARGMAX*/, nullptr, st));
        }
    } else {
    SHOWO_TRY(gemm_bf16(g, GEMM_BIAS_F32, st));
    for (int t = 0; t < max_new_tokens; ++t) {
        if (greedy) {
            SHOWO_TRY(argmax_rows(e->logits_ws, V, B, V, e->tok_ws, st));
            SHOWO_CUDA_OK(cudaMemcpy2DAsync(out_tokens_dev + t, (size_t)max_new_tokens * 8, e->tok_ws, 8, 8, B,
                                            cudaMemcpyDeviceToDevice, st));
        } else {
            MmuSampleArgs ms{};
            ms.logits = e->logits_ws; ms.ld = V; ms.B = B; ms.V = V; ms.temperature = temperature; ms.top_k = top_k;
            ms.noise_expo = noise_expo_dev ? noise_expo_dev + (size_t)t * B * V : nullptr;
            ms.seed = seed; ms.step = (uint32_t)t; ms.row_base = e->rng_row_base;
            ms.out = out_tokens_dev + t; ms.out_stride = max_new_tokens; ms.out_next = e->tok_ws;
            SHOWO_TRY(mmu_sample(ms, st));
        }
        if (t == max_new_tokens - 1) break;
        if (eot_token >= 0) {
            mmu_mark_finished_kernel<<<1, 64, 0, st>>>(e->tok_ws, B, eot_token, e->finished_ws);
            note_launch();
            if ((t & 15) == 15 && all_rows_finished(e, B, st)) break;
        }
        // ---- decode one token per row at position L0 + t
        SHOWO_TRY(embed_gather(e->tok_ws, 1, 0, e->embed, e->x, B, 1, D, V, st));
        SHOWO_TRY(decode_step_layers(e, B, L0 + t, Ltot, st));
        SHOWO_TRY(gemm_bf16(g, GEMM_BIAS_F32, st));
    }
    }
    if (out_lengths_dev) {
        mmu_lengths_kernel<<<1, B, 0, st>>>(out_tokens_dev, max_new_tokens, eot_token, out_lengths_dev);
        SHOWO_CUDA_OK(cudaGetLastError());
        note_launch();
    }
    e->launches_last = launches_total() - l0;
    return 0;
}

int showo_cross_entropy(const float* logits_dev, const int64_t* labels_dev, int64_t L, int V, int b0, int nb, int t0, int nt,
                        int shift, int64_t ignore_index, float* out2_dev, void* stream) {
    SHOWO_CHECK(logits_dev && labels_dev && out2_dev && V > 0 && L > 0 && nb >= 0 && nt >= 0, "cross_entropy: bad arguments");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = nullptr;
    const size_t n = (size_t)nb * nt;
    if (n > 0) SHOWO_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), 2 * n * 4, st));
    const int rc = cross_entropy_mean(logits_dev, labels_dev, L, V, b0, nb, t0, nt, shift, ignore_index, ws, out2_dev, st);
    if (ws) cudaFreeAsync(ws, st);
    return rc;
}

int showo_mmu_sample(const float* logits_dev, int64_t ld, int B, int V, float temperature, int top_k,
                     const float* noise_expo_dev, uint64_t seed, uint32_t step, int64_t* out_tokens_dev, void* stream) {
    SHOWO_CHECK(logits_dev && out_tokens_dev && B > 0 && V > 0 && ld >= V, "mmu_sample: bad arguments");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    MmuSampleArgs ms{};
    ms.logits = logits_dev; ms.ld = ld; ms.B = B; ms.V = V; ms.temperature = temperature; ms.top_k = top_k;
    ms.noise_expo = noise_expo_dev; ms.seed = seed; ms.step = step; ms.out = out_tokens_dev; ms.out_stride = 1;
    return mmu_sample(ms, (cudaStream_t)stream);
}

int showo_mm_projector(showo_engine_t* e, const float* feats_dev, int64_t n, float* out_dev, void* stream) {
    SHOWO_CHECK(e && feats_dev && out_dev && n > 0, "mm_projector: bad arguments");
    SHOWO_TRY(engine_set_device(e));
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    for (const char* k : {"mm_projector.0.weight", "mm_projector.0.bias", "mm_projector.2.weight", "mm_projector.2.bias"})
        SHOWO_CHECK(e->loaded.count(k) == 1, std::string("mm_projector: weight not loaded: ") + k);
    constexpr int kIn = 1024, kMid = 2048, kOut = 2048;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > e->mmp_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        dev_free(e->mmp_in); dev_free(e->mmp_mid); dev_free(e->mmp_pre);
        SHOWO_TRY(dev_alloc(&e->mmp_in, (size_t)n * kIn));
        SHOWO_TRY(dev_alloc(&e->mmp_mid, (size_t)n * kMid));
        SHOWO_TRY(dev_alloc(&e->mmp_pre, (size_t)n * kMid));
        e->mmp_cap = n;
    }
    e->mmp_n = n;                                     // showo_mm_projector_backward differentiates THIS call
    SHOWO_TRY(f32_to_bf16(feats_dev, e->mmp_in, n * kIn, st));
    GemmArgs g0{};
    g0.A = e->mmp_in; g0.lda = kIn; g0.B = e->mmp_w0; g0.ldb = kIn; g0.M = (int)n; g0.N = kMid; g0.K = kIn;
    g0.out = e->mmp_pre; g0.ldc = kMid; g0.bias = e->mmp_b0; g0.gelu_from = kMid;          // no gelu_new here: nn.GELU() is the erf form
    g0.block_n = 128;                                 // keeps the M <= 16 case off the decode path's fp32-only epilogues
    SHOWO_TRY(gemm_bf16(g0, GEMM_BIAS_BF16, st));
    SHOWO_TRY(gelu_erf_bf16(e->mmp_pre, e->mmp_mid, n * kMid, st));      // the pre-activation stays for the backward
    GemmArgs g1{};
    g1.A = e->mmp_mid; g1.lda = kMid; g1.B = e->mmp_w2; g1.ldb = kMid; g1.M = (int)n; g1.N = kOut; g1.K = kMid;
    g1.out = out_dev; g1.ldc = kOut; g1.bias = e->mmp_b2; g1.block_n = 128;
    return gemm_bf16(g1, GEMM_BIAS_F32, st);
}

int showo_set_rng_row_base(showo_engine_t* e, int64_t first_row) {
    SHOWO_CHECK(e != nullptr && first_row >= 0 && first_row < (1 << 20), "set_rng_row_base: bad arguments");
    e->rng_row_base = (int)first_row;
    return 0;
}

int64_t showo_kernel_launches(showo_engine_t* e) { return e ? e->launches_last : 0; }

// ------------------------------------------------------------------------------------------------ raw kernel test entries
int showo_gemm_bf16(const void* A_dev, int64_t lda, const void* B_dev, int64_t ldb, int M, int N, int K, void* out_dev,
                    int64_t ldc, const float* bias_dev, const float* resid_dev, int64_t ldr, int gelu_from, int epi,
                    int block_n, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    GemmArgs g{};
    g.A = (const bf16*)A_dev; g.lda = lda; g.B = (const bf16*)B_dev; g.ldb = ldb; g.M = M; g.N = N; g.K = K;
    g.out = out_dev; g.ldc = ldc; g.bias = bias_dev; g.resid = resid_dev; g.ldr = ldr; g.gelu_from = gelu_from;
    g.block_n = block_n;
    if (epi == 3) return gemm_bf16_tn(g, (cudaStream_t)stream);        // token-major operands: C = A^T B, A = [K, lda], B = [K, ldb]
    SHOWO_CHECK(epi >= 0 && epi <= 2, "gemm: epi must be 0, 1, 2 or 3");
    return gemm_bf16(g, (GemmEpi)epi, (cudaStream_t)stream);
}

int showo_layernorm_test(const float* x_dev, const float* gamma_dev, const float* beta_dev, float eps, void* out_bf16_dev,
                         int rows, int D, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    return layernorm_bf16(x_dev, gamma_dev, beta_dev, eps, (bf16*)out_bf16_dev, rows, D, rows, rows, 0, (cudaStream_t)stream);
}

int showo_attention_test(void* qkv_dev, int64_t ld, int n_seq, int rows_per_seq, int pos0, int H,
                         const float* qg, const float* qb, const float* kg, const float* kb, float eps,
                         float rope_theta, int rotary_dim, void* kcache_dev, void* vtcache_dev, int Lmax, int n_keys,
                         const showo_seq_mask_t* masks_host, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(rotary_dim == 32, "rotary_dim must be 32");
    cudaStream_t st = (cudaStream_t)stream;
    const int D = H * 64;
    std::vector<float> c((size_t)Lmax * 32), s((size_t)Lmax * 32);
    for (int p = 0; p < Lmax; ++p)
        for (int i = 0; i < 16; ++i) {
            const float inv = 1.0f / powf(rope_theta, (float)(2 * i) / 32.0f);
            const float f = (float)p * inv;
            c[(size_t)p * 32 + i] = c[(size_t)p * 32 + 16 + i] = cosf(f);
            s[(size_t)p * 32 + i] = s[(size_t)p * 32 + 16 + i] = sinf(f);
        }
    float *dc = nullptr, *ds = nullptr; showo_seq_mask_t* dm = nullptr;
    SHOWO_TRY(dev_alloc(&dc, c.size())); SHOWO_TRY(dev_alloc(&ds, s.size())); SHOWO_TRY(dev_alloc(&dm, (size_t)n_seq));
    SHOWO_CUDA_OK(cudaMemcpyAsync(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
    SHOWO_CUDA_OK(cudaMemcpyAsync(ds, s.data(), s.size() * 4, cudaMemcpyHostToDevice, st));
    SHOWO_CUDA_OK(cudaMemcpyAsync(dm, masks_host, (size_t)n_seq * sizeof(showo_seq_mask_t), cudaMemcpyHostToDevice, st));
    QkRopeArgs r{};
    r.qkv = (bf16*)qkv_dev; r.ld = ld; r.n_rows = n_seq * rows_per_seq; r.rows_per_seq = rows_per_seq; r.pos0 = pos0;
    r.H = H; r.D = D; r.q_gamma = qg; r.q_beta = qb; r.k_gamma = kg; r.k_beta = kb; r.eps = eps;
    r.cos_tab = dc; r.sin_tab = ds; r.kcache = (bf16*)kcache_dev; r.vtcache = (bf16*)vtcache_dev; r.Lmax = Lmax;
    int rc = qk_norm_rope_scatter(r, st);
    if (!rc) {
        AttnArgs a{};
        a.q = (bf16*)qkv_dev + 2 * D; a.ld = ld; a.n_seq = n_seq; a.H = H; a.rows_per_seq = rows_per_seq; a.pos0 = pos0;
        a.kcache = (const bf16*)kcache_dev; a.vtcache = (const bf16*)vtcache_dev; a.Lmax = Lmax; a.n_keys = n_keys;
        a.masks = dm; a.scale = 0.125f;
        rc = (rows_per_seq == 1 && pos0 == n_keys - 1 && n_keys > 1) ? omni_attention_decode(a, st) : omni_attention(a, st);
    }
    cudaStreamSynchronize(st);
    cudaFree(dc); cudaFree(ds); cudaFree(dm);
    return rc;
}

int showo_attention_run(void* q_dev, int64_t ld, int n_seq, int rows_per_seq, int pos0, int H, const void* kcache_dev,
                        const void* vtcache_dev, int Lmax, int n_keys, const showo_seq_mask_t* masks_dev, void* out_dev,
                        int64_t out_ld, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(q_dev && kcache_dev && vtcache_dev && masks_dev, "attention_run: null argument");
    AttnArgs a{};
    a.q = (bf16*)q_dev; a.ld = ld; a.n_seq = n_seq; a.H = H; a.rows_per_seq = rows_per_seq; a.pos0 = pos0;
    a.kcache = (const bf16*)kcache_dev; a.vtcache = (const bf16*)vtcache_dev; a.Lmax = Lmax; a.n_keys = n_keys;
    a.masks = masks_dev; a.scale = 0.125f; a.out = (bf16*)out_dev; a.out_ld = out_ld;
    return omni_attention(a, (cudaStream_t)stream);
}

int showo_conv_test(const void* x_dev, const void* w_dev, const float* bias_dev, const void* resid_dev, void* out_dev,
                    int NB, int H, int W, int cin, int cout, int taps, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    ConvArgs a{};
    a.x = (const bf16*)x_dev; a.w = (const bf16*)w_dev; a.bias = bias_dev; a.resid = (const bf16*)resid_dev; a.ldr = cout;
    a.out = (bf16*)out_dev; a.ldc = cout; a.NB = NB; a.H = H; a.W = W; a.cin = cin; a.cout = cout; a.taps = taps;
    return conv_nhwc_bf16(a, (cudaStream_t)stream);
}

}  // extern "C"
