// fp32 verification forward of the Show-o backbone ("an fp32 / TF32-accumulate verification mode is needed for any
// stricter claim" than the bf16 tolerance).  A second, deliberately plain implementation of Showo.forward without labels
// (models/modeling_showo.py:76-79 -> models/phi.py:629-799,953-1208): fp32 activations, the fp32 MASTER weights the engine keeps when
// the optimizer is enabled (showo_optimizer_enable), CUDA cores only, one Linear at a time exactly as the reference writes them
// (six separate projections per layer, no fusion, no tensor cores, no bf16 anywhere), the omni mask as the same closed-form
// predicate.  Used by the parity tests to (a) pin the engine to the oracle at fp32 re-association level, where token decisions are
// bit-identical, and (b) measure the fast path's bf16 error against it.  Not a fallback: nothing on the product path calls it.
#include "attn_common.cuh"
#include "engine_state.h"

namespace showo {

// C[m][n] = bias[n] + sum_k A[m][k] W[n][k]      (A: [M, lda], W: [N, ldw] row-major, fp32), 64 x 64 tile, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256) vlinear_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                                                      const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K) {
    __shared__ float As[16][64 + 4], Ws[16][64 + 4];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, kk = i & 15;
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(int64_t)(m0 + r) * lda + k0 + kk] : 0.f;
            Ws[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(int64_t)(n0 + r) * ldw + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[(int64_t)m * ldc + n] = acc[i][j] + (bias ? bias[n] : 0.f);
        }
}
__global__ void __launch_bounds__(256) vembed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table, float* __restrict__ x, int D, int V) {
    int64_t id = ids[blockIdx.x];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    for (int d = threadIdx.x; d < D; d += 256) x[(int64_t)blockIdx.x * D + d] = table[id * D + d];
}
// LayerNorm over the last dimension (n elements per row), one CTA per row; rows may be the 64-wide head slices of a [M, D] buffer
__global__ void __launch_bounds__(64) vlayernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                        float* __restrict__ y, int n, float eps) {
    __shared__ float red[4];
    const float* r = x + (int64_t)blockIdx.x * n;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 64) s += r[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float mean = (red[0] + red[1]) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += 64) { const float d = r[i] - mean; q += d * d; }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[2 + (threadIdx.x >> 5)] = q;
    __syncthreads();
    const float rstd = rsqrtf((red[2] + red[3]) / (float)n + eps);
    for (int i = threadIdx.x; i < n; i += 64) y[(int64_t)blockIdx.x * n + i] = (r[i] - mean) * rstd * g[i] + b[i];
}
// partial rotary on the first 32 dims of every 64-wide head slice, rotate_half pairing (i, i + 16) (phi.py:163-196,680-694), in place
__global__ void vrotary_kernel(float* __restrict__ x, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, int64_t n_slices, int H, int L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slices * 16) return;
    const int j = (int)(i & 15);
    const int64_t slice = i >> 4;
    const int pos = (int)((slice / H) % L);
    float* p = x + slice * 64;
    const float a = p[j], b = p[j + 16];
    const float c0 = cos_tab[pos * 32 + j], s0 = sin_tab[pos * 32 + j], c1 = cos_tab[pos * 32 + 16 + j], s1 = sin_tab[pos * 32 + 16 + j];
    p[j] = a * c0 - b * s0;
    p[j + 16] = b * c1 + a * s1;
}
// softmax(q k^T / 8 + mask) v for one (sequence, head, query row) per CTA; q / k / v / out: [n_seq * L, H * 64] fp32
__global__ void __launch_bounds__(128) vattention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                         float* __restrict__ out, const showo_seq_mask_t* __restrict__ masks, int L, int H) {
    extern __shared__ float sc[];              // [L]
    __shared__ float qs[64];
    __shared__ float red[8];
    const int qi = blockIdx.x, h = blockIdx.y, s = blockIdx.z;
    const int64_t D = (int64_t)H * 64;
    const showo_seq_mask_t m = masks[s];
    if (threadIdx.x < 64) qs[threadIdx.x] = q[((int64_t)s * L + qi) * D + h * 64 + threadIdx.x];
    __syncthreads();
    float mx = -3.0e38f;
    for (int kk = threadIdx.x; kk < L; kk += 128) {
        float a = -3.0e38f;
        if (omni_allowed(m, qi, kk)) {
            const float* kr = k + ((int64_t)s * L + kk) * D + h * 64;
            a = 0.f;
            for (int d = 0; d < 64; ++d) a = fmaf(qs[d], kr[d], a);
            a *= 0.125f;
        }
        sc[kk] = a;
        mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int kk = threadIdx.x; kk < L; kk += 128) {
        const float p = sc[kk] > -1.0e38f ? expf(sc[kk] - mx) : 0.f;
        sc[kk] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[4 + (threadIdx.x >> 5)] = sum;
    __syncthreads();
    sum = red[4] + red[5] + red[6] + red[7];
    if (threadIdx.x < 64) {
        float acc = 0.f;
        for (int kk = 0; kk < L; ++kk) acc = fmaf(sc[kk], v[((int64_t)s * L + kk) * D + h * 64 + threadIdx.x], acc);
        out[((int64_t)s * L + qi) * D + h * 64 + threadIdx.x] = acc / sum;
    }
}
__global__ void vgelu_new_kernel(float* __restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        x[i] = 0.5f * v * (1.f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));
    }
}
__global__ void vadd3_kernel(float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = a[i] + b[i] + x[i];
}

}  // namespace showo

using namespace showo;

namespace {
struct Lin { float* w; float* b; int64_t rows, cols, ld; };
int linear_of(showo_engine* e, const std::string& prefix, Lin* out) {
    int64_t r, c, l;
    SHOWO_TRY(opt_master_slot(e, prefix + ".weight", &out->w, &out->rows, &out->cols, &out->ld));
    SHOWO_TRY(opt_master_slot(e, prefix + ".bias", &out->b, &r, &c, &l));
    return 0;
}
int vlinear(const float* A, int64_t lda, const Lin& w, float* C, int64_t ldc, int M, cudaStream_t st) {
    vlinear_kernel<<<dim3(cdiv((int)w.rows, 64), cdiv(M, 64)), 256, 0, st>>>(A, lda, w.w, w.ld, w.b, C, ldc, M, (int)w.rows, (int)w.cols);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
int egrid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 32); }
}  // namespace

extern "C" int showo_forward_fp32(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                                  const showo_seq_mask_t* masks_host, float* logits_out_dev, void* stream) {
    SHOWO_TRY(engine_check_ready(e));
    SHOWO_CHECK(e->opt != nullptr, "forward_fp32: the verification path reads the fp32 master weights -- call showo_optimizer_enable before loading them");
    SHOWO_CHECK((ids_dev != nullptr) != (embeds_dev != nullptr), "forward_fp32: exactly one of ids / embeds");
    SHOWO_CHECK(B > 0 && L > 0 && L <= e->cfg.max_pos && (size_t)L * 4 <= 48 * 1024 && masks_host && logits_out_dev, "forward_fp32: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    const int M = B * L, D = e->D, F = e->F, H = e->H, V = e->V;
    const size_t mD = (size_t)M * D;
    float *x = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *ao = nullptr, *mo = nullptr, *mid = nullptr;
    showo_seq_mask_t* dm = nullptr;
    int rc = 0;
    auto body = [&]() -> int {
        for (float** p : {&x, &xn, &q, &k, &v, &att, &ao, &mo}) SHOWO_TRY(dev_alloc(p, mD));
        SHOWO_TRY(dev_alloc(&mid, (size_t)M * F));
        SHOWO_TRY(dev_alloc(&dm, (size_t)B));
        SHOWO_CUDA_OK(cudaMemcpyAsync(dm, masks_host, (size_t)B * sizeof(showo_seq_mask_t), cudaMemcpyHostToDevice, st));
        float* embed; int64_t r, c, l;
        SHOWO_TRY(opt_master_slot(e, "showo.model.embed_tokens.weight", &embed, &r, &c, &l));
        if (ids_dev) { vembed_kernel<<<M, 256, 0, st>>>(ids_dev, embed, x, D, V); note_launch(); }
        else SHOWO_CUDA_OK(cudaMemcpyAsync(x, embeds_dev, mD * 4, cudaMemcpyDeviceToDevice, st));
        for (int li = 0; li < e->NL; ++li) {
            const std::string p = "showo.model.layers." + std::to_string(li) + ".";
            float *g, *bb, *qg, *qb, *kg, *kb;
            SHOWO_TRY(opt_master_slot(e, p + "input_layernorm.weight", &g, &r, &c, &l));
            SHOWO_TRY(opt_master_slot(e, p + "input_layernorm.bias", &bb, &r, &c, &l));
            SHOWO_TRY(opt_master_slot(e, p + "self_attn.q_layernorm.weight", &qg, &r, &c, &l));
            SHOWO_TRY(opt_master_slot(e, p + "self_attn.q_layernorm.bias", &qb, &r, &c, &l));
            SHOWO_TRY(opt_master_slot(e, p + "self_attn.k_layernorm.weight", &kg, &r, &c, &l));
            SHOWO_TRY(opt_master_slot(e, p + "self_attn.k_layernorm.bias", &kb, &r, &c, &l));
            Lin wq, wk, wv, wd, w1, w2;
            SHOWO_TRY(linear_of(e, p + "self_attn.q_proj", &wq)); SHOWO_TRY(linear_of(e, p + "self_attn.k_proj", &wk));
            SHOWO_TRY(linear_of(e, p + "self_attn.v_proj", &wv)); SHOWO_TRY(linear_of(e, p + "self_attn.dense", &wd));
            SHOWO_TRY(linear_of(e, p + "mlp.fc1", &w1)); SHOWO_TRY(linear_of(e, p + "mlp.fc2", &w2));
            vlayernorm_kernel<<<M, 64, 0, st>>>(x, g, bb, xn, D, e->cfg.ln_eps);                      // one shared pre-LN (phi.py:774-790)
            note_launch();
            SHOWO_TRY(vlinear(xn, D, wq, q, D, M, st));
            SHOWO_TRY(vlinear(xn, D, wk, k, D, M, st));
            SHOWO_TRY(vlinear(xn, D, wv, v, D, M, st));
            vlayernorm_kernel<<<M * H, 64, 0, st>>>(q, qg, qb, q, 64, e->cfg.ln_eps);                 // q / k LayerNorm(64), weights shared by the heads
            vlayernorm_kernel<<<M * H, 64, 0, st>>>(k, kg, kb, k, 64, e->cfg.ln_eps);
            const int64_t slices = (int64_t)M * H;
            vrotary_kernel<<<(int)((slices * 16 + 255) / 256), 256, 0, st>>>(q, e->cos_tab, e->sin_tab, slices, H, L);   // position = arange(L) for every row
            vrotary_kernel<<<(int)((slices * 16 + 255) / 256), 256, 0, st>>>(k, e->cos_tab, e->sin_tab, slices, H, L);
            vattention_kernel<<<dim3(L, H, B), 128, (size_t)L * 4, st>>>(q, k, v, att, dm, L, H);
            note_launch(5);
            SHOWO_TRY(vlinear(att, D, wd, ao, D, M, st));
            SHOWO_TRY(vlinear(xn, D, w1, mid, F, M, st));
            vgelu_new_kernel<<<egrid((int64_t)M * F), 256, 0, st>>>(mid, (int64_t)M * F);
            SHOWO_TRY(vlinear(mid, F, w2, mo, D, M, st));
            vadd3_kernel<<<egrid((int64_t)mD), 256, 0, st>>>(x, ao, mo, (int64_t)mD);                 // attn_out + mlp_out + x
            note_launch(2);
            SHOWO_CUDA_OK(cudaGetLastError());
        }
        float *fg, *fb;
        SHOWO_TRY(opt_master_slot(e, "showo.model.final_layernorm.weight", &fg, &r, &c, &l));
        SHOWO_TRY(opt_master_slot(e, "showo.model.final_layernorm.bias", &fb, &r, &c, &l));
        vlayernorm_kernel<<<M, 64, 0, st>>>(x, fg, fb, xn, D, e->cfg.ln_eps);
        note_launch();
        Lin head;
        SHOWO_TRY(linear_of(e, "showo.lm_head", &head));
        SHOWO_TRY(vlinear(xn, D, head, logits_out_dev, V, M, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));            // the scratch buffers are freed below
        return 0;
    };
    rc = body();
    for (float** p : {&x, &xn, &q, &k, &v, &att, &ao, &mo, &mid}) dev_free(*p);
    dev_free(dm);
    e->launches_last = launches_total() - l0;
    return rc;
}
