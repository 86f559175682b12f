// Memory-bound helper kernels of the backbone: LayerNorm, embedding gather, q/k LayerNorm + partial rotary + KV-cache
// scatter, fp32 -> bf16 weight packing.  All are coalesced, vectorised (16 B per thread per access) and sized so that
// every SM gets several CTAs.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace showo {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per output row; the row lives in registers between the two reductions (two-pass variance like ATen).
// phi.py:776 (input_layernorm), :1065 (final_layernorm): eps 1e-5, affine + bias, fp32 in, bf16 out (GEMM operand).
template <int kMaxVec>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             bf16* __restrict__ out, int n_rows_out, int D,
                                                             int rows_out_per_seq, int rows_in_per_seq, int row_off) {
    pdl_trigger();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    // gamma / beta are weights: pull them into L2 while the predecessor is still running (the decode step evicts them
    // every time: 2.9 GB of weights stream through a 126 MB L2), so the loads after the wait are L2 hits
    if (warp < n_rows_out) {
        for (int i = lane * 32; i < D; i += 32 * 32) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gamma + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(beta + i));
        }
    }
    pdl_wait();
    if (warp >= n_rows_out) return;
    const int64_t in_row = (int64_t)(warp / rows_out_per_seq) * rows_in_per_seq + row_off + warp % rows_out_per_seq;
    const float4* xr = reinterpret_cast<const float4*>(x + in_row * D);
    const int nvec = D >> 7;  // float4 per lane
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            v[i] = xr[i * 32 + lane];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + (int64_t)warp * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        if (i < nvec) {
            const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
            uint2 pk;
            pk.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            pk.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            orow[i * 32 + lane] = pk;
        }
    }
}

int layernorm_bf16(const float* x, const float* gamma, const float* beta, float eps, bf16* out, int n_rows_out, int D,
                   int rows_out_per_seq, int rows_in_per_seq, int row_off, cudaStream_t st) {
    SHOWO_CHECK(D % 128 == 0 && D <= 2048, "layernorm: D must be a multiple of 128 and <= 2048");
    if (n_rows_out == 0) return 0;
    const int wpb = 8;
    const int grid = cdiv(n_rows_out, wpb);
    if (D <= 512)
        SHOWO_CUDA_OK(launch_kernel(layernorm_bf16_kernel<4>, dim3(grid), dim3(wpb * 32), 0, st, 1, x, gamma, beta, eps, out, n_rows_out, D,
                                    rows_out_per_seq, rows_in_per_seq, row_off));
    else
        SHOWO_CUDA_OK(launch_kernel(layernorm_bf16_kernel<16>, dim3(grid), dim3(wpb * 32), 0, st, 1, x, gamma, beta, eps, out, n_rows_out, D,
                                    rows_out_per_seq, rows_in_per_seq, row_off));
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ embedding gather
// phi.py:1006 nn.Embedding: x[r,:] = table[id,:] (bf16 table -> fp32 residual stream)
__global__ void __launch_bounds__(128) embed_gather_kernel(const int64_t* __restrict__ ids, int64_t ids_stride, int pos0,
                                                           const bf16* __restrict__ table, float* __restrict__ x,
                                                           int n_rows, int rows_per_seq, int D, int vocab) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    int64_t id = ids[(int64_t)(r / rows_per_seq) * ids_stride + pos0 + r % rows_per_seq];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    const uint4* src = reinterpret_cast<const uint4*>(table + id * D);
    float4* dst = reinterpret_cast<float4*>(x + (int64_t)r * D);
    for (int i = threadIdx.x; i < D / 8; i += blockDim.x) {
        const uint4 u = __ldg(src + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]),
                     d = __bfloat1622float2(h[3]);
        dst[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        dst[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
}
int embed_gather(const int64_t* ids, int64_t ids_stride, int pos0, const bf16* table, float* x, int n_rows,
                 int rows_per_seq, int D, int vocab, cudaStream_t st) {
    if (n_rows == 0) return 0;
    SHOWO_CHECK(D % 8 == 0, "embed: D must be a multiple of 8");
    embed_gather_kernel<<<n_rows, 128, 0, st>>>(ids, ids_stride, pos0, table, x, n_rows, rows_per_seq, D, vocab);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ conversions
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
}
// nn.GELU() (exact, erf form) on bf16 -- the activation of Showo.mm_projector (modeling_showo.py:51), out of place:
// the pre-activation is kept for the backward
__global__ void gelu_erf_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __bfloat162float(src[i]);
        dst[i] = __float2bfloat16(0.5f * v * (1.f + erff(v * 0.70710678118654752f)));
    }
}
int gelu_erf_bf16(const bf16* src, bf16* dst, int64_t n, cudaStream_t st) {
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    gelu_erf_bf16_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(src, dst, n);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
// d/dv [v Phi(v)] = Phi(v) + v phi(v),  Phi(v) = (1 + erf(v / sqrt 2)) / 2,  phi(v) = exp(-v^2 / 2) / sqrt(2 pi)
__global__ void gelu_erf_bwd_bf16_kernel(bf16* __restrict__ d_io, const bf16* __restrict__ pre, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __bfloat162float(pre[i]);
        const float g = 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v);
        d_io[i] = __float2bfloat16(__bfloat162float(d_io[i]) * g);
    }
}
int gelu_erf_bwd_bf16(bf16* d_io, const bf16* pre, int64_t n, cudaStream_t st) {
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    gelu_erf_bwd_bf16_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(d_io, pre, n);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
// rows of a mixed ids / embeddings input: ids[r] < 0 marks a row whose vector comes from the caller (the mm_projector output of
// train_w_clip_vit.py:532-537), every other row keeps what embed_gather put there
__global__ void __launch_bounds__(128) embed_override_kernel(const int64_t* __restrict__ ids, const float* __restrict__ embeds,
                                                             float* __restrict__ x, int n_rows, int D) {
    const int r = blockIdx.x;
    if (r >= n_rows || ids[r] >= 0) return;
    const float4* s = reinterpret_cast<const float4*>(embeds + (int64_t)r * D);
    float4* d = reinterpret_cast<float4*>(x + (int64_t)r * D);
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) d[i] = s[i];
}
int embed_override(const int64_t* ids, const float* embeds, float* x, int n_rows, int D, cudaStream_t st) {
    SHOWO_CHECK(D % 4 == 0, "embed_override: D must be a multiple of 4");
    embed_override_kernel<<<n_rows, 128, 0, st>>>(ids, embeds, x, n_rows, D);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

int f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    f32_to_bf16_kernel<<<grid, 256, 0, st>>>(src, dst, n);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
__global__ void pack_block_kernel(const float* __restrict__ src, int64_t src_ld, bf16* __restrict__ dst, int64_t dst_ld,
                                  int rows, int cols) {
    const int64_t n = (int64_t)rows * cols;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t r = i / cols, c = i % cols;
        dst[r * dst_ld + c] = __float2bfloat16(src[r * src_ld + c]);
    }
}
int pack_block_bf16(const float* src, int64_t src_ld, bf16* dst, int64_t dst_ld, int rows, int cols, cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols;
    if (n == 0) return 0;
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    pack_block_kernel<<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
__global__ void copy_rows_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ dst, int64_t dst_ld,
                                 int rows, int cols) {
    const int64_t n = (int64_t)rows * cols;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t r = i / cols, c = i % cols;
        dst[r * dst_ld + c] = src[r * src_ld + c];
    }
}
int copy_f32_to_f32_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int rows, int cols,
                         cudaStream_t st) {
    const int64_t n = (int64_t)rows * cols;
    if (n == 0) return 0;
    int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    copy_rows_kernel<<<grid, 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, cols);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ q/k LN + rotary + scatter
// One CTA per (sequence, block of 32 positions, head).  Each warp normalises 4 rows of q and k (LayerNorm over the 64
// head dims, weights shared across heads, phi.py:264-271,665-667), rotates the first 32 dims with the rotate_half pairing
// (i, i+16) (phi.py:163-196,680-694) using position = pos0 + row, writes q back in place, K to the cache row-major and V
// to the cache TRANSPOSED ([64][Lmax], keys contiguous) through a padded smem tile so both stores are coalesced.
__global__ void __launch_bounds__(256) qk_norm_rope_scatter_kernel(QkRopeArgs a) {
    __shared__ bf16 vs[32][66];
    const int blocks_per_seq = (a.rows_per_seq + 31) >> 5;
    const int seq = blockIdx.x / blocks_per_seq;
    const int rb = blockIdx.x % blocks_per_seq;
    const int h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D;
    const float2 qg = reinterpret_cast<const float2*>(a.q_gamma)[lane], qb = reinterpret_cast<const float2*>(a.q_beta)[lane];
    const float2 kg = reinterpret_cast<const float2*>(a.k_gamma)[lane], kb = reinterpret_cast<const float2*>(a.k_beta)[lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = warp * 4 + i;            // row within the 32-row block
        const int rs = rb * 32 + rl;            // row within the sequence's row range
        if (rs >= a.rows_per_seq) continue;     // warp-uniform
        const int pos = a.pos0 + rs;
        bf16* row = a.qkv + ((int64_t)seq * a.rows_per_seq + rs) * a.ld + h * 64;
        const __nv_bfloat162 k2 = reinterpret_cast<const __nv_bfloat162*>(row)[lane];
        const __nv_bfloat162 v2 = reinterpret_cast<const __nv_bfloat162*>(row + D)[lane];
        const __nv_bfloat162 q2 = reinterpret_cast<const __nv_bfloat162*>(row + 2 * D)[lane];
        *reinterpret_cast<__nv_bfloat162*>(&vs[rl][2 * lane]) = v2;
        float2 kf = __bfloat1622float2(k2), qf = __bfloat1622float2(q2);
        // LayerNorm(64)
        {
            const float mk = warp_sum(kf.x + kf.y) * (1.f / 64.f), mq = warp_sum(qf.x + qf.y) * (1.f / 64.f);
            kf.x -= mk; kf.y -= mk; qf.x -= mq; qf.y -= mq;
            const float rk = rsqrtf(warp_sum(kf.x * kf.x + kf.y * kf.y) * (1.f / 64.f) + a.eps);
            const float rq = rsqrtf(warp_sum(qf.x * qf.x + qf.y * qf.y) * (1.f / 64.f) + a.eps);
            kf.x = kf.x * rk * kg.x + kb.x; kf.y = kf.y * rk * kg.y + kb.y;
            qf.x = qf.x * rq * qg.x + qb.x; qf.y = qf.y * rq * qg.y + qb.y;
        }
        // partial rotary on dims [0,32): lanes 0..7 hold dims 2l,2l+1 (< 16); partner dims (+16) sit in lane l+8
        {
            const float pkx = __shfl_xor_sync(0xffffffffu, kf.x, 8), pky = __shfl_xor_sync(0xffffffffu, kf.y, 8);
            const float pqx = __shfl_xor_sync(0xffffffffu, qf.x, 8), pqy = __shfl_xor_sync(0xffffffffu, qf.y, 8);
            if (lane < 16) {
                const float2 c = reinterpret_cast<const float2*>(a.cos_tab + (int64_t)pos * 32)[lane];
                const float2 s = reinterpret_cast<const float2*>(a.sin_tab + (int64_t)pos * 32)[lane];
                const float sgn = lane < 8 ? -1.f : 1.f;      // rotate_half: first half gets -x2, second half +x1
                kf.x = kf.x * c.x + sgn * pkx * s.x; kf.y = kf.y * c.y + sgn * pky * s.y;
                qf.x = qf.x * c.x + sgn * pqx * s.x; qf.y = qf.y * c.y + sgn * pqy * s.y;
            }
        }
        reinterpret_cast<__nv_bfloat162*>(row + 2 * D)[lane] = __floats2bfloat162_rn(qf.x, qf.y);
        bf16* kdst = a.kcache + (((int64_t)seq * a.H + h) * a.Lmax + pos) * 64;
        reinterpret_cast<__nv_bfloat162*>(kdst)[lane] = __floats2bfloat162_rn(kf.x, kf.y);
    }
    __syncthreads();
    // V^T scatter: a warp writes 32 consecutive positions of one head-dim row
    const int r = lane;
    const int rs = rb * 32 + r;
    if (rs < a.rows_per_seq) {
        bf16* vt = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax + a.pos0 + rs;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = warp + 8 * i;
            vt[(int64_t)d * a.Lmax] = vs[r][d];
        }
    }
}

int qk_norm_rope_scatter(const QkRopeArgs& a, cudaStream_t st) {
    if (a.n_rows == 0) return 0;
    SHOWO_CHECK(a.n_rows % a.rows_per_seq == 0, "qk_rope: n_rows must be a multiple of rows_per_seq");
    SHOWO_CHECK(a.pos0 + a.rows_per_seq <= a.Lmax, "qk_rope: positions exceed the KV cache length");
    const int n_seq = a.n_rows / a.rows_per_seq;
    dim3 grid(n_seq * cdiv(a.rows_per_seq, 32), a.H);
    qk_norm_rope_scatter_kernel<<<grid, 256, 0, st>>>(a);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ cross-entropy
// The three F.cross_entropy(ignore_index=-100) terms of Showo.forward (modeling_showo.py:81-100) over fp32 logits
// [n_seq, L, V]: row (b, t) of a term pairs logits[b0 + b, t0 + t, :] with labels[b0 + b, t0 + t + shift].  One CTA per row,
// one pass over the row (per-thread online max / sum-exp, merged across the block), per-row losses to a buffer and a
// single-CTA fixed-order reduction -> mean over the non-ignored rows (0/0 = NaN for an empty term, like torch).
struct CeRowsArgs {
    const float* logits; const int64_t* labels; int64_t L; int V;
    int b0, nb, t0, nt, shift; int64_t ignore_index;
    float* row_loss; float* row_valid;
};

__global__ void __launch_bounds__(256) ce_rows_kernel(CeRowsArgs a) {
    __shared__ float sm_m[8], sm_s[8];
    const int row = blockIdx.x, b = row / a.nt, t = row % a.nt;
    const int64_t label = a.labels[(int64_t)(a.b0 + b) * a.L + a.t0 + t + a.shift];
    const int tid = threadIdx.x;
    if (label == a.ignore_index) {                // CTA-uniform
        if (tid == 0) { a.row_loss[row] = 0.f; a.row_valid[row] = 0.f; }
        return;
    }
    const float* x = a.logits + ((int64_t)(a.b0 + b) * a.L + a.t0 + t) * (int64_t)a.V;
    float m = -FLT_MAX, sum = 0.f;
    for (int i = tid; i < a.V; i += 256) {
        const float v = x[i];
        if (v > m) { sum = sum * expf(m - v) + 1.f; m = v; }
        else sum += expf(v - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
        const float nm = fmaxf(m, om);
        sum = sum * expf(m - nm) + os * expf(om - nm);
        m = nm;
    }
    if ((tid & 31) == 0) { sm_m[tid >> 5] = m; sm_s[tid >> 5] = sum; }
    __syncthreads();
    if (tid == 0) {
        float M = sm_m[0], S = sm_s[0];
        for (int w = 1; w < 8; ++w) {
            const float nm = fmaxf(M, sm_m[w]);
            S = S * expf(M - nm) + sm_s[w] * expf(sm_m[w] - nm);
            M = nm;
        }
        const bool in_range = label >= 0 && label < a.V;
        a.row_loss[row] = in_range ? (M + logf(S)) - x[label] : __int_as_float(0x7fc00000);
        a.row_valid[row] = 1.f;
    }
}

__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float* row_loss, const float* row_valid, int n, float* out) {
    __shared__ float s_l[1024], s_v[1024];
    float l = 0.f, v = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) { l += row_loss[i]; v += row_valid[i]; }
    s_l[threadIdx.x] = l; s_v[threadIdx.x] = v;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { s_l[threadIdx.x] += s_l[threadIdx.x + o]; s_v[threadIdx.x] += s_v[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s_l[0] / s_v[0]; out[1] = s_v[0]; }
}

int cross_entropy_mean(const float* logits, const int64_t* labels, int64_t L, int V, int b0, int nb, int t0, int nt, int shift,
                       int64_t ignore_index, float* ws, float* out2, cudaStream_t st) {
    const int n = nb * nt;
    if (n <= 0) {                                 // empty slice: torch's mean over nothing is NaN
        const float h[2] = {__builtin_nanf(""), 0.f};
        SHOWO_CUDA_OK(cudaMemcpyAsync(out2, h, 8, cudaMemcpyHostToDevice, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    }
    CeRowsArgs a{logits, labels, L, V, b0, nb, t0, nt, shift, ignore_index, ws, ws + n};
    ce_rows_kernel<<<n, 256, 0, st>>>(a);
    ce_reduce_kernel<<<1, 1024, 0, st>>>(ws, ws + n, n, out2);
    note_launch(); note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace showo
