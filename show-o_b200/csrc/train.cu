// Training step of the Show-o backbone behind the C ABI: Showo.forward with labels under autograd + loss.backward()
// (models/modeling_showo.py:59-102, training/train.py:589-612; PhiDecoderLayer phi.py:774-790 differentiated).
//
// Forward (activations kept for the backward, per layer l, M = B * L token rows):
//   xs[l] fp32 residual stream -> LN (+ row mean / rstd) -> xh[l] bf16 -> GEMM1 (tcgen05, + bias) -> pre[l] = k | v | q | fc1 (raw, bf16)
//   -> q/k LayerNorm(64) + partial rotary -> qrot[l], krot[l] (+ the K / V^T tiles the forward attention kernel reads)
//   -> gelu_new(fc1) and omni attention (+ row log-sum-exp) -> a2[l] = attn | act -> GEMM2 (+ bias + residual) -> xs[l+1];
//   final LN -> head GEMM -> logits -> the three cross-entropy means.
// Backward, all contractions on the tcgen05 GEMM of gemm_tcgen05.cuh (bf16 operands, fp32 accumulate, fp32 gradients):
//   dgrad  dX = dY W        : A = dY [M, N] row-major, B = W^T kept as a second bf16 copy (w1t / w2t / head_wt), so both operands
//                             stay K-contiguous for TMA / UMMA (cost: 2.9 GB of HBM and one transpose per weight update);
//   wgrad  dW = dY^T A      : both operands transposed into [features, M] scratch (token dimension = K) by a tiled transpose;
//                             bias gradients are row sums of the transposed dY;
//   attention              : attention_bwd.cu (recompute S per tile, dK/dV by key block, dQ by query block);
//   elementwise            : d gelu_new, rotary^T + LayerNorm(64) backward, LayerNorm(D) backward (+ two-stage deterministic
//                             column reductions for all LayerNorm weights), softmax - onehot for the three loss terms, embedding
//                             scatter-add.
// Gradients land in one engine-owned fp32 buffer in the packed layout of the weights (W1 = [Wk; Wv; Wq; Wfc1], W2 = [Wdense | Wfc2])
// and are read back per reference parameter name with showo_read_grad.
#include <float.h>

#include "engine_state.h"

namespace showo {

// ------------------------------------------------------------------------------------------------ LayerNorm (training)
template <int kMaxVec>
__global__ void __launch_bounds__(256) layernorm_train_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, bf16* __restrict__ out,
                                                              float2* __restrict__ stats, int n_rows, int D) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)warp * D);
    const int nvec = D >> 7;
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) { v[i] = xr[i * 32 + lane]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) stats[warp] = make_float2(mean, rstd);
    uint2* orow = reinterpret_cast<uint2*>(out + (int64_t)warp * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            const float4 g = __ldg(g4 + i * 32 + lane), b = __ldg(b4 + i * 32 + lane);
            uint2 pk;
            pk.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            pk.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            orow[i * 32 + lane] = pk;
        }
}

// dx_io[r] (+)= rstd * (g o dy - mean(g o dy) - xhat * mean(g o dy o xhat))        (one warp per row)
template <int kMaxVec>
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                        float* __restrict__ dx_io, int accumulate, int n_rows, int D) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)warp * D);
    const float4* dr = reinterpret_cast<const float4*>(dy + (int64_t)warp * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float2 st = stats[warp];
    const int nvec = D >> 7;
    float4 xh[kMaxVec], gd[kMaxVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            const float4 xv = xr[i * 32 + lane], dv = dr[i * 32 + lane], g = __ldg(g4 + i * 32 + lane);
            xh[i] = make_float4((xv.x - st.x) * st.y, (xv.y - st.x) * st.y, (xv.z - st.x) * st.y, (xv.w - st.x) * st.y);
            gd[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
            s1 += (gd[i].x + gd[i].y) + (gd[i].z + gd[i].w);
            s2 += (gd[i].x * xh[i].x + gd[i].y * xh[i].y) + (gd[i].z * xh[i].z + gd[i].w * xh[i].w);
        }
    const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
    float4* orow = reinterpret_cast<float4*>(dx_io + (int64_t)warp * D);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
            float4 o = make_float4(st.y * (gd[i].x - m1 - xh[i].x * m2), st.y * (gd[i].y - m1 - xh[i].y * m2),
                                   st.y * (gd[i].z - m1 - xh[i].z * m2), st.y * (gd[i].w - m1 - xh[i].w * m2));
            if (accumulate) {
                const float4 p = orow[i * 32 + lane];
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            orow[i * 32 + lane] = o;
        }
}

// the same, one CTA of 256 threads per row (D <= 4096): a few registers per thread instead of a whole row per warp, so 8 CTAs / 64
// warps are resident per SM and the load, reduce and store phases of different rows overlap (the warp-per-row version at D = 2048 holds
// 149 registers, one 8-warp CTA per SM: 154 us per call for 303 MB of traffic)
__global__ void __launch_bounds__(256) ln_bwd_dx_row_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                            float* __restrict__ dx_io, int accumulate, int D) {
    __shared__ float rs[2][8];
    const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * D);
    const float4* dr = reinterpret_cast<const float4*>(dy + (int64_t)row * D);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    float4* orow = reinterpret_cast<float4*>(dx_io + (int64_t)row * D);
    const float2 st = stats[row];
    const int nvec = D >> 2;                  // float4 per row, <= 4 per thread
    float4 xh[4], gd[4], prev[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = tid + i * 256;
        if (v < nvec) {
            const float4 xv = xr[v], dv = dr[v], g = __ldg(g4 + v);
            if (accumulate) prev[i] = orow[v];
            xh[i] = make_float4((xv.x - st.x) * st.y, (xv.y - st.x) * st.y, (xv.z - st.x) * st.y, (xv.w - st.x) * st.y);
            gd[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
            s1 += (gd[i].x + gd[i].y) + (gd[i].z + gd[i].w);
            s2 += (gd[i].x * xh[i].x + gd[i].y * xh[i].y) + (gd[i].z * xh[i].z + gd[i].w * xh[i].w);
        }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { rs[0][warp] = s1; rs[1][warp] = s2; }
    __syncthreads();
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { m1 += rs[0][w]; m2 += rs[1][w]; }
    m1 /= (float)D; m2 /= (float)D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = tid + i * 256;
        if (v < nvec) {
            float4 o = make_float4(st.y * (gd[i].x - m1 - xh[i].x * m2), st.y * (gd[i].y - m1 - xh[i].y * m2),
                                   st.y * (gd[i].z - m1 - xh[i].z * m2), st.y * (gd[i].w - m1 - xh[i].w * m2));
            if (accumulate) { o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w; }
            orow[v] = o;
        }
    }
}

// partial[blockIdx.y][0][c] = sum_r dy[r,c] * xhat[r,c], partial[blockIdx.y][1][c] = sum_r dy[r,c] over this block's rows
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float2* __restrict__ stats, int n_rows, int D,
                                                           float* __restrict__ partial) {
    __shared__ float sg[8][33], sb[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float ag = 0.f, ab = 0.f;
    if (c < D)
        for (int r = blockIdx.y * 8 + ty; r < n_rows; r += 8 * gridDim.y) {
            const float2 st = stats[r];
            const float d = dy[(int64_t)r * D + c];
            ag += d * (x[(int64_t)r * D + c] - st.x) * st.y;
            ab += d;
        }
    sg[ty][tx] = ag; sb[ty][tx] = ab;
    __syncthreads();
    if (ty == 0 && c < D) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { ag += sg[j][tx]; ab += sb[j][tx]; }
        partial[((int64_t)blockIdx.y * 2 + 0) * D + c] = ag;
        partial[((int64_t)blockIdx.y * 2 + 1) * D + c] = ab;
    }
}
// out[i] = sum_s partial[s][i]   (fixed order -> deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int S, int n, float* __restrict__ out) {
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (i < n)
        for (int s = ty; s < S; s += 8) a += partial[(int64_t)s * n + i];
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && i < n) {
#pragma unroll
        for (int j = 1; j < 8; ++j) a += sm[j][tx];
        out[i] = a;
    }
}

// ------------------------------------------------------------------------------------------------ q/k LayerNorm(64) + rotary
struct QkTrainArgs {
    const bf16* pre; int64_t ld;          // [M, ld]: k at col 0, v at col D, q at col 2D (raw projections + bias)
    int n_seq, L, H, D;
    const float *qg, *qb, *kg, *kb; float eps;
    const float* cos_tab; const float* sin_tab;
    bf16* qrot; bf16* krot;               // [M, D] row-major
    bf16* kcache; bf16* vtcache; int Lmax;   // forward attention operands ([seq][H][Lmax][64], [seq][H][64][Lmax])
    // backward
    const bf16* dq; const bf16* dk;       // [M, D] gradients wrt the rotated q / k
    bf16* dpre;                           // [M, ld]: receives d k_raw at col 0 and d q_raw at col 2D
    float* partial;                       // [gridDim.x][4][64]: q gamma, q beta, k gamma, k beta
};

__global__ void __launch_bounds__(256) qk_rope_train_kernel(QkTrainArgs a) {
    __shared__ bf16 vs[32][66];
    const int blocks_per_seq = (a.L + 31) >> 5;
    const int seq = blockIdx.x / blocks_per_seq, rb = blockIdx.x % blocks_per_seq, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.D;
    const float2 qg = reinterpret_cast<const float2*>(a.qg)[lane], qb = reinterpret_cast<const float2*>(a.qb)[lane];
    const float2 kg = reinterpret_cast<const float2*>(a.kg)[lane], kb = reinterpret_cast<const float2*>(a.kb)[lane];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = warp * 4 + i, pos = rb * 32 + rl;
        if (pos >= a.L) continue;
        const int64_t m = (int64_t)seq * a.L + pos;
        const bf16* row = a.pre + m * a.ld + h * 64;
        const __nv_bfloat162 k2 = reinterpret_cast<const __nv_bfloat162*>(row)[lane];
        const __nv_bfloat162 v2 = reinterpret_cast<const __nv_bfloat162*>(row + D)[lane];
        const __nv_bfloat162 q2 = reinterpret_cast<const __nv_bfloat162*>(row + 2 * D)[lane];
        *reinterpret_cast<__nv_bfloat162*>(&vs[rl][2 * lane]) = v2;
        float2 kf = __bfloat1622float2(k2), qf = __bfloat1622float2(q2);
        const float mk = warp_sum(kf.x + kf.y) * (1.f / 64.f), mq = warp_sum(qf.x + qf.y) * (1.f / 64.f);
        kf.x -= mk; kf.y -= mk; qf.x -= mq; qf.y -= mq;
        const float rk = rsqrtf(warp_sum(kf.x * kf.x + kf.y * kf.y) * (1.f / 64.f) + a.eps);
        const float rq = rsqrtf(warp_sum(qf.x * qf.x + qf.y * qf.y) * (1.f / 64.f) + a.eps);
        kf.x = kf.x * rk * kg.x + kb.x; kf.y = kf.y * rk * kg.y + kb.y;
        qf.x = qf.x * rq * qg.x + qb.x; qf.y = qf.y * rq * qg.y + qb.y;
        const float pkx = __shfl_xor_sync(0xffffffffu, kf.x, 8), pky = __shfl_xor_sync(0xffffffffu, kf.y, 8);
        const float pqx = __shfl_xor_sync(0xffffffffu, qf.x, 8), pqy = __shfl_xor_sync(0xffffffffu, qf.y, 8);
        if (lane < 16) {
            const float2 c = reinterpret_cast<const float2*>(a.cos_tab + (int64_t)pos * 32)[lane];
            const float2 s = reinterpret_cast<const float2*>(a.sin_tab + (int64_t)pos * 32)[lane];
            const float sgn = lane < 8 ? -1.f : 1.f;
            kf.x = kf.x * c.x + sgn * pkx * s.x; kf.y = kf.y * c.y + sgn * pky * s.y;
            qf.x = qf.x * c.x + sgn * pqx * s.x; qf.y = qf.y * c.y + sgn * pqy * s.y;
        }
        const __nv_bfloat162 qo = __floats2bfloat162_rn(qf.x, qf.y), ko = __floats2bfloat162_rn(kf.x, kf.y);
        reinterpret_cast<__nv_bfloat162*>(a.qrot + m * D + h * 64)[lane] = qo;
        reinterpret_cast<__nv_bfloat162*>(a.krot + m * D + h * 64)[lane] = ko;
        reinterpret_cast<__nv_bfloat162*>(a.kcache + (((int64_t)seq * a.H + h) * a.Lmax + pos) * 64)[lane] = ko;
    }
    __syncthreads();
    const int pos = rb * 32 + lane;
    if (pos < a.L) {
        bf16* vt = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax + pos;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int d = warp + 8 * i;
            vt[(int64_t)d * a.Lmax] = vs[lane][d];
        }
    }
}

// backward of rotary (transpose of the rotation) and of LayerNorm(64).  Eight lanes per (position, head) -- lane l8 holds dims
// 8 l8 .. 8 l8 + 7 as one 16-byte vector -- so a warp covers four heads of a position per step, the 64-wide reductions are three
// shuffles, and the four 16-byte loads of a step are issued together (the first version, one warp per (position, head) with 4-byte
// loads behind five-shuffle reductions, was a latency chain: 307 us per layer at 8 x 1155 rows for 227 MB of traffic).
constexpr int kQkBwdPos = 16;            // positions per CTA (two per warp)
__global__ void __launch_bounds__(256) qk_rope_bwd_kernel(QkTrainArgs a) {
    __shared__ float red[8][256];
    const int blocks_per_seq = (a.L + kQkBwdPos - 1) / kQkBwdPos;
    const int seq = blockIdx.x / blocks_per_seq, rb = blockIdx.x % blocks_per_seq;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l8 = lane & 7, hs = lane >> 3;
    const int D = a.D;
    const bf16* __restrict__ pre = a.pre; const bf16* __restrict__ dqp = a.dq; const bf16* __restrict__ dkp = a.dk;
    bf16* __restrict__ dpre = a.dpre;
    float gq[8], gk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gq[j] = __ldg(a.qg + 8 * l8 + j); gk[j] = __ldg(a.kg + 8 * l8 + j); }
    float acc[4][8];                     // q gamma, q beta, k gamma, k beta of this lane's 8 dims
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    auto sum8 = [](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        return v;
    };
    for (int pp = 0; pp < kQkBwdPos / 8; ++pp) {
        const int pos = rb * kQkBwdPos + warp * (kQkBwdPos / 8) + pp;
        if (pos >= a.L) break;           // warp-uniform
        const int64_t m = (int64_t)seq * a.L + pos;
        // rotary dims [0, 32) = lanes l8 < 4; the partner of dim i is i ^ 16 = lane l8 ^ 2; cos / sin tables hold cat(freqs, freqs)
        float cs[8], sn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { cs[j] = 1.f; sn[j] = 0.f; }
        if (l8 < 4) {
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.cos_tab + (int64_t)pos * 32 + 8 * l8 + j));
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.sin_tab + (int64_t)pos * 32 + 8 * l8 + j));
                cs[j] = c4.x; cs[j + 1] = c4.y; cs[j + 2] = c4.z; cs[j + 3] = c4.w;
                sn[j] = s4.x; sn[j + 1] = s4.y; sn[j + 2] = s4.z; sn[j + 3] = s4.w;
            }
        }
        const float sgn = l8 < 2 ? 1.f : -1.f;               // transpose of the forward rotation
#pragma unroll 2
        for (int hb = 0; hb < a.H; hb += 4) {
            const bool hv = hb + hs < a.H;                    // H % 4 != 0: the spare head slots run on head H - 1 and drop the result
            const int h = hv ? hb + hs : a.H - 1;
            const int64_t oq = m * a.ld + 2 * D + h * 64 + 8 * l8, ok = m * a.ld + h * 64 + 8 * l8;
            uint4 u[4];
            u[0] = *reinterpret_cast<const uint4*>(dqp + m * D + h * 64 + 8 * l8);
            u[1] = *reinterpret_cast<const uint4*>(pre + oq);
            u[2] = *reinterpret_cast<const uint4*>(dkp + m * D + h * 64 + 8 * l8);
            u[3] = *reinterpret_cast<const uint4*>(pre + ok);
#pragma unroll
            for (int t = 0; t < 2; ++t) {                     // t = 0: q, t = 1: k
                float dz[8], xr[8];
                const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&u[2 * t]);
                const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&u[2 * t + 1]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gv = __bfloat1622float2(g2[j]), xv = __bfloat1622float2(x2[j]);
                    dz[2 * j] = gv.x; dz[2 * j + 1] = gv.y; xr[2 * j] = xv.x; xr[2 * j + 1] = xv.y;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float pr = __shfl_xor_sync(0xffffffffu, dz[j], 2);
                    dz[j] = dz[j] * cs[j] + sgn * pr * sn[j];          // lanes l8 >= 4: c = 1, s = 0
                }
                float sx = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) sx += xr[j];
                const float mu = sum8(sx) * (1.f / 64.f);
                float sv = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { xr[j] -= mu; sv = fmaf(xr[j], xr[j], sv); }
                const float rstd = rsqrtf(sum8(sv) * (1.f / 64.f) + a.eps);
                float s1 = 0.f, s2 = 0.f, gd[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xr[j] *= rstd;                                       // xhat
                    if (hv) {
                        acc[2 * t][j] = fmaf(dz[j], xr[j], acc[2 * t][j]);   // d gamma
                        acc[2 * t + 1][j] += dz[j];                          // d beta
                    }
                    gd[j] = dz[j] * (t == 0 ? gq[j] : gk[j]);
                    s1 += gd[j]; s2 = fmaf(gd[j], xr[j], s2);
                }
                const float m1 = sum8(s1) * (1.f / 64.f), m2 = sum8(s2) * (1.f / 64.f);
                uint4 o;
                o.x = pack_bf16(rstd * (gd[0] - m1 - xr[0] * m2), rstd * (gd[1] - m1 - xr[1] * m2));
                o.y = pack_bf16(rstd * (gd[2] - m1 - xr[2] * m2), rstd * (gd[3] - m1 - xr[3] * m2));
                o.z = pack_bf16(rstd * (gd[4] - m1 - xr[4] * m2), rstd * (gd[5] - m1 - xr[5] * m2));
                o.w = pack_bf16(rstd * (gd[6] - m1 - xr[6] * m2), rstd * (gd[7] - m1 - xr[7] * m2));
                if (hv) *reinterpret_cast<uint4*>(dpre + (t == 0 ? oq : ok)) = o;
            }
        }
    }
    // partial[block][tensor][dim], tensors q gamma, q beta, k gamma, k beta: the four head slots of a warp, then the eight warps
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = acc[t][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (hs == 0) red[warp][t * 64 + 8 * l8 + j] = v;
        }
    __syncthreads();
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    a.partial[(int64_t)blockIdx.x * 256 + threadIdx.x] = v;
}

// (gelu_new forward / backward live in the tcgen05 GEMM's epilogue: kernels.h GemmArgs::gelu_mode)
// dst[r, c] = src[r, c] for a column block (bf16, 16-byte vectors)
__global__ void __launch_bounds__(256) copy_cols_bf16_kernel(const bf16* __restrict__ src, int64_t ld_s, bf16* __restrict__ dst, int64_t ld_d,
                                                             int64_t rows, int cols) {
    const int cv = cols >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cv; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cv; const int c = (int)(i % cv) * 8;
        *reinterpret_cast<uint4*>(dst + r * ld_d + c) = *reinterpret_cast<const uint4*>(src + r * ld_s + c);
    }
}

// ------------------------------------------------------------------------------------------------ transposes / row sums
// dst[c, r] = bf16(src[r, c]) over 64 x 64 tiles (R, C even); optional row-major bf16 copy of the source (fp32 -> bf16)
template <class TSrc>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const TSrc* __restrict__ src, int64_t ld_s, int R, int C, int R_pad,
                                                                bf16* __restrict__ dst, int64_t ld_d, bf16* __restrict__ copy, int64_t ld_c) {
    __shared__ bf16 tile[64][66];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = r0 + ty + 8 * j, c = c0 + 2 * tx;
        __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
        if (r < R && c < C) {
            if constexpr (sizeof(TSrc) == 4) {
                const float2 f = *reinterpret_cast<const float2*>(src + (int64_t)r * ld_s + c);
                v = __floats2bfloat162_rn(f.x, f.y);
            } else {
                v = *reinterpret_cast<const __nv_bfloat162*>(src + (int64_t)r * ld_s + c);
            }
            if (copy != nullptr) *reinterpret_cast<__nv_bfloat162*>(copy + (int64_t)r * ld_c + c) = v;
        }
        *reinterpret_cast<__nv_bfloat162*>(&tile[ty + 8 * j][2 * tx]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + ty + 8 * j, r = r0 + 2 * tx;          // dst row c, dst columns r, r + 1 (zeros in the pad columns [R, R_pad))
        if (c < C && r < R_pad) {
            __nv_bfloat162 v;
            v.x = tile[2 * tx][ty + 8 * j];
            v.y = tile[2 * tx + 1][ty + 8 * j];
            *reinterpret_cast<__nv_bfloat162*>(dst + (int64_t)c * ld_d + r) = v;
        }
    }
}
// out[r] = sum_c src[r, c]  (one warp per row, fixed order)
__global__ void __launch_bounds__(256) rowsum_bf16_kernel(const bf16* __restrict__ src, int64_t ld, int R, int C, float* __restrict__ out) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= R) return;
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(src + (int64_t)r * ld);
    float s = 0.f;
    for (int i = lane; i < (C >> 1); i += 32) { const float2 f = __bfloat1622float2(p[i]); s += f.x + f.y; }
    if ((C & 1) && lane == 0) s += __bfloat162float(src[(int64_t)r * ld + C - 1]);
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}

// partial[blockIdx.y][c] = sum over this block's rows of src[r, c]  (c < C; the row stride ld is a multiple of 4 so the 8-byte loads of
// the last, partly valid column group stay inside the row); reduce_partials_kernel finishes in a fixed order
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ src, int64_t ld, int R, int C, float* __restrict__ partial) {
    __shared__ float sm[8][32][5];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + tx * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < C)
        for (int r = blockIdx.y * 8 + ty; r < R; r += 8 * gridDim.y) {
            const uint2 u = *reinterpret_cast<const uint2*>(src + (int64_t)r * ld + c);
            const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            a0 += f0.x; a1 += f0.y; a2 += f1.x; a3 += f1.y;
        }
    sm[ty][tx][0] = a0; sm[ty][tx][1] = a1; sm[ty][tx][2] = a2; sm[ty][tx][3] = a3;
    __syncthreads();
    if (ty < 4) {                      // thread (ty, tx) finishes column c + ty
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v += sm[j][tx][ty];
        if (c + ty < C) partial[(int64_t)blockIdx.y * C + c + ty] = v;
    }
}
__global__ void __launch_bounds__(256) train_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        uint2 o; o.x = pack_bf16(v.x, v.y); o.y = pack_bf16(v.z, v.w);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------ cross-entropy backward
struct CeBwdArgs {
    const float* logits; const int64_t* labels; int L, V; int64_t Vp;
    int b0[3], nb[3], t0[3], nt[3], shift[3];
    int64_t ignore_index;
    const float* loss_grads;      // [3] upstream gradients of the three mean losses (device)
    const float* counts;          // [3][2] {mean, count} written by the forward (device)
    bf16* dlogits;                // [B * L, Vp]
};
// dlogits[r, :] = sum_i w_i (softmax(logits[r]) - onehot(label_i)),  w_i = loss_grads[i] / count_i for every loss term i that
// counts row r (modeling_showo.py:81-100); rows no term counts get zeros.
__global__ void __launch_bounds__(256) ce_bwd_kernel(CeBwdArgs a) {
    __shared__ float sm_m[8], sm_s[8];
    __shared__ float s_lse;
    const int r = blockIdx.x, b = r / a.L, t = r % a.L, tid = threadIdx.x;
    float w[3]; int lab[3];
    float wsum = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        w[i] = 0.f; lab[i] = -1;
        if (b >= a.b0[i] && b < a.b0[i] + a.nb[i] && t >= a.t0[i] && t < a.t0[i] + a.nt[i]) {
            const int64_t l = a.labels[(int64_t)b * a.L + t + a.shift[i]];
            if (l != a.ignore_index) { w[i] = a.loss_grads[i] / a.counts[2 * i + 1]; lab[i] = (int)l; wsum += w[i]; }
        }
    }
    bf16* out = a.dlogits + (int64_t)r * a.Vp;
    if (lab[0] < 0 && lab[1] < 0 && lab[2] < 0) {
        for (int i = tid; i < a.Vp; i += 256) out[i] = __float2bfloat16(0.f);
        return;
    }
    const float* x = a.logits + (int64_t)r * a.V;
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
        for (int q = 1; q < 8; ++q) {
            const float nm = fmaxf(M, sm_m[q]);
            S = S * expf(M - nm) + sm_s[q] * expf(sm_m[q] - nm);
            M = nm;
        }
        s_lse = M + logf(S);
    }
    __syncthreads();
    const float lse = s_lse;
    for (int i = tid; i < a.Vp; i += 256) {
        float v = 0.f;
        if (i < a.V) {
            v = wsum * expf(x[i] - lse);
#pragma unroll
            for (int q = 0; q < 3; ++q) if (i == lab[q]) v -= w[q];
        }
        out[i] = __float2bfloat16(v);
    }
}

// g[ids[r], :] += dx[r, :]   (fp32 atomics; rows sharing a token id -- pads, the mask token -- collide on purpose)
__global__ void __launch_bounds__(128) embed_scatter_add_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dx,
                                                                float* __restrict__ g, int n_rows, int D, int vocab) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    int64_t id = ids[r];
    if (id < 0 || id >= vocab) return;
    for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(g + id * D + d, dx[(int64_t)r * D + d]);
}

// ================================================================================================ host side
struct TrainState {
    int cap_M = 0;
    int64_t Vp = 0;
    bf16 *w1t = nullptr, *w2t = nullptr, *head_wt = nullptr;
    int64_t wt_version = -1;
    float* xs = nullptr; float2* stats = nullptr; bf16* xh = nullptr; bf16* pre = nullptr; bf16* a2 = nullptr;
    bf16 *qrot = nullptr, *krot = nullptr; float* lse = nullptr;
    float *dx = nullptr, *dxh = nullptr; bf16 *dxb = nullptr, *dA2 = nullptr, *dpre = nullptr, *dqk = nullptr;
    float* delta = nullptr; bf16 *tA = nullptr, *tB = nullptr, *dlogits = nullptr;
    float* partials = nullptr; size_t partials_cap = 0;
    float* logits = nullptr; int64_t logits_cap = 0;
    float* grads = nullptr; int64_t n_grads = 0;
    int64_t* ids = nullptr; int64_t* labels = nullptr; float* losses = nullptr;   // device copies for the backward
    // what the last forward ran
    int B = 0, L = 0, M = 0; bool from_ids = false, have_forward = false;
    const float* logits_used = nullptr;
    int terms[15] = {0}; int64_t ignore_index = -100;
};

void train_state_destroy(TrainState* t) {
    if (!t) return;
    dev_free(t->w1t); dev_free(t->w2t); dev_free(t->head_wt);
    dev_free(t->xs); dev_free(t->stats); dev_free(t->xh); dev_free(t->pre); dev_free(t->a2); dev_free(t->qrot); dev_free(t->krot);
    dev_free(t->lse); dev_free(t->dx); dev_free(t->dxh); dev_free(t->dxb); dev_free(t->dA2); dev_free(t->dpre); dev_free(t->dqk);
    dev_free(t->delta); dev_free(t->tA); dev_free(t->tB); dev_free(t->dlogits); dev_free(t->partials); dev_free(t->logits);
    dev_free(t->grads); dev_free(t->ids); dev_free(t->labels); dev_free(t->losses);
    delete t;
}

// ---- gradient buffer layout (fp32, packed like the weights)
struct GradLayout {
    int64_t per_layer, w1, b1, w2, b2, ln_g, ln_b, qg, qb, kg, kb;      // offsets inside a layer block
    int64_t head_w, head_b, fln_g, fln_b, embed, total;
};
static GradLayout grad_layout(const showo_engine* e) {
    GradLayout g{};
    const int64_t D = e->D, W1N = e->W1N, W2K = e->W2K, V = e->V;
    int64_t o = 0;
    g.w1 = o; o += W1N * D; g.b1 = o; o += W1N; g.w2 = o; o += D * W2K; g.b2 = o; o += D;
    g.ln_g = o; o += D; g.ln_b = o; o += D; g.qg = o; o += 64; g.qb = o; o += 64; g.kg = o; o += 64; g.kb = o; o += 64;
    g.per_layer = o;
    o = g.per_layer * e->NL;
    g.head_w = o; o += V * D; g.head_b = o; o += (V + 3) / 4 * 4; g.fln_g = o; o += D; g.fln_b = o; o += D; g.embed = o; o += V * D;
    g.total = o;
    return g;
}

static int layernorm_train(const float* x, const float* g, const float* b, float eps, bf16* out, float2* stats, int rows, int D, cudaStream_t st) {
    const int grid = cdiv(rows, 8);
    if (D <= 512) layernorm_train_kernel<4><<<grid, 256, 0, st>>>(x, g, b, eps, out, stats, rows, D);
    else layernorm_train_kernel<16><<<grid, 256, 0, st>>>(x, g, b, eps, out, stats, rows, D);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int ensure_partials(TrainState* t, size_t n) {
    if (n > t->partials_cap) {
        dev_free(t->partials);
        SHOWO_TRY(dev_alloc(&t->partials, n));
        t->partials_cap = n;
    }
    return 0;
}
// LayerNorm(D) backward: dx_io (+)= ..., d gamma / d beta -> gg / gb
static int layernorm_backward(TrainState* t, const float* dy, const float* x, const float2* stats, const float* gamma, float* dx_io,
                              bool accumulate, float* gg, float* gb, int rows, int D, cudaStream_t st) {
    const int S = 32;
    SHOWO_TRY(ensure_partials(t, (size_t)S * 2 * D + 2 * D));
    ln_bwd_param_kernel<<<dim3(cdiv(D, 32), S), 256, 0, st>>>(dy, x, stats, rows, D, t->partials);
    // partial layout [S][2][D]: reduced as one vector of 2 * D (gamma | beta) behind the partials, then copied out
    reduce_partials_kernel<<<cdiv(2 * D, 32), 256, 0, st>>>(t->partials, S, 2 * D, t->partials + (size_t)S * 2 * D);
    SHOWO_CUDA_OK(cudaMemcpyAsync(gg, t->partials + (size_t)S * 2 * D, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
    SHOWO_CUDA_OK(cudaMemcpyAsync(gb, t->partials + (size_t)S * 2 * D + D, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
    const int grid = cdiv(rows, 8);
    if (D <= 512) ln_bwd_dx_kernel<4><<<grid, 256, 0, st>>>(dy, x, stats, gamma, dx_io, accumulate ? 1 : 0, rows, D);
    else if (D <= 4096 && D % 4 == 0) ln_bwd_dx_row_kernel<<<rows, 256, 0, st>>>(dy, x, stats, gamma, dx_io, accumulate ? 1 : 0, D);
    else ln_bwd_dx_kernel<16><<<grid, 256, 0, st>>>(dy, x, stats, gamma, dx_io, accumulate ? 1 : 0, rows, D);
    note_launch(3);
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
template <class TSrc>
static int transpose_to_bf16(const TSrc* src, int64_t ld_s, int R, int C, bf16* dst, int64_t ld_d, bf16* copy, int64_t ld_c, cudaStream_t st) {
    SHOWO_CHECK(C % 2 == 0 && ld_s % 2 == 0 && ld_d % 2 == 0, "transpose: even sizes expected");
    // the destination's row stride is the token count padded to the wgrad GEMM's 128-wide k block: the pad columns are written as zeros
    const int R_pad = (ld_d >= R && ld_d % 128 == 0 && ld_d - R < 128) ? (int)ld_d : R;
    transpose_to_bf16_kernel<TSrc><<<dim3(cdiv(C, 64), cdiv(R_pad, 64)), 256, 0, st>>>(src, ld_s, R, C, R_pad, dst, ld_d, copy, ld_c);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int rowsum_bf16(const bf16* src, int64_t ld, int R, int C, float* out, cudaStream_t st) {
    rowsum_bf16_kernel<<<cdiv(R, 8), 256, 0, st>>>(src, ld, R, C, out);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int gemm_plain(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, int M, int N, int K, void* out, int64_t ldc, bool f32_out,
                      cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.out = out; g.ldc = ldc; g.gelu_from = N;
    if (M <= 16) g.block_n = 64;          // keep the training path on the tcgen05 kernel whatever the row count
    return gemm_bf16(g, f32_out ? GEMM_BIAS_F32 : GEMM_BIAS_BF16, st);
}

// SHOWO_WGRAD_MN=0: the first version of the weight-gradient GEMMs (explicit transposed bf16 copies of dY and X, K-major operands)
static bool wgrad_mn() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_WGRAD_MN"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}
// dW[n_out, n_in] (fp32) = dY^T X over the M tokens and dbias[n_out] = column sums of dY, with dY = [M, ld_y], X = [M, ld_x] bf16.
// Default: the tcgen05 GEMM reads both operands as they lie (MN-major descriptors, gemm_bf16_tn) -- no transposed copies.
static int wgrad(TrainState* t, const bf16* dY, int64_t ld_y, int n_out, const bf16* X, int64_t ld_x, int n_in, int M, int64_t Mp,
                 float* dW, int64_t ld_w, float* dbias, cudaStream_t st, bool force_mn = false) {
    if (wgrad_mn() || force_mn) {
        GemmArgs g{};
        g.A = dY; g.lda = ld_y; g.B = X; g.ldb = ld_x; g.M = n_out; g.N = n_in; g.K = M; g.out = dW; g.ldc = ld_w;
        SHOWO_TRY(gemm_bf16_tn(g, st));
        const int S = 32;
        SHOWO_TRY(ensure_partials(t, (size_t)S * n_out));
        colsum_bf16_kernel<<<dim3(cdiv(n_out, 128), S), 256, 0, st>>>(dY, ld_y, M, n_out, t->partials);
        reduce_partials_kernel<<<cdiv(n_out, 32), 256, 0, st>>>(t->partials, S, n_out, dbias);
        note_launch(2);
        SHOWO_CUDA_OK(cudaGetLastError());
        return 0;
    }
    SHOWO_TRY(transpose_to_bf16<bf16>(dY, ld_y, M, (int)((n_out + 1) / 2 * 2), t->tA, Mp, nullptr, 0, st));
    SHOWO_TRY(transpose_to_bf16<bf16>(X, ld_x, M, n_in, t->tB, Mp, nullptr, 0, st));
    SHOWO_TRY(gemm_plain(t->tA, Mp, t->tB, Mp, n_out, n_in, (int)Mp, dW, ld_w, true, st));
    return rowsum_bf16(t->tA, Mp, n_out, M, dbias, st);
}

static int ensure_train(showo_engine* e, int M, cudaStream_t st) {
    if (!e->train) e->train = new TrainState();
    TrainState* t = e->train;
    const int64_t D = e->D, W1N = e->W1N, W2K = e->W2K, V = e->V, NL = e->NL, H = e->H;
    t->Vp = (V + 7) / 8 * 8;
    if (!t->w1t) {
        SHOWO_TRY(dev_alloc(&t->w1t, (size_t)(NL * D * W1N)));
        SHOWO_TRY(dev_alloc(&t->w2t, (size_t)(NL * W2K * D)));
        SHOWO_TRY(dev_alloc(&t->head_wt, (size_t)(D * t->Vp)));
        SHOWO_CUDA_OK(cudaMemsetAsync(t->head_wt, 0, (size_t)(D * t->Vp) * 2, st));
        const GradLayout gl = grad_layout(e);
        SHOWO_TRY(dev_alloc(&t->grads, (size_t)gl.total));
        t->n_grads = gl.total;
        SHOWO_TRY(dev_alloc(&t->losses, 8));
    }
    if (t->wt_version != e->weights_version) {
        for (int l = 0; l < NL; ++l) {
            SHOWO_TRY(transpose_to_bf16<bf16>(e->layers[l].w1, D, (int)W1N, (int)D, t->w1t + (size_t)l * D * W1N, W1N, nullptr, 0, st));
            SHOWO_TRY(transpose_to_bf16<bf16>(e->layers[l].w2, W2K, (int)D, (int)W2K, t->w2t + (size_t)l * W2K * D, D, nullptr, 0, st));
        }
        SHOWO_TRY(transpose_to_bf16<bf16>(e->head_w, D, (int)V, (int)D, t->head_wt, t->Vp, nullptr, 0, st));
        t->wt_version = e->weights_version;
    }
    if (M > t->cap_M) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        dev_free(t->xs); dev_free(t->stats); dev_free(t->xh); dev_free(t->pre); dev_free(t->a2); dev_free(t->qrot); dev_free(t->krot);
        dev_free(t->lse); dev_free(t->dx); dev_free(t->dxh); dev_free(t->dxb); dev_free(t->dA2); dev_free(t->dpre); dev_free(t->dqk);
        dev_free(t->delta); dev_free(t->tA); dev_free(t->tB); dev_free(t->dlogits); dev_free(t->ids); dev_free(t->labels);
        const size_t m = (size_t)M, Mp = (size_t)(M + 127) / 128 * 128;    // wgrad contracts over the tokens: K padded to the GEMM's 128-wide k block, pad columns zero
        SHOWO_TRY(dev_alloc(&t->xs, (size_t)(NL + 1) * m * D));
        SHOWO_TRY(dev_alloc(&t->stats, (size_t)(NL + 1) * m));
        SHOWO_TRY(dev_alloc(&t->xh, (size_t)(NL + 1) * m * D));
        SHOWO_TRY(dev_alloc(&t->pre, (size_t)NL * m * W1N));
        SHOWO_TRY(dev_alloc(&t->a2, (size_t)NL * m * W2K));
        SHOWO_TRY(dev_alloc(&t->qrot, (size_t)NL * m * D));
        SHOWO_TRY(dev_alloc(&t->krot, (size_t)NL * m * D));
        SHOWO_TRY(dev_alloc(&t->lse, (size_t)NL * m * H));
        SHOWO_TRY(dev_alloc(&t->dx, m * D));
        SHOWO_TRY(dev_alloc(&t->dxh, m * D));
        SHOWO_TRY(dev_alloc(&t->dxb, m * D));
        SHOWO_TRY(dev_alloc(&t->dA2, m * W2K));
        SHOWO_TRY(dev_alloc(&t->dpre, m * W1N));
        SHOWO_TRY(dev_alloc(&t->dqk, 2 * m * D));
        SHOWO_TRY(dev_alloc(&t->delta, m * H));
        const size_t ra = (size_t)(t->Vp > W1N ? t->Vp : W1N), rbm = (size_t)(W2K > D ? W2K : D);
        SHOWO_TRY(dev_alloc(&t->tA, ra * Mp));
        SHOWO_TRY(dev_alloc(&t->tB, rbm * Mp));
        SHOWO_CUDA_OK(cudaMemsetAsync(t->tA, 0, ra * Mp * sizeof(bf16), st));      // the transposes only ever write columns [0, M)
        SHOWO_CUDA_OK(cudaMemsetAsync(t->tB, 0, rbm * Mp * sizeof(bf16), st));
        SHOWO_TRY(dev_alloc(&t->dlogits, m * (size_t)t->Vp));
        SHOWO_TRY(dev_alloc(&t->ids, m));
        SHOWO_TRY(dev_alloc(&t->labels, m));
        t->cap_M = M;
    }
    return 0;
}

// location of a reference parameter inside a buffer in the packed layout (the gradient buffer, or the optimizer's master / moment
// buffers, which append one [NL][D] block: dense.bias and fc2.bias share a gradient but are two parameters)
static int named_slot(showo_engine* e, float* base_ptr, bool split_bias, const std::string& name, float** ptr, int64_t* rows, int64_t* cols,
                      int64_t* ld) {
    const GradLayout g = grad_layout(e);
    const int64_t D = e->D, F = e->F, V = e->V, W2K = e->W2K;
    auto set = [&](int64_t off, int64_t r, int64_t c, int64_t l) { *ptr = base_ptr + off; *rows = r; *cols = c; *ld = l; return 0; };
    if (name == "showo.model.embed_tokens.weight") return set(g.embed, V, D, D);
    if (name == "showo.lm_head.weight") return set(g.head_w, V, D, D);
    if (name == "showo.lm_head.bias") return set(g.head_b, 1, V, V);
    if (name == "showo.model.final_layernorm.weight") return set(g.fln_g, 1, D, D);
    if (name == "showo.model.final_layernorm.bias") return set(g.fln_b, 1, D, D);
    const std::string lp = "showo.model.layers.";
    SHOWO_CHECK(name.compare(0, lp.size(), lp) == 0, "read_grad: unknown parameter " + name);
    const size_t dot = name.find('.', lp.size());
    SHOWO_CHECK(dot != std::string::npos, "read_grad: bad parameter name " + name);
    const int l = atoi(name.substr(lp.size(), dot - lp.size()).c_str());
    SHOWO_CHECK(l >= 0 && l < e->NL, "read_grad: layer index out of range in " + name);
    const int64_t base = g.per_layer * l;
    const std::string k = name.substr(dot + 1);
    if (k == "self_attn.k_proj.weight") return set(base + g.w1, D, D, D);
    if (k == "self_attn.v_proj.weight") return set(base + g.w1 + D * D, D, D, D);
    if (k == "self_attn.q_proj.weight") return set(base + g.w1 + 2 * D * D, D, D, D);
    if (k == "mlp.fc1.weight") return set(base + g.w1 + 3 * D * D, F, D, D);
    if (k == "self_attn.k_proj.bias") return set(base + g.b1, 1, D, D);
    if (k == "self_attn.v_proj.bias") return set(base + g.b1 + D, 1, D, D);
    if (k == "self_attn.q_proj.bias") return set(base + g.b1 + 2 * D, 1, D, D);
    if (k == "mlp.fc1.bias") return set(base + g.b1 + 3 * D, 1, F, F);
    if (k == "self_attn.dense.weight") return set(base + g.w2, D, D, W2K);
    if (k == "mlp.fc2.weight") return set(base + g.w2 + D, D, F, W2K);
    if (k == "mlp.fc2.bias" && split_bias) return set(g.total + (int64_t)l * D, 1, D, D);
    if (k == "self_attn.dense.bias" || k == "mlp.fc2.bias") return set(base + g.b2, 1, D, D);
    if (k == "input_layernorm.weight") return set(base + g.ln_g, 1, D, D);
    if (k == "input_layernorm.bias") return set(base + g.ln_b, 1, D, D);
    if (k == "self_attn.q_layernorm.weight") return set(base + g.qg, 1, 64, 64);
    if (k == "self_attn.q_layernorm.bias") return set(base + g.qb, 1, 64, 64);
    if (k == "self_attn.k_layernorm.weight") return set(base + g.kg, 1, 64, 64);
    if (k == "self_attn.k_layernorm.bias") return set(base + g.kb, 1, 64, 64);
    SHOWO_CHECK(false, "read_grad: unknown parameter " + name);
    return -2;
}
static int named_grad(showo_engine* e, const std::string& name, float** ptr, int64_t* rows, int64_t* cols, int64_t* ld) {
    return named_slot(e, e->train->grads, false, name, ptr, rows, cols, ld);
}

// ================================================================================================ optimizer
// torch.optim.AdamW as the reference instantiates it (training/train.py:211-236, :617): decoupled weight decay on every parameter
// whose name has no "bias" (the reference's other no_decay patterns, "layer_norm.weight" / "embeddings.weight", match no Phi
// parameter name, so LayerNorm weights and the embedding ARE decayed), same update order as torch's single-tensor AdamW:
//   p *= 1 - lr * wd;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// One pass per tensor over gradient + fp32 master + both moments, writing the engine's bf16 (or fp32) working copy on the way out.
struct OptState {
    float *master = nullptr, *m = nullptr, *v = nullptr;
    int64_t n = 0, step = 0;
    // Showo.mm_projector's four tensors in the layout of showo_engine::mmp_grads (allocated when they are handed over)
    float *mmp_master = nullptr, *mmp_m = nullptr, *mmp_v = nullptr;
};
void opt_state_destroy(OptState* o) {
    if (!o) return;
    dev_free(o->master); dev_free(o->m); dev_free(o->v);
    dev_free(o->mmp_master); dev_free(o->mmp_m); dev_free(o->mmp_v);
    delete o;
}

// Showo.mm_projector = Linear(1024, 2048) -> GELU -> Linear(2048, 2048) (modeling_showo.py:49-54): element offsets of its four tensors
// inside the projector's gradient / master / moment buffers
constexpr int64_t kMmpIn = 1024, kMmpMid = 2048, kMmpOut = 2048;
constexpr int64_t kMmpW0 = 0, kMmpB0 = kMmpW0 + kMmpMid * kMmpIn, kMmpW2 = kMmpB0 + kMmpMid, kMmpB2 = kMmpW2 + kMmpOut * kMmpMid,
                  kMmpTotal = kMmpB2 + kMmpOut;
static bool is_mmp_name(const std::string& name) { return name.compare(0, 13, "mm_projector.") == 0; }
static int mmp_slot(float* base_ptr, const std::string& name, float** ptr, int64_t* rows, int64_t* cols) {
    SHOWO_CHECK(base_ptr != nullptr, "mm_projector: no buffer for " + name + " (weights not handed over / no backward has run)");
    if (name == "mm_projector.0.weight") { *ptr = base_ptr + kMmpW0; *rows = kMmpMid; *cols = kMmpIn; return 0; }
    if (name == "mm_projector.0.bias") { *ptr = base_ptr + kMmpB0; *rows = 1; *cols = kMmpMid; return 0; }
    if (name == "mm_projector.2.weight") { *ptr = base_ptr + kMmpW2; *rows = kMmpOut; *cols = kMmpMid; return 0; }
    if (name == "mm_projector.2.bias") { *ptr = base_ptr + kMmpB2; *rows = 1; *cols = kMmpOut; return 0; }
    SHOWO_CHECK(false, "unknown parameter " + name);
    return -2;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t rows, int64_t cols, int64_t ld, float lr, float b1, float b2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, bf16* __restrict__ out16,
                                                    float* __restrict__ out32, int64_t out_ld) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / cols, c = i % cols, o = r * ld + c;
        const float gi = g[o];
        float p = w[o] * (1.f - lr * wd);
        const float mi = m[o] + (gi - m[o]) * (1.f - b1);
        const float vi = v[o] * b2 + (1.f - b2) * gi * gi;
        p -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        w[o] = p; m[o] = mi; v[o] = vi;
        if (out16) out16[r * out_ld + c] = __float2bfloat16(p);
        if (out32) out32[r * out_ld + c] = p;
    }
}

int opt_store_master(showo_engine* e, const std::string& name, const float* src_dev, int64_t numel, cudaStream_t st) {
    float* dst; int64_t rows, cols, ld;
    if (is_mmp_name(name)) {
        OptState* o = e->opt;
        if (!o->mmp_master) {
            SHOWO_TRY(dev_alloc(&o->mmp_master, (size_t)kMmpTotal)); SHOWO_TRY(dev_alloc(&o->mmp_m, (size_t)kMmpTotal));
            SHOWO_TRY(dev_alloc(&o->mmp_v, (size_t)kMmpTotal));
            SHOWO_CUDA_OK(cudaMemsetAsync(o->mmp_m, 0, (size_t)kMmpTotal * 4, st));
            SHOWO_CUDA_OK(cudaMemsetAsync(o->mmp_v, 0, (size_t)kMmpTotal * 4, st));
        }
        SHOWO_TRY(mmp_slot(o->mmp_master, name, &dst, &rows, &cols));
        SHOWO_CHECK(rows * cols == numel, "optimizer: parameter " + name + " has the wrong size");
        SHOWO_CUDA_OK(cudaMemcpyAsync(dst, src_dev, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    SHOWO_TRY(named_slot(e, e->opt->master, true, name, &dst, &rows, &cols, &ld));
    SHOWO_CHECK(rows * cols == numel, "optimizer: parameter " + name + " has the wrong size");
    SHOWO_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)ld * 4, src_dev, (size_t)cols * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int opt_master_slot(showo_engine* e, const std::string& name, float** ptr, int64_t* rows, int64_t* cols, int64_t* ld) {
    SHOWO_CHECK(e && e->opt && e->opt->master, "no optimizer state: call showo_optimizer_enable before loading the weights");
    return named_slot(e, e->opt->master, true, name, ptr, rows, cols, ld);
}

}  // namespace showo

using namespace showo;

extern "C" {

int showo_train_forward(showo_engine_t* e, const int64_t* ids_dev, const float* embeds_dev, int B, int L,
                        const showo_seq_mask_t* masks_host, const int64_t* labels_dev, const int32_t* terms_host,
                        int64_t ignore_index, float* logits_out_dev, float* losses_out_dev, void* stream) {
    SHOWO_TRY(engine_check_ready(e));
    SHOWO_CHECK(ids_dev != nullptr || embeds_dev != nullptr, "train_forward: needs ids, embeds or both");
    SHOWO_CHECK(B > 0 && L > 0 && L <= e->cfg.max_pos && masks_host && labels_dev && terms_host && losses_out_dev,
                "train_forward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    const int M = B * L, D = e->D, H = e->H, W1N = e->W1N, W2K = e->W2K, V = e->V, NL = e->NL;
    SHOWO_TRY(engine_ensure_ws(e, 0, B, L, 0, st));         // K / V^T tiles of one layer + the mask descriptors
    SHOWO_TRY(engine_upload_masks(e, masks_host, B, st));
    SHOWO_TRY(ensure_train(e, M, st));
    TrainState* t = e->train;
    t->B = B; t->L = L; t->M = M; t->from_ids = ids_dev != nullptr; t->ignore_index = ignore_index;
    for (int i = 0; i < 15; ++i) t->terms[i] = terms_host[i];
    SHOWO_CUDA_OK(cudaMemcpyAsync(t->labels, labels_dev, (size_t)M * 8, cudaMemcpyDeviceToDevice, st));
    const size_t mD = (size_t)M * D;
    if (ids_dev) {
        SHOWO_CUDA_OK(cudaMemcpyAsync(t->ids, ids_dev, (size_t)M * 8, cudaMemcpyDeviceToDevice, st));
        SHOWO_TRY(embed_gather(ids_dev, L, 0, e->embed, t->xs, M, L, D, V, st));
        // both given = mixed input (train_w_clip_vit.py:532-537): rows with ids < 0 take the caller's vector (mm_projector output);
        // the backward's embedding scatter-add skips them and dembeds carries their gradient
        if (embeds_dev) SHOWO_TRY(embed_override(ids_dev, embeds_dev, t->xs, M, D, st));
    } else {
        SHOWO_CUDA_OK(cudaMemcpyAsync(t->xs, embeds_dev, mD * 4, cudaMemcpyDeviceToDevice, st));
    }
    for (int l = 0; l < NL; ++l) {
        const LayerW& w = e->layers[l];
        float* x = t->xs + (size_t)l * mD;
        bf16* xh = t->xh + (size_t)l * mD;
        bf16* pre = t->pre + (size_t)l * M * W1N;
        bf16* a2 = t->a2 + (size_t)l * M * W2K;
        SHOWO_TRY(layernorm_train(x, w.ln_g, w.ln_b, e->cfg.ln_eps, xh, t->stats + (size_t)l * M, M, D, st));
        GemmArgs g1{};
        g1.A = xh; g1.lda = D; g1.B = w.w1; g1.ldb = D; g1.M = M; g1.N = W1N; g1.K = D; g1.out = pre; g1.ldc = W1N; g1.bias = w.b1;
        // the fc1 columns are kept raw (the backward's gelu' reads them) and written as gelu_new(.) into the second GEMM's operand
        g1.gelu_from = 3 * (int)D; g1.gelu_mode = 1; g1.gelu_out = a2 + D; g1.gelu_out_ld = W2K;
        if (M <= 16) g1.block_n = 64;
        SHOWO_TRY(gemm_bf16(g1, GEMM_BIAS_BF16, st));
        QkTrainArgs q{};
        q.pre = pre; q.ld = W1N; q.n_seq = B; q.L = L; q.H = H; q.D = D; q.qg = w.qg; q.qb = w.qb; q.kg = w.kg; q.kb = w.kb;
        q.eps = e->cfg.ln_eps; q.cos_tab = e->cos_tab; q.sin_tab = e->sin_tab;
        q.qrot = t->qrot + (size_t)l * mD; q.krot = t->krot + (size_t)l * mD; q.kcache = e->kcache; q.vtcache = e->vtcache; q.Lmax = e->cap_L;
        qk_rope_train_kernel<<<dim3(B * cdiv(L, 32), H), 256, 0, st>>>(q);
        note_launch(1);
        SHOWO_CUDA_OK(cudaGetLastError());
        AttnArgs a{};
        a.q = t->qrot + (size_t)l * mD; a.ld = D; a.n_seq = B; a.H = H; a.rows_per_seq = L; a.pos0 = 0;
        a.kcache = e->kcache; a.vtcache = e->vtcache; a.Lmax = e->cap_L; a.n_keys = L; a.masks = e->d_masks; a.scale = 0.125f; a.work_ctr = e->attn_ctr;
        a.out = a2; a.out_ld = W2K; a.lse = t->lse + (size_t)l * M * H;
        SHOWO_TRY(omni_attention(a, st));
        GemmArgs g2{};
        g2.A = a2; g2.lda = W2K; g2.B = w.w2; g2.ldb = W2K; g2.M = M; g2.N = D; g2.K = W2K;
        g2.out = t->xs + (size_t)(l + 1) * mD; g2.ldc = D; g2.bias = w.b2; g2.resid = x; g2.ldr = D; if (M <= 16) g2.block_n = 64;
        SHOWO_TRY(gemm_bf16(g2, GEMM_RESID_F32, st));
    }
    SHOWO_TRY(layernorm_train(t->xs + (size_t)NL * mD, e->fln_g, e->fln_b, e->cfg.ln_eps, t->xh + (size_t)NL * mD, t->stats + (size_t)NL * M, M, D, st));
    float* logits = logits_out_dev;
    if (!logits) {
        if ((int64_t)M * V > t->logits_cap) {
            SHOWO_CUDA_OK(cudaStreamSynchronize(st));
            dev_free(t->logits);
            SHOWO_TRY(dev_alloc(&t->logits, (size_t)M * V));
            t->logits_cap = (int64_t)M * V;
        }
        logits = t->logits;
    }
    t->logits_used = logits;
    GemmArgs gh{};
    gh.A = t->xh + (size_t)NL * mD; gh.lda = D; gh.B = e->head_w; gh.ldb = D; gh.M = M; gh.N = V; gh.K = D;
    gh.out = logits; gh.ldc = V; gh.bias = e->head_b; if (M <= 16) gh.block_n = 64;
    SHOWO_TRY(gemm_bf16(gh, GEMM_BIAS_F32, st));
    // the three cross-entropy means (modeling_showo.py:81-100); {mean, count} pairs are kept on the device for the backward
    SHOWO_TRY(ensure_partials(t, (size_t)2 * M + 64));
    for (int i = 0; i < 3; ++i) {
        const int* tm = t->terms + 5 * i;
        SHOWO_TRY(cross_entropy_mean(logits, labels_dev, L, V, tm[0], tm[1], tm[2], tm[3], tm[4], ignore_index, t->partials, t->losses + 2 * i, st));
    }
    SHOWO_CUDA_OK(cudaMemcpyAsync(losses_out_dev, t->losses, 6 * 4, cudaMemcpyDeviceToDevice, st));
    t->have_forward = true;
    e->launches_last = launches_total() - l0;
    return 0;
}

// phase -1: loss -> dlogits -> head + final LayerNorm;  phase l in [0, NL): decoder layer l (in DEcreasing order);  phase -2: embedding
static int backward_phase(showo_engine_t* e, int phase, const float* loss_grads_dev, float* dembeds_out_dev, cudaStream_t st) {
    TrainState* t = e->train;
    const int M = t->M, B = t->B, L = t->L, D = e->D, H = e->H, W1N = e->W1N, W2K = e->W2K, V = e->V, NL = e->NL;
    const int64_t Mp = (int64_t)(M + 127) / 128 * 128, Vp = t->Vp;
    const size_t mD = (size_t)M * D;
    const GradLayout gl = grad_layout(e);
    float* G = t->grads;
    if (phase == -1) {
    // ---- loss -> dlogits (bf16) -> head
    CeBwdArgs c{};
    c.logits = t->logits_used; c.labels = t->labels; c.L = L; c.V = V; c.Vp = Vp; c.ignore_index = t->ignore_index;
    for (int i = 0; i < 3; ++i) { c.b0[i] = t->terms[5 * i]; c.nb[i] = t->terms[5 * i + 1]; c.t0[i] = t->terms[5 * i + 2]; c.nt[i] = t->terms[5 * i + 3]; c.shift[i] = t->terms[5 * i + 4]; }
    c.loss_grads = loss_grads_dev; c.counts = t->losses; c.dlogits = t->dlogits;
    ce_bwd_kernel<<<M, 256, 0, st>>>(c);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    // d xh_f = dlogits * Wh                                     [M, D] fp32
    SHOWO_TRY(gemm_plain(t->dlogits, Vp, t->head_wt, Vp, M, D, (int)Vp, t->dxh, D, true, st));
    // d Wh = dlogits^T * xh_f, d bh = column sums of dlogits
    SHOWO_TRY(wgrad(t, t->dlogits, Vp, V, t->xh + (size_t)NL * mD, D, D, M, Mp, G + gl.head_w, D, G + gl.head_b, st));
    // final LayerNorm
    SHOWO_TRY(layernorm_backward(t, t->dxh, t->xs + (size_t)NL * mD, t->stats + (size_t)NL * M, e->fln_g, t->dx, false, G + gl.fln_g,
                                 G + gl.fln_b, M, D, st));
        return 0;
    }
    if (phase >= 0) {
        const int l = phase;
        const LayerW& w = e->layers[l];
        float* GL = G + gl.per_layer * l;
        const bf16* pre = t->pre + (size_t)l * M * W1N;
        const bf16* a2 = t->a2 + (size_t)l * M * W2K;
        // dx (fp32) -> bf16 row-major + transposed; d b2 = column sums
        train_f32_to_bf16_kernel<<<148 * 8, 256, 0, st>>>(t->dx, t->dxb, (int64_t)mD / 4);
        note_launch();
        // d W2 = dx^T * [attn | act]
        SHOWO_TRY(wgrad(t, t->dxb, D, D, a2, W2K, W2K, M, Mp, GL + gl.w2, W2K, GL + gl.b2, st));
        // d [attn | act] = dx * W2; the act columns leave the epilogue as d fc1 = d act * gelu_new'(fc1) in dpre[:, 3D:]
        {
            GemmArgs g{};
            g.A = t->dxb; g.lda = D; g.B = t->w2t + (size_t)l * W2K * D; g.ldb = D; g.M = M; g.N = (int)W2K; g.K = D; g.out = t->dA2; g.ldc = W2K;
            g.gelu_from = D; g.gelu_mode = 2; g.gelu_out = t->dpre + 3 * D; g.gelu_out_ld = W1N; g.gelu_pre = pre + 3 * D; g.gelu_pre_ld = W1N;
            if (M <= 16) g.block_n = 64;
            SHOWO_TRY(gemm_bf16(g, GEMM_BIAS_BF16, st));
        }
        // attention backward: d q_rot, d k_rot -> dqk, d v -> dpre[:, D:2D]
        AttnBwdArgs ab{};
        ab.q = t->qrot + (size_t)l * mD; ab.q_ld = D; ab.k = t->krot + (size_t)l * mD; ab.k_ld = D; ab.v = pre + D; ab.v_ld = W1N;
        ab.o = a2; ab.o_ld = W2K; ab.d_o = t->dA2; ab.do_ld = W2K; ab.lse = t->lse + (size_t)l * M * H; ab.delta = t->delta;
        ab.dq = t->dqk; ab.dq_ld = D; ab.dk = t->dqk + mD; ab.dk_ld = D; ab.dv = t->dpre + D; ab.dv_ld = W1N;
        ab.n_seq = B; ab.H = H; ab.L = L; ab.masks = e->d_masks; ab.scale = 0.125f;
        SHOWO_TRY(omni_attention_backward(ab, st));
        // rotary^T + LayerNorm(64) backward -> d q_raw / d k_raw; gelu' -> d fc1
        QkTrainArgs q{};
        q.pre = pre; q.ld = W1N; q.n_seq = B; q.L = L; q.H = H; q.D = D; q.qg = w.qg; q.qb = w.qb; q.kg = w.kg; q.kb = w.kb;
        q.eps = e->cfg.ln_eps; q.cos_tab = e->cos_tab; q.sin_tab = e->sin_tab; q.dq = t->dqk; q.dk = t->dqk + mD; q.dpre = t->dpre;
        const int nblk = B * cdiv(L, kQkBwdPos);
        SHOWO_TRY(ensure_partials(t, (size_t)nblk * 256 + 256));
        q.partial = t->partials;
        qk_rope_bwd_kernel<<<nblk, 256, 0, st>>>(q);
        reduce_partials_kernel<<<cdiv(256, 32), 256, 0, st>>>(t->partials, nblk, 256, GL + gl.qg);      // qg | qb | kg | kb are adjacent
        note_launch(2);
        SHOWO_CUDA_OK(cudaGetLastError());
        // d W1 = dpre^T * xh, d b1 = column sums of dpre
        SHOWO_TRY(wgrad(t, t->dpre, W1N, W1N, t->xh + (size_t)l * mD, D, D, M, Mp, GL + gl.w1, D, GL + gl.b1, st));
        // d xh = dpre * W1, then LayerNorm backward into the residual gradient
        SHOWO_TRY(gemm_plain(t->dpre, W1N, t->w1t + (size_t)l * D * W1N, W1N, M, D, W1N, t->dxh, D, true, st));
        SHOWO_TRY(layernorm_backward(t, t->dxh, t->xs + (size_t)l * mD, t->stats + (size_t)l * M, w.ln_g, t->dx, true, GL + gl.ln_g,
                                     GL + gl.ln_b, M, D, st));
        return 0;
    }
    if (t->from_ids) {
        SHOWO_CUDA_OK(cudaMemsetAsync(G + gl.embed, 0, (size_t)V * D * 4, st));
        embed_scatter_add_kernel<<<M, 128, 0, st>>>(t->ids, t->dx, G + gl.embed, M, D, V);
        note_launch();
        SHOWO_CUDA_OK(cudaGetLastError());
    } else {
        SHOWO_CUDA_OK(cudaMemsetAsync(G + gl.embed, 0, (size_t)V * D * 4, st));
    }
    if (dembeds_out_dev) SHOWO_CUDA_OK(cudaMemcpyAsync(dembeds_out_dev, t->dx, mD * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int backward_check(showo_engine_t* e, const float* loss_grads_dev) {
    SHOWO_TRY(engine_check_ready(e));
    SHOWO_CHECK(e->train && e->train->have_forward, "backward: call showo_train_forward first");
    SHOWO_CHECK(loss_grads_dev != nullptr, "backward: null loss gradients");
    return 0;
}

int showo_backward(showo_engine_t* e, const float* loss_grads_dev, float* dembeds_out_dev, void* stream) {
    SHOWO_TRY(backward_check(e, loss_grads_dev));
    const int64_t l0 = launches_total();
    SHOWO_TRY(backward_phase(e, -1, loss_grads_dev, nullptr, (cudaStream_t)stream));
    for (int l = e->NL - 1; l >= 0; --l) SHOWO_TRY(backward_phase(e, l, loss_grads_dev, nullptr, (cudaStream_t)stream));
    SHOWO_TRY(backward_phase(e, -2, loss_grads_dev, dembeds_out_dev, (cudaStream_t)stream));
    e->launches_last = launches_total() - l0;
    return 0;
}

int showo_backward_phase(showo_engine_t* e, int phase, const float* loss_grads_dev, float* dembeds_out_dev, void* stream) {
    SHOWO_TRY(backward_check(e, loss_grads_dev));
    SHOWO_CHECK(phase >= -2 && phase < e->NL, "backward_phase: phase must be -1 (head), a layer index, or -2 (embedding)");
    return backward_phase(e, phase, loss_grads_dev, dembeds_out_dev, (cudaStream_t)stream);
}

int showo_grad_buffer(showo_engine_t* e, float** base_dev, int64_t* numel) {
    SHOWO_CHECK(e && e->train && e->train->grads && base_dev && numel, "grad_buffer: no training state (call showo_train_forward first)");
    *base_dev = e->train->grads;
    *numel = grad_layout(e).total;
    return 0;
}

int showo_grad_range(showo_engine_t* e, int phase, int64_t* begin, int64_t* end) {
    SHOWO_CHECK(e && begin && end && phase >= -2 && phase < e->NL, "grad_range: bad arguments");
    const GradLayout gl = grad_layout(e);
    if (phase >= 0) { *begin = gl.per_layer * phase; *end = gl.per_layer * (phase + 1); }
    else if (phase == -1) { *begin = gl.head_w; *end = gl.embed; }
    else { *begin = gl.embed; *end = gl.total; }
    return 0;
}

int showo_optimizer_enable(showo_engine_t* e) {
    SHOWO_CHECK(e != nullptr, "null engine");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    if (e->opt) return 0;
    OptState* o = new OptState();
    o->n = grad_layout(e).total + (int64_t)e->NL * e->D;
    if (dev_alloc(&o->master, (size_t)o->n) || dev_alloc(&o->m, (size_t)o->n) || dev_alloc(&o->v, (size_t)o->n)) { opt_state_destroy(o); return -1; }
    SHOWO_CUDA_OK(cudaMemset(o->master, 0, (size_t)o->n * 4));
    SHOWO_CUDA_OK(cudaMemset(o->m, 0, (size_t)o->n * 4));
    SHOWO_CUDA_OK(cudaMemset(o->v, 0, (size_t)o->n * 4));
    e->opt = o;
    e->loaded.clear();                             // every parameter has to be handed over again so that its fp32 value is kept
    e->finalized = false;
    return 0;
}

int showo_adamw_step(showo_engine_t* e, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    SHOWO_TRY(engine_check_ready(e));
    SHOWO_CHECK(e->opt != nullptr, "adamw_step: call showo_optimizer_enable (before loading the weights) first");
    SHOWO_CHECK(e->train && e->train->grads, "adamw_step: no gradients (run showo_train_forward + showo_backward first)");
    cudaStream_t st = (cudaStream_t)stream;
    OptState* o = e->opt;
    const GradLayout gl = grad_layout(e);
    const int64_t D = e->D, F = e->F, V = e->V, W1N = e->W1N, W2K = e->W2K;
    const int64_t t = ++o->step;
    const float bc1 = 1.f - powf(beta1, (float)t), bc2s = sqrtf(1.f - powf(beta2, (float)t));
    float* G = e->train->grads;
    // one tensor: `off` in the master / moment buffers, `goff` in the gradient buffer, `decay` per the reference's grouping
    auto upd = [&](int64_t off, int64_t goff, int64_t rows, int64_t cols, int64_t ld, bool decay, bf16* out16, float* out32, int64_t out_ld) -> int {
        const int64_t n = rows * cols;
        const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
        // the gradient sits at the same (row, col) position of its own packed block: both buffers share `ld`
        adamw_kernel<<<grid, 256, 0, st>>>(o->master + off, G + goff - 0, o->m + off, o->v + off, rows, cols, ld, lr, beta1, beta2, eps,
                                           decay ? weight_decay : 0.f, bc1, bc2s, out16, out32, out_ld);
        note_launch();
        return 0;
    };
    (void)F;
    for (int l = 0; l < e->NL; ++l) {
        const LayerW& w = e->layers[l];
        const int64_t b = gl.per_layer * l;
        // NB adamw_kernel indexes gradient and master with the same offsets relative to their bases: pass bases shifted alike
        SHOWO_TRY(upd(b + gl.w1, b + gl.w1, W1N, D, D, true, w.w1, nullptr, D));
        SHOWO_TRY(upd(b + gl.b1, b + gl.b1, 1, W1N, W1N, false, nullptr, w.b1, W1N));
        SHOWO_TRY(upd(b + gl.w2, b + gl.w2, D, W2K, W2K, true, w.w2, nullptr, W2K));
        SHOWO_TRY(upd(b + gl.b2, b + gl.b2, 1, D, D, false, nullptr, w.b_dense, D));                          // self_attn.dense.bias
        SHOWO_TRY(upd(gl.total + (int64_t)l * D, b + gl.b2, 1, D, D, false, nullptr, w.b_fc2, D));           // mlp.fc2.bias: same gradient
        SHOWO_TRY(upd(b + gl.ln_g, b + gl.ln_g, 1, D, D, true, nullptr, w.ln_g, D));
        SHOWO_TRY(upd(b + gl.ln_b, b + gl.ln_b, 1, D, D, false, nullptr, w.ln_b, D));
        SHOWO_TRY(upd(b + gl.qg, b + gl.qg, 1, 64, 64, true, nullptr, w.qg, 64));
        SHOWO_TRY(upd(b + gl.qb, b + gl.qb, 1, 64, 64, false, nullptr, w.qb, 64));
        SHOWO_TRY(upd(b + gl.kg, b + gl.kg, 1, 64, 64, true, nullptr, w.kg, 64));
        SHOWO_TRY(upd(b + gl.kb, b + gl.kb, 1, 64, 64, false, nullptr, w.kb, 64));
    }
    SHOWO_TRY(upd(gl.head_w, gl.head_w, V, D, D, true, e->head_w, nullptr, D));
    SHOWO_TRY(upd(gl.head_b, gl.head_b, 1, V, V, false, nullptr, e->head_b, V));
    SHOWO_TRY(upd(gl.fln_g, gl.fln_g, 1, D, D, true, nullptr, e->fln_g, D));
    SHOWO_TRY(upd(gl.fln_b, gl.fln_b, 1, D, D, false, nullptr, e->fln_b, D));
    SHOWO_TRY(upd(gl.embed, gl.embed, V, D, D, true, e->embed, nullptr, D));
    if (o->mmp_master && e->mmp_grads && e->mmp_grads_valid) {
        // Showo.mm_projector (train_w_clip_vit.py trains it with the same AdamW groups: weights decayed, biases not)
        auto updp = [&](int64_t off, int64_t rows, int64_t cols, bool decay, bf16* out16, float* out32) {
            const int64_t n = rows * cols;
            const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
            adamw_kernel<<<grid, 256, 0, st>>>(o->mmp_master + off, e->mmp_grads + off, o->mmp_m + off, o->mmp_v + off, rows, cols, cols, lr, beta1,
                                               beta2, eps, decay ? weight_decay : 0.f, bc1, bc2s, out16, out32, cols);
            note_launch();
        };
        updp(kMmpW0, kMmpMid, kMmpIn, true, e->mmp_w0, nullptr);
        updp(kMmpB0, 1, kMmpMid, false, nullptr, e->mmp_b0);
        updp(kMmpW2, kMmpOut, kMmpMid, true, e->mmp_w2, nullptr);
        updp(kMmpB2, 1, kMmpOut, false, nullptr, e->mmp_b2);
        e->mmp_grads_valid = false;                    // consumed: a step without a projector backward leaves the projector alone
        ++e->mmp_version;
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    return engine_refresh_derived(e, st);
}

int showo_read_param(showo_engine_t* e, const char* name, float* out_dev, int64_t numel, void* stream) {
    SHOWO_CHECK(e && e->opt && name && out_dev, "read_param: needs an engine with the optimizer enabled");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    float* p; int64_t rows, cols, ld;
    if (is_mmp_name(name)) { SHOWO_TRY(mmp_slot(e->opt->mmp_master, name, &p, &rows, &cols)); ld = cols; }
    else
    SHOWO_TRY(named_slot(e, e->opt->master, true, std::string(name), &p, &rows, &cols, &ld));
    SHOWO_CHECK(rows * cols == numel, std::string("read_param: ") + name + " has " + std::to_string(rows * cols) + " elements");
    SHOWO_CUDA_OK(cudaMemcpy2DAsync(out_dev, (size_t)cols * 4, p, (size_t)ld * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int showo_read_grad(showo_engine_t* e, const char* name, float* out_dev, int64_t numel, void* stream) {
    SHOWO_CHECK(e && name && out_dev, "read_grad: null argument");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    float* p = nullptr; int64_t rows = 0, cols = 0, ld = 0;
    if (is_mmp_name(name)) {
        SHOWO_CHECK(e->mmp_grads && e->mmp_grads_valid, "read_grad: no showo_mm_projector_backward has run since the last optimizer step");
        SHOWO_TRY(mmp_slot(e->mmp_grads, name, &p, &rows, &cols));
        ld = cols;
    } else {
        SHOWO_CHECK(e->train && e->train->grads, "read_grad: no backward has run");
        SHOWO_TRY(named_grad(e, name, &p, &rows, &cols, &ld));
    }
    SHOWO_CHECK(numel == rows * cols, std::string("read_grad: ") + name + " has " + std::to_string(rows * cols) + " elements, got " + std::to_string(numel));
    return copy_f32_to_f32_rows(p, ld, out_dev, cols, (int)rows, (int)cols, (cudaStream_t)stream);
}

// Backward of Showo.mm_projector for the rows of the LAST showo_mm_projector call (train_w_clip_vit.py:599-601 differentiates it through
// `input_embeddings`): dY [n, 2048] fp32 = the slice of showo_backward's dembeds that the projector's output occupied.
//   dW2 = dY^T gelu(H),  db2 = colsum dY,  dH = (dY W2) o gelu'(H),  dW0 = dH^T X,  db0 = colsum dH      (H = X W0^T + b0)
// The CLIP features X are not differentiated (the tower is frozen, train_w_clip_vit.py:199-203).  bf16 operands, fp32 accumulation,
// fp32 gradients in e->mmp_grads = [w0 | b0 | w2 | b2].
int showo_mm_projector_backward(showo_engine_t* e, const float* dy_dev, int64_t n, void* stream) {
    SHOWO_CHECK(e && dy_dev && n > 0, "mm_projector_backward: bad arguments");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(e->mmp_pre && e->mmp_n == n, "mm_projector_backward: differentiates the last showo_mm_projector call, which had " +
                                                 std::to_string(e->mmp_n) + " rows (got " + std::to_string(n) + ")");
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->train) e->train = new TrainState();
    TrainState* t = e->train;
    if (!e->mmp_grads) SHOWO_TRY(dev_alloc(&e->mmp_grads, (size_t)kMmpTotal));
    if (!e->mmp_w2t) SHOWO_TRY(dev_alloc(&e->mmp_w2t, (size_t)(kMmpMid * kMmpOut)));
    if (n > e->mmp_bwd_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        dev_free(e->mmp_dy); dev_free(e->mmp_dmid);
        SHOWO_TRY(dev_alloc(&e->mmp_dy, (size_t)(n * kMmpOut)));
        SHOWO_TRY(dev_alloc(&e->mmp_dmid, (size_t)(n * kMmpMid)));
        e->mmp_bwd_cap = n;
    }
    if (e->mmp_w2t_version != e->mmp_version) {       // w2t[j][o] = w2[o][j]: the dgrad GEMM's K-contiguous B operand
        SHOWO_TRY(transpose_to_bf16<bf16>(e->mmp_w2, kMmpMid, (int)kMmpOut, (int)kMmpMid, e->mmp_w2t, kMmpOut, nullptr, 0, st));
        e->mmp_w2t_version = e->mmp_version;
    }
    const int64_t l0 = launches_total();
    const int M = (int)n;
    const int64_t Mp = (int64_t)(M + 127) / 128 * 128;
    float* G = e->mmp_grads;
    SHOWO_TRY(f32_to_bf16(dy_dev, e->mmp_dy, n * kMmpOut, st));
    // (always the token-major GEMM: the SHOWO_WGRAD_MN=0 scratch belongs to the backbone's training state and is sized by its rows)
    SHOWO_TRY(wgrad(t, e->mmp_dy, kMmpOut, (int)kMmpOut, e->mmp_mid, kMmpMid, (int)kMmpMid, M, Mp, G + kMmpW2, kMmpMid, G + kMmpB2, st, true));
    SHOWO_TRY(gemm_plain(e->mmp_dy, kMmpOut, e->mmp_w2t, kMmpOut, M, (int)kMmpMid, (int)kMmpOut, e->mmp_dmid, kMmpMid, false, st));
    SHOWO_TRY(gelu_erf_bwd_bf16(e->mmp_dmid, e->mmp_pre, n * kMmpMid, st));
    SHOWO_TRY(wgrad(t, e->mmp_dmid, kMmpMid, (int)kMmpMid, e->mmp_in, kMmpIn, (int)kMmpIn, M, Mp, G + kMmpW0, kMmpIn, G + kMmpB0, st, true));
    e->mmp_grads_valid = true;
    e->launches_last = launches_total() - l0;
    return 0;
}

// the projector's gradient buffer ([w0 | b0 | w2 | b2] fp32): the extra bucket of the data-parallel gradient all-reduce
int showo_mm_projector_grad_buffer(showo_engine_t* e, float** base_dev, int64_t* numel) {
    SHOWO_CHECK(e && base_dev && numel, "mm_projector_grad_buffer: null argument");
    SHOWO_CHECK(e->mmp_grads != nullptr, "mm_projector_grad_buffer: no showo_mm_projector_backward has run");
    *base_dev = e->mmp_grads;
    *numel = kMmpTotal;
    return 0;
}

int showo_attention_bwd_test(const void* q_dev, const void* k_dev, const void* v_dev, const void* o_dev, const void* do_dev,
                             const float* lse_dev, void* dq_dev, void* dk_dev, void* dv_dev, int n_seq, int L, int H,
                             const showo_seq_mask_t* masks_host, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    showo_seq_mask_t* dm = nullptr; float* delta = nullptr;
    SHOWO_TRY(dev_alloc(&dm, (size_t)n_seq));
    SHOWO_TRY(dev_alloc(&delta, (size_t)n_seq * L * H));
    SHOWO_CUDA_OK(cudaMemcpyAsync(dm, masks_host, (size_t)n_seq * sizeof(showo_seq_mask_t), cudaMemcpyHostToDevice, st));
    const int64_t ld = (int64_t)H * 64;
    AttnBwdArgs a{};
    a.q = (const bf16*)q_dev; a.k = (const bf16*)k_dev; a.v = (const bf16*)v_dev; a.o = (const bf16*)o_dev; a.d_o = (const bf16*)do_dev;
    a.q_ld = a.k_ld = a.v_ld = a.o_ld = a.do_ld = a.dq_ld = a.dk_ld = a.dv_ld = ld;
    a.lse = lse_dev; a.delta = delta; a.dq = (bf16*)dq_dev; a.dk = (bf16*)dk_dev; a.dv = (bf16*)dv_dev;
    a.n_seq = n_seq; a.H = H; a.L = L; a.masks = dm; a.scale = 0.125f;
    const int rc = omni_attention_backward(a, st);
    cudaStreamSynchronize(st);
    cudaFree(dm); cudaFree(delta);
    return rc;
}

}  // extern "C"
