// MAGVIT-v2 VQ tokenizer (models/modeling_magvitv2.py, models/common_modules.py) on sm_100a.
//
// Activations are NHWC bf16.  Every 3x3 / 1x1 convolution with >= 64 input channels is an implicit GEMM on the tcgen05
// kernel (gemm_tcgen05.cuh, A_CONV3 mode: TMA fetches the shifted pixel patch per filter tap, zero padding = TMA
// out-of-bounds fill) with bias and the ResnetBlock skip connection fused into the epilogue.  GroupNorm(32, eps 1e-6)
// + swish, nearest x2 upsampling, the LFQ lookup / sign-pack and the 3-channel image-side convolutions are coalesced
// HBM-bound kernels (16 B per thread per access).
#include <map>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace showo {

// ------------------------------------------------------------------------------------------------ weight packing
// src [cout, cin, k, k] fp32 (torch Conv2d) -> dst [cout_pad, k*k*cin_pad] bf16, K index = (ky*k + kx) * cin_pad + c
__global__ void pack_conv_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin, int k,
                                 int cin_pad, int row0, int ld) {
    const int taps = k * k;
    const int64_t n = (int64_t)cout * taps * cin_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cin_pad);
        const int tap = (int)((i / cin_pad) % taps);
        const int o = (int)(i / ((int64_t)cin_pad * taps));
        const float v = (c < cin) ? src[((int64_t)o * cin + c) * taps + tap] : 0.f;
        dst[(int64_t)(row0 + o) * ld + (int64_t)tap * cin_pad + c] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm (+swish)
// stats[n][g] = (sum, sumsq) over H*W*(C/32) elements.  One thread owns 8 consecutive channels of a pixel.
// Deterministic two-stage reduction (no atomics): each CTA reduces its pixel slab in a fixed order into
// partials[n][block][g], gn_finalize_kernel sums the blocks in order -> bit-reproducible images run to run.
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ partials, int HW,
                                                       int C, int pix_per_block) {
    __shared__ float4 part[256];
    const int n = blockIdx.y;
    const int oct = C >> 3;                    // octets per pixel
    const int o = threadIdx.x % oct;
    const int pl = threadIdx.x / oct;
    const int pstride = 256 / oct;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(p0 + pix_per_block, HW);
    float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
    const bf16* base = x + (int64_t)n * HW * C + o * 8;
    for (int p = p0 + pl; p < p1; p += pstride) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + (int64_t)p * C);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]),
                     f3 = __bfloat1622float2(h[3]);
        a0 += (f0.x + f0.y) + (f1.x + f1.y);
        q0 += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
        a1 += (f2.x + f2.y) + (f3.x + f3.y);
        q1 += (f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y);
    }
    part[threadIdx.x] = make_float4(a0, q0, a1, q1);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x;
        const int cpg = C >> 5;                // 4, 8 or 16 channels per group
        float s = 0.f, q = 0.f;
        // half-octets [h0, h1) of this group; half-octet index = channel / 4
        const int h0 = g * cpg / 4, h1 = (g + 1) * cpg / 4;
        for (int hh = h0; hh < h1; ++hh) {
            const int oo = hh >> 1, half = hh & 1;
            for (int r = 0; r < pstride; ++r) {
                const float4 v = part[r * oct + oo];
                s += half ? v.z : v.x;
                q += half ? v.w : v.y;
            }
        }
        float* dst = partials + (((int64_t)n * gridDim.x + blockIdx.x) * 32 + g) * 2;
        dst[0] = s; dst[1] = q;
    }
}
__global__ void gn_finalize_kernel(const float* __restrict__ partials, float* __restrict__ stats, int nblk) {
    const int n = blockIdx.x, g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int b = 0; b < nblk; ++b) {
        const float* src = partials + (((int64_t)n * nblk + b) * 32 + g) * 2;
        s += src[0]; q += src[1];
    }
    stats[((int64_t)n * 32 + g) * 2] = s;
    stats[((int64_t)n * 32 + g) * 2 + 1] = q;
}
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       bf16* __restrict__ y, int NB, int HW, int C, float eps, int swish) {
    const int oct = C >> 3;
    const int cpg = C >> 5;
    const int64_t total = (int64_t)NB * HW * oct;
    const float inv_n = 1.f / ((float)HW * (float)cpg);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % oct);
        const int n = (int)(i / ((int64_t)HW * oct));
        const uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
        float out[8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int g = (o * 8 + half * 4) / cpg;
            const float s = stats[((int64_t)n * 32 + g) * 2], q = stats[((int64_t)n * 32 + g) * 2 + 1];
            const float mean = s * inv_n;
            const float var = fmaxf(q * inv_n - mean * mean, 0.f);
            const float rstd = rsqrtf(var + eps);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = o * 8 + half * 4 + j;
                float t = (v[half * 4 + j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
                if (swish) t = t / (1.f + __expf(-t));
                out[half * 4 + j] = t;
            }
        }
        uint4 pk;
        pk.x = pack_bf16(out[0], out[1]); pk.y = pack_bf16(out[2], out[3]);
        pk.z = pack_bf16(out[4], out[5]); pk.w = pack_bf16(out[6], out[7]);
        *reinterpret_cast<uint4*>(y + i * 8) = pk;
    }
}

// ------------------------------------------------------------------------------------------------ nearest x2 upsample
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int NB, int H,
                                                         int W, int C) {
    const int oct = C >> 3;
    const int64_t total = (int64_t)NB * (2 * H) * (2 * W) * oct;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % oct);
        int64_t p = i / oct;
        const int xo = (int)(p % (2 * W)); p /= (2 * W);
        const int yo = (int)(p % (2 * H));
        const int n = (int)(p / (2 * H));
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + (yo >> 1)) * W + (xo >> 1)) * C + o * 8);
        *reinterpret_cast<uint4*>(y + i * 8) = u;
    }
}

// ------------------------------------------------------------------------------------------------ LFQ lookup + post_quant_conv
// ids [B, hw] -> z [B, hw, 64] bf16 (13 live channels): bit c of the code (channel 0 = MSB) -> +-1, then the 1x1
// post_quant_conv (13x13) in fp32.            modeling_magvitv2.py:186-221,371
__global__ void lfq_decode_kernel(const int64_t* __restrict__ ids, const float* __restrict__ w /*[13][13]*/,
                                  const float* __restrict__ b, bf16* __restrict__ z, int64_t n_pix) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const int64_t id = ids[i];
    float e[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) e[c] = ((id >> (12 - c)) & 1) ? 1.f : -1.f;
    bf16* o = z + i * 64;
#pragma unroll
    for (int co = 0; co < 13; ++co) {
        float acc = b[co];
#pragma unroll
        for (int c = 0; c < 13; ++c) acc += w[co * 13 + c] * e[c];
        o[co] = __float2bfloat16(acc);
    }
    for (int c = 13; c < 64; ++c) o[c] = __float2bfloat16(0.f);
}

// encoder tail: quant_conv (1x1, 13x13, fp32) on the conv_out result [n_pix, ld] bf16, then sign -> 13-bit code
__global__ void lfq_encode_kernel(const bf16* __restrict__ h, int ld, const float* __restrict__ w, const float* __restrict__ b,
                                  int64_t* __restrict__ ids, int64_t n_pix) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float e[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) e[c] = __bfloat162float(h[i * ld + c]);
    int64_t code = 0;
#pragma unroll
    for (int co = 0; co < 13; ++co) {
        float acc = b[co];
#pragma unroll
        for (int c = 0; c < 13; ++c) acc += w[co * 13 + c] * e[c];
        code |= (int64_t)(acc > 0.f ? 1 : 0) << (12 - co);
    }
    ids[i] = code;
}

// ------------------------------------------------------------------------------------------------ conv_out (C -> 3), fp32 out
// 3x3, pad 1, NHWC bf16 in, weights fp32 [3][C][3][3] staged in smem; writes NCHW fp32 and/or uint8 NHWC image
// (inference_t2i.py:338-341: clamp((x+1)/2, 0, 1) * 255 -> uint8).
__global__ void __launch_bounds__(256) conv_out3_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ out_nchw,
                                                        uint8_t* __restrict__ out_u8, int NB, int H, int W, int C) {
    extern __shared__ float ws[];      // [9][C][3]
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
        const int co = i % 3, c = (i / 3) % C, tap = i / (3 * C);
        ws[i] = w[((int64_t)co * C + c) * 9 + tap];
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)NB * H * W) return;
    const int xo = (int)(p % W), yo = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    float a0 = b[0], a1 = b[1], a2 = b[2];
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = yo + tap / 3 - 1, xx = xo + tap % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const bf16* px = x + (((int64_t)n * H + yy) * W + xx) * C;
        const float* wt = ws + tap * C * 3;
        for (int c8 = 0; c8 < C; c8 += 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(px + c8);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h[j]);
                const float* w0 = wt + (c8 + 2 * j) * 3;
                a0 += f.x * w0[0] + f.y * w0[3];
                a1 += f.x * w0[1] + f.y * w0[4];
                a2 += f.x * w0[2] + f.y * w0[5];
            }
        }
    }
    if (out_nchw) {
        const int64_t hw = (int64_t)H * W, base = (int64_t)n * 3 * hw + (int64_t)yo * W + xo;
        out_nchw[base] = a0; out_nchw[base + hw] = a1; out_nchw[base + 2 * hw] = a2;
    }
    if (out_u8) {
        const float v[3] = {a0, a1, a2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // same fp32 op sequence as torch.clamp((x + 1) / 2, 0, 1) * 255 (no FMA contraction)
            float t = __fmul_rn(fminf(fmaxf(__fmul_rn(__fadd_rn(v[c], 1.0f), 0.5f), 0.f), 1.f), 255.0f);
            out_u8[p * 3 + c] = (uint8_t)t;          // .astype(uint8) truncates
        }
    }
}

// encoder head: conv_in 3 -> C (3x3, pad 1) from NCHW fp32 pixels to NHWC bf16; one thread = one pixel x 8 out channels
__global__ void __launch_bounds__(256) conv_in3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, bf16* __restrict__ y, int NB, int H,
                                                       int W, int C) {
    extern __shared__ float ws[];      // [27][C]
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
        const int co = i % C, k = i / C;          // k = ci*9 + tap
        ws[i] = w[(int64_t)co * 27 + k];
    }
    __syncthreads();
    const int oct = C >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)NB * H * W * oct) return;
    const int o = (int)(i % oct);
    const int64_t p = i / oct;
    const int xo = (int)(p % W), yo = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = b[o * 8 + j];
    for (int ci = 0; ci < 3; ++ci)
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = yo + tap / 3 - 1, xx = xo + tap % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float v = x[(((int64_t)n * 3 + ci) * H + yy) * W + xx];
            const float* wk = ws + (ci * 9 + tap) * C + o * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v * wk[j];
        }
    uint4 pk;
    pk.x = pack_bf16(acc[0], acc[1]); pk.y = pack_bf16(acc[2], acc[3]);
    pk.z = pack_bf16(acc[4], acc[5]); pk.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(y + p * C + o * 8) = pk;
}

// space-to-depth for the stride-2 Downsample (common_modules.py:83-87: pad (0,1,0,1) then 3x3 stride 2):
// y[n, yo, xo, (py*2+px)*C + c] = x[n, 2yo+py, 2xo+px, c]
__global__ void __launch_bounds__(256) space_to_depth_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int NB,
                                                             int H, int W, int C) {
    const int oct = C >> 3;
    const int Ho = H >> 1, Wo = W >> 1;
    const int64_t total = (int64_t)NB * Ho * Wo * 4 * oct;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % oct);
        int64_t r = i / oct;
        const int par = (int)(r % 4); r /= 4;
        const int xo = (int)(r % Wo); r /= Wo;
        const int yo = (int)(r % Ho);
        const int n = (int)(r / Ho);
        const uint4 u = *reinterpret_cast<const uint4*>(
            x + (((int64_t)n * H + 2 * yo + (par >> 1)) * W + 2 * xo + (par & 1)) * C + o * 8);
        *reinterpret_cast<uint4*>(y + i * 8) = u;
    }
}
// weights of the stride-2 conv rearranged for the space-to-depth input: a 2x2 "conv" with taps (sy,sx) over 4C channels.
// dst [cout][ (sy*2+sx) * 4C + (py*2+px)*C + c ] = w[cout][c][dy][dx] with dy = 2*sy+py, dx = 2*sx+px (zero if dy or dx > 2)
__global__ void pack_down_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int C) {
    const int64_t n = (int64_t)cout * 16 * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int par = (int)((i / C) % 4);
        const int s = (int)((i / (4 * (int64_t)C)) % 4);
        const int o = (int)(i / (16 * (int64_t)C));
        const int dy = 2 * (s >> 1) + (par >> 1), dx = 2 * (s & 1) + (par & 1);
        const float v = (dy < 3 && dx < 3) ? src[((int64_t)o * C + c) * 9 + dy * 3 + dx] : 0.f;
        dst[i] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------------------------------------ mid-block attention helpers
// row softmax with scale: s fp32 [rows, n] -> p bf16 [rows, n]           (common_modules.py:196-198)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int n,
                                                           float scale) {
    __shared__ float red[8];
    const float* r = s + (int64_t)blockIdx.x * n;
    float mx = -3.0e38f;
    for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, r[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) sum += __expf((r[i] - mx) * scale);
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int i = 0; i < 8; ++i) sum += red[i];
    const float inv = 1.f / sum;
    for (int i = threadIdx.x; i < n; i += 256)
        p[(int64_t)blockIdx.x * n + i] = __float2bfloat16(__expf((r[i] - mx) * scale) * inv);
}
// v [n_tok, ld] (C columns starting at v) -> vt [C, n_tok]
__global__ void transpose_kernel(const bf16* __restrict__ v, int64_t ld, bf16* __restrict__ vt, int n_tok, int C) {
    __shared__ bf16 tile[32][34];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8)
        if (t0 + j < n_tok) tile[j][threadIdx.x] = v[(int64_t)(t0 + j) * ld + c0 + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8)
        if (t0 + threadIdx.x < n_tok) vt[(int64_t)(c0 + j) * n_tok + t0 + threadIdx.x] = tile[threadIdx.x][j];
}


// ================================================================================================ fp32 verification path
// SURVEY section 7 ("an fp32-accumulate verification mode is needed for any stricter claim"): a second, deliberately plain
// implementation of get_code / decode_code -- fp32 NCHW activations, fp32 weights in torch's own layout, CUDA cores only, no tensor
// cores, no bf16 anywhere -- that follows models/modeling_magvitv2.py / common_modules.py operation by operation.  It exists to
// separate "bf16 rounding of the fast path" from "a defect": against the CPU oracle it agrees to fp32 re-association level, and the
// fast path's LFQ sign flips are shown to sit only where THIS path's pre-sign value is ~0 (tests/test_gpu_parity.py).  Not a
// fallback: nothing on the product path calls it.
__global__ void __launch_bounds__(256) vconv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                    const float* __restrict__ resid, float* __restrict__ y, int NB, int cin, int H, int W,
                                                    int cout, int Ho, int Wo, int k, int stride, int pad_lo) {
    const int64_t total = (int64_t)NB * cout * Ho * Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho), co = (int)((i / ((int64_t)Wo * Ho)) % cout), n = (int)(i / ((int64_t)Wo * Ho * cout));
        float acc = b ? b[co] : 0.f;
        for (int ci = 0; ci < cin; ++ci) {
            const float* xp = x + ((int64_t)n * cin + ci) * H * W;
            const float* wp = w + ((int64_t)co * cin + ci) * k * k;
            for (int ky = 0; ky < k; ++ky) {
                const int yi = yo * stride + ky - pad_lo;
                if (yi < 0 || yi >= H) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int xi = xo * stride + kx - pad_lo;
                    if (xi < 0 || xi >= W) continue;
                    acc = fmaf(wp[ky * k + kx], xp[(int64_t)yi * W + xi], acc);
                }
            }
        }
        if (resid) acc += resid[i];
        y[i] = acc;
    }
}
// GroupNorm(32 groups, eps) (+ swish) on NCHW fp32: one CTA per (image, group), two-pass statistics
__global__ void __launch_bounds__(256) vgn_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  float* __restrict__ y, int C, int HW, float eps, int swish) {
    __shared__ float red[8];
    __shared__ float stat[2];
    const int cpg = C / 32, g = blockIdx.x % 32, n = blockIdx.x / 32;
    const int64_t base = ((int64_t)n * C + (int64_t)g * cpg) * HW, cnt = (int64_t)cpg * HW;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < cnt; i += 256) s += x[base + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; stat[0] = t / (float)cnt; }
    __syncthreads();
    const float mean = stat[0];
    float q = 0.f;
    for (int64_t i = threadIdx.x; i < cnt; i += 256) { const float d = x[base + i] - mean; q += d * d; }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; stat[1] = rsqrtf(t / (float)cnt + eps); }
    __syncthreads();
    const float rstd = stat[1];
    for (int64_t i = threadIdx.x; i < cnt; i += 256) {
        const int c = g * cpg + (int)(i / HW);
        float t = (x[base + i] - mean) * rstd * gamma[c] + beta[c];
        if (swish) t = t / (1.f + expf(-t));
        y[base + i] = t;
    }
}
__global__ void vupsample_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int H, int W) {
    const int64_t total = planes * 4 * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % (2 * W)), yo = (int)((i / (2 * W)) % (2 * H));
        const int64_t pl = i / ((int64_t)4 * H * W);
        y[i] = x[(pl * H + (yo >> 1)) * W + (xo >> 1)];
    }
}
// AttnBlock (common_modules.py:168-211) on fused q|k|v [n, 3C, HW]: S[i][j] = sum_c q[c][i] k[c][j] * C^-0.5, softmax over j, out[c][i] = sum_j v[c][j] P[i][j]
__global__ void vattn_scores_kernel(const float* __restrict__ qkv, float* __restrict__ S, int C, int HW, float scale) {
    const int64_t total = (int64_t)HW * HW;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % HW), i = (int)(t / HW);
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(qkv[(int64_t)c * HW + i], qkv[(int64_t)(C + c) * HW + j], acc);
        S[t] = acc * scale;
    }
}
__global__ void __launch_bounds__(256) vsoftmax_kernel(float* __restrict__ S, int n) {
    __shared__ float red[8];
    float* r = S + (int64_t)blockIdx.x * n;
    float mx = -3.0e38f;
    for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, r[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) { const float e = expf(r[i] - mx); r[i] = e; sum += e; }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int i = 0; i < 8; ++i) sum += red[i];
    const float inv = 1.f / sum;
    for (int i = threadIdx.x; i < n; i += 256) r[i] *= inv;
}
__global__ void vattn_out_kernel(const float* __restrict__ qkv, const float* __restrict__ P, float* __restrict__ out, int C, int HW) {
    const int64_t total = (int64_t)C * HW;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % HW), c = (int)(t / HW);
        const float* v = qkv + (int64_t)(2 * C + c) * HW;
        const float* p = P + (int64_t)i * HW;
        float acc = 0.f;
        for (int j = 0; j < HW; ++j) acc = fmaf(v[j], p[j], acc);
        out[t] = acc;
    }
}
// LFQ (modeling_magvitv2.py:186-221): code -> +-1 entries [n, 13, hw]; pre-sign values [n, 13, hw] -> code (channel 0 = MSB)
__global__ void vlfq_entry_kernel(const int64_t* __restrict__ ids, float* __restrict__ z, int64_t n_img, int hw) {
    const int64_t total = n_img * 13 * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % hw), c = (int)((i / hw) % 13);
        const int64_t n = i / ((int64_t)13 * hw);
        z[i] = ((ids[n * hw + p] >> (12 - c)) & 1) ? 1.f : -1.f;
    }
}
__global__ void vlfq_index_kernel(const float* __restrict__ z, int64_t* __restrict__ ids, int64_t n_img, int hw) {
    const int64_t total = n_img * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % hw);
        const int64_t n = i / hw;
        int64_t code = 0;
        for (int c = 0; c < 13; ++c) code |= (int64_t)(z[(n * 13 + c) * hw + p] > 0.f ? 1 : 0) << (12 - c);
        ids[i] = code;
    }
}

}  // namespace showo

using namespace showo;

// ================================================================================================== engine
struct ConvW {
    bf16* w = nullptr; float* b = nullptr; float* w32 = nullptr;   // w32: fp32 copy for the SIMT image-side convs / 13x13
    int cin = 0, cin_pad = 0, cout = 0, cout_pad = 0, k = 0, ld = 0;
    bool down = false;
};
struct NormW { float* g = nullptr; float* b = nullptr; int c = 0; };

struct magvit_engine {
    int device = 0;
    std::map<std::string, ConvW> conv;
    std::map<std::string, NormW> norm;
    std::set<std::string> loaded;
    float* stage = nullptr; size_t stage_cap = 0;
    bf16* bufs[4] = {nullptr, nullptr, nullptr, nullptr}; size_t buf_cap = 0;
    float* stats = nullptr; float* partials = nullptr;
    float* scores = nullptr; size_t scores_cap = 0; bf16* probs = nullptr; bf16* vt = nullptr;
    float* vbufs[4] = {nullptr, nullptr, nullptr, nullptr}; size_t vbuf_cap = 0;     // fp32 verification path (NCHW)
    float* vscores = nullptr; size_t vscores_cap = 0;
    int64_t launches_last = 0;
};

namespace showo { int64_t launches_total(); }

static const int kEncMult[5] = {1, 2, 2, 4, 4}, kEncBlocks[5] = {4, 3, 4, 3, 4};
static const int kDecMult[5] = {1, 1, 2, 2, 4}, kDecBlocks[5] = {4, 4, 3, 4, 3};
static const int kCh = 128, kZ = 13;

static void reg_conv(magvit_engine* m, const std::string& name, int cout, int cin, int k, bool down = false) {
    ConvW c; c.cin = cin; c.cout = cout; c.k = k; c.down = down;
    c.cin_pad = cdiv(cin, 64) * 64; c.cout_pad = cdiv(cout, 64) * 64;
    c.ld = down ? 16 * cin : k * k * c.cin_pad;
    m->conv[name] = c;
}
static void reg_norm(magvit_engine* m, const std::string& name, int c) { NormW n; n.c = c; m->norm[name] = n; }
static void reg_res(magvit_engine* m, const std::string& p, int cin, int cout) {
    reg_norm(m, p + ".norm1", cin); reg_conv(m, p + ".conv1", cout, cin, 3);
    reg_norm(m, p + ".norm2", cout); reg_conv(m, p + ".conv2", cout, cout, 3);
    if (cin != cout) reg_conv(m, p + ".nin_shortcut", cout, cin, 1);
}
static void reg_attn(magvit_engine* m, const std::string& p, int c) {
    reg_norm(m, p + ".norm", c);
    reg_conv(m, p + ".qkv", 3 * c, c, 1);      // fused q | k | v rows
    reg_conv(m, p + ".proj_out", c, c, 1);
}

static int magvit_alloc_all(magvit_engine* m) {
    for (auto& kv : m->conv) {
        ConvW& c = kv.second;
        SHOWO_CUDA_OK(cudaMalloc(&c.w, (size_t)c.cout_pad * c.ld * sizeof(bf16)));
        SHOWO_CUDA_OK(cudaMemset(c.w, 0, (size_t)c.cout_pad * c.ld * sizeof(bf16)));
        SHOWO_CUDA_OK(cudaMalloc(&c.b, (size_t)c.cout * 4));
        SHOWO_CUDA_OK(cudaMalloc(&c.w32, (size_t)c.cout * c.cin * c.k * c.k * 4));
    }
    for (auto& kv : m->norm) {
        SHOWO_CUDA_OK(cudaMalloc(&kv.second.g, (size_t)kv.second.c * 4));
        SHOWO_CUDA_OK(cudaMalloc(&kv.second.b, (size_t)kv.second.c * 4));
    }
    return 0;
}

static int ensure_bufs(magvit_engine* m, size_t elems, cudaStream_t st) {
    if (elems <= m->buf_cap) return 0;
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; ++i) {
        if (m->bufs[i]) cudaFree(m->bufs[i]);
        SHOWO_CUDA_OK(cudaMalloc(&m->bufs[i], elems * sizeof(bf16)));
    }
    m->buf_cap = elems;
    return 0;
}

// ---- layer helpers (all NHWC bf16)
static int gn(magvit_engine* m, const std::string& name, const bf16* x, bf16* y, int NB, int H, int W, int C, bool swish,
              cudaStream_t st) {
    auto it = m->norm.find(name);
    SHOWO_CHECK(it != m->norm.end() && it->second.c == C, "groupnorm " + name + ": unknown or channel mismatch");
    const int HW = H * W;
    SHOWO_CHECK(NB <= 32, "groupnorm: batch too large for the partials buffer (max 32 images per call)");
    const int nblk = cdiv(HW, 64) < 128 ? cdiv(HW, 64) : 128;
    const int ppb = cdiv(HW, nblk);
    dim3 grid(cdiv(HW, ppb), NB);
    gn_stats_kernel<<<grid, 256, 0, st>>>(x, m->partials, HW, C, ppb);
    note_launch();
    gn_finalize_kernel<<<NB, 32, 0, st>>>(m->partials, m->stats, (int)grid.x);
    note_launch();
    const int64_t total = (int64_t)NB * HW * (C / 8);
    const int g2 = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    gn_apply_kernel<<<g2, 256, 0, st>>>(x, m->stats, it->second.g, it->second.b, y, NB, HW, C, 1e-6f, swish ? 1 : 0);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int conv(magvit_engine* m, const std::string& name, const bf16* x, bf16* y, const bf16* resid, int NB, int H,
                int W, cudaStream_t st) {
    auto it = m->conv.find(name);
    SHOWO_CHECK(it != m->conv.end(), "conv " + name + ": unknown");
    const ConvW& c = it->second;
    ConvArgs a{};
    a.x = x; a.w = c.w; a.bias = c.b; a.resid = resid; a.ldr = c.cout; a.out = y; a.ldc = c.cout;
    a.NB = NB; a.H = H; a.W = W; a.cin = c.cin_pad; a.cout = c.cout; a.taps = c.k * c.k;
    return conv_nhwc_bf16(a, st);
}
// X -> out (returns index of the buffer that holds the result); scratch: the other three buffers
static int resblock(magvit_engine* m, const std::string& p, int& X, int NB, int H, int W, int cin, int cout,
                    cudaStream_t st) {
    int others[3], k = 0;
    for (int i = 0; i < 4; ++i) if (i != X) others[k++] = i;
    const int T = others[0], U = others[1], V = others[2];
    SHOWO_TRY(gn(m, p + ".norm1", m->bufs[X], m->bufs[T], NB, H, W, cin, true, st));
    SHOWO_TRY(conv(m, p + ".conv1", m->bufs[T], m->bufs[U], nullptr, NB, H, W, st));
    SHOWO_TRY(gn(m, p + ".norm2", m->bufs[U], m->bufs[T], NB, H, W, cout, true, st));
    const bf16* skip = m->bufs[X];
    if (cin != cout) {
        SHOWO_TRY(conv(m, p + ".nin_shortcut", m->bufs[X], m->bufs[V], nullptr, NB, H, W, st));
        skip = m->bufs[V];
    }
    SHOWO_TRY(conv(m, p + ".conv2", m->bufs[T], m->bufs[U], skip, NB, H, W, st));
    X = U;
    return 0;
}
static int attnblock(magvit_engine* m, const std::string& p, int& X, int NB, int H, int W, int C, cudaStream_t st) {
    int others[3], k = 0;
    for (int i = 0; i < 4; ++i) if (i != X) others[k++] = i;
    const int T = others[0], U = others[1], V = others[2];
    const int HW = H * W;
    SHOWO_TRY(gn(m, p + ".norm", m->bufs[X], m->bufs[T], NB, H, W, C, false, st));
    SHOWO_TRY(conv(m, p + ".qkv", m->bufs[T], m->bufs[U], nullptr, NB, H, W, st));      // [NB*HW, 3C]
    const size_t need = (size_t)HW * HW;
    if (need > m->scores_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (m->scores) cudaFree(m->scores); if (m->probs) cudaFree(m->probs); if (m->vt) cudaFree(m->vt);
        SHOWO_CUDA_OK(cudaMalloc(&m->scores, need * 4));
        SHOWO_CUDA_OK(cudaMalloc(&m->probs, need * 2));
        SHOWO_CUDA_OK(cudaMalloc(&m->vt, (size_t)C * HW * 2));
        m->scores_cap = need;
    }
    const float scale = 1.0f / sqrtf((float)C);
    for (int n = 0; n < NB; ++n) {
        const bf16* qkv = m->bufs[U] + (size_t)n * HW * 3 * C;
        GemmArgs g{};
        g.A = qkv; g.lda = 3 * C; g.B = qkv + C; g.ldb = 3 * C; g.M = HW; g.N = HW; g.K = C;
        g.out = m->scores; g.ldc = HW; g.bias = nullptr;
        SHOWO_TRY(gemm_bf16(g, GEMM_BIAS_F32, st));
        softmax_rows_kernel<<<HW, 256, 0, st>>>(m->scores, m->probs, HW, scale);
        note_launch();
        transpose_kernel<<<dim3(cdiv(HW, 32), C / 32), dim3(32, 8), 0, st>>>(qkv + 2 * C, 3 * C, m->vt, HW, C);
        note_launch();
        GemmArgs g2{};
        g2.A = m->probs; g2.lda = HW; g2.B = m->vt; g2.ldb = HW; g2.M = HW; g2.N = C; g2.K = HW;
        g2.out = m->bufs[T] + (size_t)n * HW * C; g2.ldc = C; g2.bias = nullptr; g2.gelu_from = C;
        SHOWO_TRY(gemm_bf16(g2, GEMM_BIAS_BF16, st));
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    SHOWO_TRY(conv(m, p + ".proj_out", m->bufs[T], m->bufs[V], m->bufs[X], NB, H, W, st));
    X = V;
    return 0;
}

static int decode_impl(magvit_engine* m, const int64_t* ids, int B, int h, int w, float* out_nchw, uint8_t* out_u8,
                       cudaStream_t st) {
    SHOWO_CHECK(m && ids && B > 0 && h > 0 && w > 0, "magvit decode: bad arguments");
    // any grid the extrapolation mode produces (inference_t2i.py:274-276, e.g. 16 x 25 with offset 1): the implicit-GEMM convolutions
    // take partial 8 x 16 pixel tiles (TMA zero fill = the padding, stores masked); the mid-block attention reads its [hw, hw]
    // probabilities as a bf16 GEMM operand, whose row stride has to be a multiple of 16 bytes
    SHOWO_CHECK((w >= 16 || (w >= 1 && (w & (w - 1)) == 0)) && ((int64_t)h * w) % 8 == 0,
                "magvit decode: grid width must be >= 16 (or a power of two) and h * w a multiple of 8");
    SHOWO_CUDA_OK(cudaSetDevice(m->device));
    SHOWO_TRY(magvit_weights_complete(m));
    const int64_t l0 = launches_total();
    const size_t max_elems = (size_t)B * (16 * h) * (16 * w) * kCh;
    SHOWO_TRY(ensure_bufs(m, max_elems, st));
    int H = h, W = w;
    const int64_t npix = (int64_t)B * h * w;
    const ConvW& pq = m->conv["decoder.post_quant_conv"];
    lfq_decode_kernel<<<(int)cdiv64(npix, 128), 128, 0, st>>>(ids, pq.w32, pq.b, m->bufs[0], npix);
    note_launch();
    int X = 1;
    SHOWO_TRY(conv(m, "decoder.conv_in", m->bufs[0], m->bufs[1], nullptr, B, H, W, st));
    int C = kCh * kDecMult[4];
    SHOWO_TRY(resblock(m, "decoder.mid.block_1", X, B, H, W, C, C, st));
    SHOWO_TRY(attnblock(m, "decoder.mid.attn_1", X, B, H, W, C, st));
    SHOWO_TRY(resblock(m, "decoder.mid.block_2", X, B, H, W, C, C, st));
    for (int lvl = 4; lvl >= 0; --lvl) {
        const int cout = kCh * kDecMult[lvl];
        for (int blk = 0; blk < kDecBlocks[lvl]; ++blk) {
            SHOWO_TRY(resblock(m, "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(blk), X, B, H, W, C,
                               cout, st));
            C = cout;
        }
        if (lvl != 0) {
            const int T = (X + 1) & 3, U = (X + 2) & 3;
            const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
            const int g = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
            upsample2x_kernel<<<g, 256, 0, st>>>(m->bufs[X], m->bufs[T], B, H, W, C);
            note_launch();
            H *= 2; W *= 2;
            SHOWO_TRY(conv(m, "decoder.up." + std::to_string(lvl) + ".upsample.conv", m->bufs[T], m->bufs[U], nullptr, B, H,
                           W, st));
            X = U;
        }
    }
    const int T = (X + 1) & 3;
    SHOWO_TRY(gn(m, "decoder.norm_out", m->bufs[X], m->bufs[T], B, H, W, C, true, st));
    const ConvW& co = m->conv["decoder.conv_out"];
    const int64_t np = (int64_t)B * H * W;
    conv_out3_kernel<<<(int)cdiv64(np, 256), 256, 27 * C * sizeof(float), st>>>(m->bufs[T], co.w32, co.b, out_nchw, out_u8, B,
                                                                               H, W, C);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    m->launches_last = launches_total() - l0;
    return 0;
}


// ------------------------------------------------------------------------------------------------ fp32 verification path (host side)
static int v_ensure(magvit_engine* m, size_t elems, cudaStream_t st) {
    if (elems <= m->vbuf_cap) return 0;
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; ++i) {
        if (m->vbufs[i]) cudaFree(m->vbufs[i]);
        m->vbufs[i] = nullptr;
        SHOWO_CUDA_OK(cudaMalloc(&m->vbufs[i], elems * sizeof(float)));
    }
    m->vbuf_cap = elems;
    return 0;
}
static int v_grid(int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 32); }
static int v_conv(magvit_engine* m, const std::string& name, const float* x, float* y, const float* resid, int NB, int H, int W, int stride,
                  cudaStream_t st) {
    auto it = m->conv.find(name);
    SHOWO_CHECK(it != m->conv.end(), "conv " + name + ": unknown");
    const ConvW& c = it->second;
    const int Ho = H / stride, Wo = W / stride;
    // Downsample (common_modules.py:83-87): pad (0, 1, 0, 1) then 3x3 stride 2 = no low-side padding; 3x3 stride 1: 'same'; 1x1: none
    const int pad_lo = (c.k == 3 && stride == 1) ? 1 : 0;
    vconv_kernel<<<v_grid((int64_t)NB * c.cout * Ho * Wo), 256, 0, st>>>(x, c.w32, c.b, resid, y, NB, c.cin, H, W, c.cout, Ho, Wo, c.k, stride, pad_lo);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int v_gn(magvit_engine* m, const std::string& name, const float* x, float* y, int NB, int C, int HW, bool swish, cudaStream_t st) {
    auto it = m->norm.find(name);
    SHOWO_CHECK(it != m->norm.end() && it->second.c == C, "groupnorm " + name + ": unknown or channel mismatch");
    vgn_kernel<<<NB * 32, 256, 0, st>>>(x, it->second.g, it->second.b, y, C, HW, 1e-6f, swish ? 1 : 0);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}
static int v_res(magvit_engine* m, const std::string& p, int& X, int NB, int H, int W, int cin, int cout, cudaStream_t st) {
    int others[3], k = 0;
    for (int i = 0; i < 4; ++i) if (i != X) others[k++] = i;
    const int T = others[0], U = others[1], V = others[2];
    SHOWO_TRY(v_gn(m, p + ".norm1", m->vbufs[X], m->vbufs[T], NB, cin, H * W, true, st));
    SHOWO_TRY(v_conv(m, p + ".conv1", m->vbufs[T], m->vbufs[U], nullptr, NB, H, W, 1, st));
    SHOWO_TRY(v_gn(m, p + ".norm2", m->vbufs[U], m->vbufs[T], NB, cout, H * W, true, st));
    const float* skip = m->vbufs[X];
    if (cin != cout) {
        SHOWO_TRY(v_conv(m, p + ".nin_shortcut", m->vbufs[X], m->vbufs[V], nullptr, NB, H, W, 1, st));
        skip = m->vbufs[V];
    }
    SHOWO_TRY(v_conv(m, p + ".conv2", m->vbufs[T], m->vbufs[U], skip, NB, H, W, 1, st));
    X = U;
    return 0;
}
static int v_attn(magvit_engine* m, const std::string& p, int& X, int NB, int H, int W, int C, cudaStream_t st) {
    int others[3], k = 0;
    for (int i = 0; i < 4; ++i) if (i != X) others[k++] = i;
    const int T = others[0], U = others[1], V = others[2];
    const int HW = H * W;
    if ((size_t)HW * HW > m->vscores_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (m->vscores) cudaFree(m->vscores);
        m->vscores = nullptr;
        SHOWO_CUDA_OK(cudaMalloc(&m->vscores, (size_t)HW * HW * 4));
        m->vscores_cap = (size_t)HW * HW;
    }
    SHOWO_TRY(v_gn(m, p + ".norm", m->vbufs[X], m->vbufs[T], NB, C, HW, false, st));
    SHOWO_TRY(v_conv(m, p + ".qkv", m->vbufs[T], m->vbufs[U], nullptr, NB, H, W, 1, st));       // [NB, 3C, HW]
    for (int n = 0; n < NB; ++n) {
        const float* qkv = m->vbufs[U] + (size_t)n * 3 * C * HW;
        vattn_scores_kernel<<<v_grid((int64_t)HW * HW), 256, 0, st>>>(qkv, m->vscores, C, HW, 1.0f / sqrtf((float)C));
        vsoftmax_kernel<<<HW, 256, 0, st>>>(m->vscores, HW);
        vattn_out_kernel<<<v_grid((int64_t)C * HW), 256, 0, st>>>(qkv, m->vscores, m->vbufs[T] + (size_t)n * C * HW, C, HW);
        note_launch(3);
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    SHOWO_TRY(v_conv(m, p + ".proj_out", m->vbufs[T], m->vbufs[V], m->vbufs[X], NB, H, W, 1, st));
    X = V;
    return 0;
}

extern "C" {

int magvit_engine_create(int device, magvit_engine_t** out) {
    SHOWO_CHECK(out, "null argument");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CUDA_OK(cudaSetDevice(device));
    magvit_engine* m = new magvit_engine();
    m->device = device;
    // ---- encoder (modeling_magvitv2.py:59-141)
    reg_conv(m, "encoder.conv_in", kCh, 3, 3);
    int bi = kCh;
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int bo = kCh * kEncMult[lvl];
        for (int blk = 0; blk < kEncBlocks[lvl]; ++blk) {
            reg_res(m, "encoder.down." + std::to_string(lvl) + ".block." + std::to_string(blk), bi, bo);
            bi = bo;
        }
        if (lvl != 4) reg_conv(m, "encoder.down." + std::to_string(lvl) + ".downsample.conv", bi, bi, 3, true);
    }
    reg_res(m, "encoder.mid.block_1", bi, bi); reg_attn(m, "encoder.mid.attn_1", bi); reg_res(m, "encoder.mid.block_2", bi, bi);
    reg_norm(m, "encoder.norm_out", bi); reg_conv(m, "encoder.conv_out", kZ, bi, 3);
    reg_conv(m, "encoder.quant_conv", kZ, kZ, 1);
    // ---- decoder (:278-362)
    bi = kCh * kDecMult[4];
    reg_conv(m, "decoder.conv_in", bi, kZ, 3);
    reg_res(m, "decoder.mid.block_1", bi, bi); reg_attn(m, "decoder.mid.attn_1", bi); reg_res(m, "decoder.mid.block_2", bi, bi);
    for (int lvl = 4; lvl >= 0; --lvl) {
        const int bo = kCh * kDecMult[lvl];
        for (int blk = 0; blk < kDecBlocks[lvl]; ++blk) {
            reg_res(m, "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(blk), bi, bo);
            bi = bo;
        }
        if (lvl != 0) reg_conv(m, "decoder.up." + std::to_string(lvl) + ".upsample.conv", bi, bi, 3);
    }
    reg_norm(m, "decoder.norm_out", bi); reg_conv(m, "decoder.conv_out", 3, bi, 3);
    reg_conv(m, "decoder.post_quant_conv", kZ, kZ, 1);
    SHOWO_TRY(magvit_alloc_all(m));
    SHOWO_CUDA_OK(cudaMalloc(&m->stats, 32 * 64 * 4));
    SHOWO_CUDA_OK(cudaMalloc(&m->partials, (size_t)32 * 128 * 64 * 4));
    *out = m;
    return 0;
}

int magvit_engine_destroy(magvit_engine_t* m) {
    if (!m) return 0;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (auto& kv : m->conv) { cudaFree(kv.second.w); cudaFree(kv.second.b); cudaFree(kv.second.w32); }
    for (auto& kv : m->norm) { cudaFree(kv.second.g); cudaFree(kv.second.b); }
    for (int i = 0; i < 4; ++i) if (m->bufs[i]) cudaFree(m->bufs[i]);
    if (m->stage) cudaFree(m->stage);
    if (m->stats) cudaFree(m->stats);
    if (m->partials) cudaFree(m->partials);
    if (m->scores) cudaFree(m->scores); if (m->probs) cudaFree(m->probs); if (m->vt) cudaFree(m->vt);
    for (int i = 0; i < 4; ++i) if (m->vbufs[i]) cudaFree(m->vbufs[i]);
    if (m->vscores) cudaFree(m->vscores);
    delete m;
    return 0;
}

int magvit_load_weight(magvit_engine_t* m, const char* name_c, const float* data, int64_t numel, int is_device) {
    SHOWO_CHECK(m && name_c && data, "null argument");
    SHOWO_CUDA_OK(cudaSetDevice(m->device));
    std::string name(name_c);
    const float* src = data;
    if (!is_device) {
        if ((size_t)numel > m->stage_cap) {
            if (m->stage) cudaFree(m->stage);
            SHOWO_CUDA_OK(cudaMalloc(&m->stage, (size_t)numel * 4));
            m->stage_cap = (size_t)numel;
        }
        SHOWO_CUDA_OK(cudaMemcpy(m->stage, data, (size_t)numel * 4, cudaMemcpyHostToDevice));
        src = m->stage;
    }
    const bool is_w = name.size() > 7 && name.compare(name.size() - 7, 7, ".weight") == 0;
    const bool is_b = name.size() > 5 && name.compare(name.size() - 5, 5, ".bias") == 0;
    SHOWO_CHECK(is_w || is_b, "magvit: unknown tensor " + name);
    std::string base = name.substr(0, name.size() - (is_w ? 7 : 5));
    // q / k / v of an AttnBlock are fused into one conv
    int qkv_row = -1;
    for (int i = 0; i < 3; ++i) {
        const std::string suf = std::string(".") + "qkv"[i];
        if (base.size() > 2 && base.compare(base.size() - 2, 2, suf) == 0 && base.find("attn_1") != std::string::npos) {
            qkv_row = i;
            base = base.substr(0, base.size() - 2) + ".qkv";
        }
    }
    auto ni = m->norm.find(base);
    if (ni != m->norm.end()) {
        SHOWO_CHECK(numel == ni->second.c, "magvit: size mismatch for " + name);
        SHOWO_CUDA_OK(cudaMemcpy(is_w ? ni->second.g : ni->second.b, src, (size_t)numel * 4, cudaMemcpyDeviceToDevice));
        m->loaded.insert(name);
        return 0;
    }
    auto ci = m->conv.find(base);
    SHOWO_CHECK(ci != m->conv.end(), "magvit: unknown tensor " + name);
    ConvW& c = ci->second;
    const int cout = qkv_row >= 0 ? c.cout / 3 : c.cout;
    const int row0 = qkv_row >= 0 ? qkv_row * cout : 0;
    if (is_b) {
        SHOWO_CHECK(numel == cout, "magvit: size mismatch for " + name);
        SHOWO_CUDA_OK(cudaMemcpy(c.b + row0, src, (size_t)numel * 4, cudaMemcpyDeviceToDevice));
    } else {
        SHOWO_CHECK(numel == (int64_t)cout * c.cin * c.k * c.k, "magvit: size mismatch for " + name);
        SHOWO_CUDA_OK(cudaMemcpy(c.w32 + (size_t)row0 * c.cin * c.k * c.k, src, (size_t)numel * 4, cudaMemcpyDeviceToDevice));
        const int grid = (int)((numel + 255) / 256 < 2048 ? (numel + 255) / 256 : 2048);
        if (c.down) pack_down_kernel<<<grid, 256>>>(src, c.w, cout, c.cin);
        else pack_conv_kernel<<<grid, 256>>>(src, c.w, cout, c.cin, c.k, c.cin_pad, row0, c.ld);
        SHOWO_CUDA_OK(cudaGetLastError());
        SHOWO_CUDA_OK(cudaDeviceSynchronize());
    }
    m->loaded.insert(name);
    return 0;
}

int magvit_weights_complete(magvit_engine_t* m) {
    SHOWO_CHECK(m, "null engine");
    for (auto& kv : m->norm) {
        SHOWO_CHECK(m->loaded.count(kv.first + ".weight") && m->loaded.count(kv.first + ".bias"),
                    "magvit weight not loaded: " + kv.first);
    }
    for (auto& kv : m->conv) {
        std::string b = kv.first;
        if (b.size() > 4 && b.compare(b.size() - 4, 4, ".qkv") == 0) {
            const std::string p = b.substr(0, b.size() - 4);
            for (const char* s : {".q", ".k", ".v"})
                SHOWO_CHECK(m->loaded.count(p + s + ".weight") && m->loaded.count(p + s + ".bias"),
                            "magvit weight not loaded: " + p + s);
        } else {
            SHOWO_CHECK(m->loaded.count(b + ".weight") && m->loaded.count(b + ".bias"), "magvit weight not loaded: " + b);
        }
    }
    return 0;
}

int magvit_decode_code(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, float* pixels_out_dev,
                       void* stream) {
    SHOWO_CHECK(pixels_out_dev, "null output");
    return decode_impl(m, ids_dev, B, h, w, pixels_out_dev, nullptr, (cudaStream_t)stream);
}
int magvit_decode_code_u8(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, uint8_t* images_out_dev,
                          void* stream) {
    SHOWO_CHECK(images_out_dev, "null output");
    return decode_impl(m, ids_dev, B, h, w, nullptr, images_out_dev, (cudaStream_t)stream);
}

int magvit_get_code(magvit_engine_t* m, const float* pixels_dev, int B, int R, int64_t* ids_out_dev, void* stream) {
    SHOWO_CHECK(m && pixels_dev && ids_out_dev && B > 0, "magvit get_code: bad arguments");
    SHOWO_CHECK(R % 256 == 0 && R >= 256, "magvit get_code: resolution must be a multiple of 256");
    SHOWO_CUDA_OK(cudaSetDevice(m->device));
    SHOWO_TRY(magvit_weights_complete(m));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    SHOWO_TRY(ensure_bufs(m, (size_t)B * R * R * kCh, st));
    int H = R, W = R, C = kCh;
    {
        const ConvW& ci = m->conv["encoder.conv_in"];
        const int64_t total = (int64_t)B * H * W * (C / 8);
        conv_in3_kernel<<<(int)cdiv64(total, 256), 256, 27 * C * sizeof(float), st>>>(pixels_dev, ci.w32, ci.b, m->bufs[0], B, H,
                                                                                     W, C);
        note_launch();
        SHOWO_CUDA_OK(cudaGetLastError());
    }
    int X = 0;
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int cout = kCh * kEncMult[lvl];
        for (int blk = 0; blk < kEncBlocks[lvl]; ++blk) {
            SHOWO_TRY(resblock(m, "encoder.down." + std::to_string(lvl) + ".block." + std::to_string(blk), X, B, H, W, C,
                               cout, st));
            C = cout;
        }
        if (lvl != 4) {
            // Downsample: pad (0,1,0,1) + 3x3 stride 2  ==  2x2 stride-1 "conv" over the space-to-depth tensor [H/2,W/2,4C]
            const int T = (X + 1) & 3, U = (X + 2) & 3;
            const int64_t total = (int64_t)B * (H / 2) * (W / 2) * 4 * (C / 8);
            const int g = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
            space_to_depth_kernel<<<g, 256, 0, st>>>(m->bufs[X], m->bufs[T], B, H, W, C);
            note_launch();
            H /= 2; W /= 2;
            const ConvW& c = m->conv["encoder.down." + std::to_string(lvl) + ".downsample.conv"];
            ConvArgs a{};
            a.x = m->bufs[T]; a.w = c.w; a.bias = c.b; a.resid = nullptr; a.out = m->bufs[U]; a.ldc = c.cout; a.ldr = c.cout;
            a.NB = B; a.H = H; a.W = W; a.cin = 4 * C; a.cout = c.cout; a.taps = 4;
            SHOWO_TRY(conv_nhwc_bf16(a, st));
            X = U;
        }
    }
    SHOWO_TRY(resblock(m, "encoder.mid.block_1", X, B, H, W, C, C, st));
    SHOWO_TRY(attnblock(m, "encoder.mid.attn_1", X, B, H, W, C, st));
    SHOWO_TRY(resblock(m, "encoder.mid.block_2", X, B, H, W, C, C, st));
    const int T = (X + 1) & 3, U = (X + 2) & 3;
    SHOWO_TRY(gn(m, "encoder.norm_out", m->bufs[X], m->bufs[T], B, H, W, C, true, st));
    SHOWO_TRY(conv(m, "encoder.conv_out", m->bufs[T], m->bufs[U], nullptr, B, H, W, st));     // [B*H*W, 13]
    const ConvW& qc = m->conv["encoder.quant_conv"];
    const int64_t npix = (int64_t)B * H * W;
    lfq_encode_kernel<<<(int)cdiv64(npix, 128), 128, 0, st>>>(m->bufs[U], kZ, qc.w32, qc.b, ids_out_dev, npix);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    m->launches_last = launches_total() - l0;
    return 0;
}

// MAGVITv2.decode_code on the fp32 verification path: ids_dev [B, h*w] -> pixels_out_dev [B, 3, 16h, 16w] fp32 (NCHW)
int magvit_decode_code_fp32(magvit_engine_t* m, const int64_t* ids_dev, int B, int h, int w, float* pixels_out_dev, void* stream) {
    SHOWO_CHECK(m && ids_dev && pixels_out_dev && B > 0 && h > 0 && w > 0, "magvit decode (fp32): bad arguments");
    SHOWO_CUDA_OK(cudaSetDevice(m->device));
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_TRY(magvit_weights_complete(m));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    SHOWO_TRY(v_ensure(m, (size_t)B * kCh * (16 * h) * (16 * w), st));
    int H = h, W = w;
    vlfq_entry_kernel<<<v_grid((int64_t)B * 13 * h * w), 256, 0, st>>>(ids_dev, m->vbufs[0], B, h * w);
    note_launch();
    SHOWO_TRY(v_conv(m, "decoder.post_quant_conv", m->vbufs[0], m->vbufs[1], nullptr, B, H, W, 1, st));
    SHOWO_TRY(v_conv(m, "decoder.conv_in", m->vbufs[1], m->vbufs[0], nullptr, B, H, W, 1, st));
    int X = 0;
    int C = kCh * kDecMult[4];
    SHOWO_TRY(v_res(m, "decoder.mid.block_1", X, B, H, W, C, C, st));
    SHOWO_TRY(v_attn(m, "decoder.mid.attn_1", X, B, H, W, C, st));
    SHOWO_TRY(v_res(m, "decoder.mid.block_2", X, B, H, W, C, C, st));
    for (int lvl = 4; lvl >= 0; --lvl) {
        const int cout = kCh * kDecMult[lvl];
        for (int blk = 0; blk < kDecBlocks[lvl]; ++blk) {
            SHOWO_TRY(v_res(m, "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(blk), X, B, H, W, C, cout, st));
            C = cout;
        }
        if (lvl != 0) {
            const int T = (X + 1) & 3, U = (X + 2) & 3;
            vupsample_kernel<<<v_grid((int64_t)B * C * 4 * H * W), 256, 0, st>>>(m->vbufs[X], m->vbufs[T], (int64_t)B * C, H, W);
            note_launch();
            H *= 2; W *= 2;
            SHOWO_TRY(v_conv(m, "decoder.up." + std::to_string(lvl) + ".upsample.conv", m->vbufs[T], m->vbufs[U], nullptr, B, H, W, 1, st));
            X = U;
        }
    }
    const int T = (X + 1) & 3;
    SHOWO_TRY(v_gn(m, "decoder.norm_out", m->vbufs[X], m->vbufs[T], B, C, H * W, true, st));
    {
        auto it = m->conv.find("decoder.conv_out");
        const ConvW& c = it->second;
        vconv_kernel<<<v_grid((int64_t)B * 3 * H * W), 256, 0, st>>>(m->vbufs[T], c.w32, c.b, nullptr, pixels_out_dev, B, c.cin, H, W, 3, H, W, 3, 1, 1);
        note_launch();
    }
    SHOWO_CUDA_OK(cudaGetLastError());
    m->launches_last = launches_total() - l0;
    return 0;
}

// MAGVITv2.get_code on the fp32 verification path: pixels_dev [B, 3, R, R] -> ids_out_dev [B, (R/16)^2]; z_out_dev (optional)
// receives the quantizer's pre-sign values [B, 13, R/16, R/16] -- what the sign of every code bit is taken from
int magvit_get_code_fp32(magvit_engine_t* m, const float* pixels_dev, int B, int R, int64_t* ids_out_dev, float* z_out_dev, void* stream) {
    SHOWO_CHECK(m && pixels_dev && ids_out_dev && B > 0 && R >= 16 && R % 16 == 0, "magvit get_code (fp32): bad arguments");
    SHOWO_CUDA_OK(cudaSetDevice(m->device));
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_TRY(magvit_weights_complete(m));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    SHOWO_TRY(v_ensure(m, (size_t)B * kCh * R * R, st));
    int H = R, W = R, C = kCh;
    {
        const ConvW& c = m->conv["encoder.conv_in"];
        vconv_kernel<<<v_grid((int64_t)B * C * H * W), 256, 0, st>>>(pixels_dev, c.w32, c.b, nullptr, m->vbufs[0], B, 3, H, W, C, H, W, 3, 1, 1);
        note_launch();
    }
    int X = 0;
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int cout = kCh * kEncMult[lvl];
        for (int blk = 0; blk < kEncBlocks[lvl]; ++blk) {
            SHOWO_TRY(v_res(m, "encoder.down." + std::to_string(lvl) + ".block." + std::to_string(blk), X, B, H, W, C, cout, st));
            C = cout;
        }
        if (lvl != 4) {
            const int U = (X + 1) & 3;
            SHOWO_TRY(v_conv(m, "encoder.down." + std::to_string(lvl) + ".downsample.conv", m->vbufs[X], m->vbufs[U], nullptr, B, H, W, 2, st));
            H /= 2; W /= 2;
            X = U;
        }
    }
    SHOWO_TRY(v_res(m, "encoder.mid.block_1", X, B, H, W, C, C, st));
    SHOWO_TRY(v_attn(m, "encoder.mid.attn_1", X, B, H, W, C, st));
    SHOWO_TRY(v_res(m, "encoder.mid.block_2", X, B, H, W, C, C, st));
    const int T = (X + 1) & 3, U = (X + 2) & 3;
    SHOWO_TRY(v_gn(m, "encoder.norm_out", m->vbufs[X], m->vbufs[T], B, C, H * W, true, st));
    SHOWO_TRY(v_conv(m, "encoder.conv_out", m->vbufs[T], m->vbufs[U], nullptr, B, H, W, 1, st));       // [B, 13, h, w]
    SHOWO_TRY(v_conv(m, "encoder.quant_conv", m->vbufs[U], m->vbufs[T], nullptr, B, H, W, 1, st));
    if (z_out_dev) SHOWO_CUDA_OK(cudaMemcpyAsync(z_out_dev, m->vbufs[T], (size_t)B * 13 * H * W * 4, cudaMemcpyDeviceToDevice, st));
    vlfq_index_kernel<<<v_grid((int64_t)B * H * W), 256, 0, st>>>(m->vbufs[T], ids_out_dev, B, H * W);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    m->launches_last = launches_total() - l0;
    return 0;
}

int64_t magvit_kernel_launches(magvit_engine_t* m) { return m ? m->launches_last : 0; }

}  // extern "C"
