// Fused MaskGIT sampler step (models/modeling_showo.py:143-179 + models/sampling.py:10-36), HBM-bound.
//
// kernel 1 (one CTA per image token): CFG combine of the cond/uncond image-vocab logits, fp32 softmax over the 8192
//   codes, categorical draw as argmax(p / Exp(1)) (== torch.multinomial(p,1)), gather of the selected
//   probability and confidence = log p + temperature * gumbel(U).  Logits are read exactly once (16 B per thread per
//   access, fully coalesced); nothing of size [*, 8192] is written back.
// kernel 2 (one CTA per batch row): mask_len = max(1, min(#unknown-1, floor(N*ratio))), the mask_len-th order statistic
//   of the N confidences by rank counting (== sort()[mask_len]), re-mask `confidence < cut_off`, write the ids back.
// Noise: host-supplied tensors in the order torch consumes them (parity mode) or counter-based Philox4x32-10.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

#include "philox.cuh"

namespace showo {

struct ArgMax { float v; int i; float p; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <int kThreads>
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) r = fmaxf(r, red[i]);
    __syncthreads();
    return r;
}
template <int kThreads>
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r += red[i];
    __syncthreads();
    return r;
}

__device__ __forceinline__ float gumbel_from_uniform(float u) {
    // -log(-log(clamp(u,1e-20)) clamped at 1e-20)      (sampling.py:10-16)
    const float inner = -logf(fmaxf(u, 1e-20f));
    return -logf(fmaxf(inner, 1e-20f));
}

constexpr int kSampThreads = 256;

__global__ void __launch_bounds__(kSampThreads) sampler_token_kernel(SamplerArgs a) {
    __shared__ float red[kSampThreads / 32];
    __shared__ ArgMax red_am[kSampThreads / 32];
    const int tok = blockIdx.x;
    const int b = tok / a.N, n = tok % a.N;
    const int tid = threadIdx.x;
    const int64_t cur = a.ids[(int64_t)b * a.ids_stride + a.ids_pos0 + n];
    const bool unknown = (cur == (int64_t)a.mask_token_id);

    float u_g;
    if (a.noise_unif) u_g = a.noise_unif[tok];
    else {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)(tok + a.row_base * a.N), 0u, a.step, 0x9u), make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        u_g = u01(r.x);
    }
    const float gum = gumbel_from_uniform(u_g);

    if (!unknown) {   // CTA-uniform: known tokens keep their code, confidence = log(finfo.max) + t*g
        if (tid == 0) {
            a.sampled_ws[tok] = (int)(cur - a.image_offset);
            a.conf_ws[tok] = __fadd_rn(logf(FLT_MAX), __fmul_rn(a.temperature, gum));
        }
        return;
    }

    const int64_t row = (int64_t)b * a.rows_per_seq + n;
    const float4* lc = reinterpret_cast<const float4*>(a.logits_cond + row * a.ld);
    const float4* lu = a.logits_uncond ? reinterpret_cast<const float4*>(a.logits_uncond + row * a.ld) : nullptr;
    const int nvec = a.C >> 2;                  // float4 per row
    const float w1 = 1.0f + a.guidance, w = a.guidance;
    float4 z[8];
    float mx = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = tid + i * kSampThreads;
        if (j < nvec) {
            float4 c = __ldg(lc + j);
            if (lu) {
                const float4 u = __ldg(lu + j);
                c.x = __fsub_rn(__fmul_rn(w1, c.x), __fmul_rn(w, u.x));
                c.y = __fsub_rn(__fmul_rn(w1, c.y), __fmul_rn(w, u.y));
                c.z = __fsub_rn(__fmul_rn(w1, c.z), __fmul_rn(w, u.z));
                c.w = __fsub_rn(__fmul_rn(w1, c.w), __fmul_rn(w, u.w));
            }
            z[i] = c;
            mx = fmaxf(mx, fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)));
        }
    }
    mx = block_reduce_max<kSampThreads>(mx, red);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = tid + i * kSampThreads;
        if (j < nvec) {
            z[i].x = expf(z[i].x - mx); z[i].y = expf(z[i].y - mx);
            z[i].z = expf(z[i].z - mx); z[i].w = expf(z[i].w - mx);
            sum += (z[i].x + z[i].y) + (z[i].z + z[i].w);
        }
    }
    sum = block_reduce_sum<kSampThreads>(sum, red);

    ArgMax best{-1.f, 0x7fffffff, 0.f};
    const float4* ex = a.noise_expo ? reinterpret_cast<const float4*>(a.noise_expo + (int64_t)tok * a.C) : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = tid + i * kSampThreads;
        if (j < nvec) {
            float4 q;
            if (ex) q = __ldg(ex + j);
            else {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)((tok + a.row_base * a.N) * nvec + j), 0u, a.step, 0x5u),
                                              make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                q = make_float4(-logf(u01(r.x)), -logf(u01(r.y)), -logf(u01(r.z)), -logf(u01(r.w)));
            }
            const float p0 = __fdiv_rn(z[i].x, sum), p1 = __fdiv_rn(z[i].y, sum), p2 = __fdiv_rn(z[i].z, sum),
                        p3 = __fdiv_rn(z[i].w, sum);
            best = better(best, ArgMax{__fdiv_rn(p0, q.x), 4 * j + 0, p0});
            best = better(best, ArgMax{__fdiv_rn(p1, q.y), 4 * j + 1, p1});
            best = better(best, ArgMax{__fdiv_rn(p2, q.z), 4 * j + 2, p2});
            best = better(best, ArgMax{__fdiv_rn(p3, q.w), 4 * j + 3, p3});
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        other.p = __shfl_xor_sync(0xffffffffu, best.p, o);
        best = better(best, other);
    }
    if ((tid & 31) == 0) red_am[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        ArgMax r = red_am[0];
#pragma unroll
        for (int i = 1; i < kSampThreads / 32; ++i) r = better(r, red_am[i]);
        a.sampled_ws[tok] = r.i;
        a.conf_ws[tok] = __fadd_rn(logf(fmaxf(r.p, 1e-20f)), __fmul_rn(a.temperature, gum));
    }
}

__global__ void __launch_bounds__(1024) sampler_row_kernel(SamplerArgs a) {
    extern __shared__ float conf[];     // [N]
    __shared__ int s_unknown;
    __shared__ float s_cut;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) { s_unknown = 0; s_cut = -FLT_MAX; }
    __syncthreads();
    int my_unknown = 0;
    for (int i = tid; i < a.N; i += blockDim.x) {
        conf[i] = a.conf_ws[b * a.N + i];
        my_unknown += (a.ids[(int64_t)b * a.ids_stride + a.ids_pos0 + i] == (int64_t)a.mask_token_id) ? 1 : 0;
    }
    if (my_unknown) atomicAdd(&s_unknown, my_unknown);
    __syncthreads();
    int mask_len = min(s_unknown - 1, a.mask_len_floor);      // modeling_showo.py:166-171
    mask_len = max(1, mask_len);
    mask_len = min(mask_len, a.N - 1);
    for (int i = tid; i < a.N; i += blockDim.x) {
        const float c = conf[i];
        int less = 0, leq = 0;
        for (int j = 0; j < a.N; ++j) {
            const float o = conf[j];
            less += (o < c) ? 1 : 0;
            leq += (o <= c) ? 1 : 0;
        }
        if (less <= mask_len && mask_len < leq) s_cut = c;   // every writer holds the same value
    }
    __syncthreads();
    const float cut = s_cut;
    for (int i = tid; i < a.N; i += blockDim.x) {
        const bool masking = conf[i] < cut;                   // strict: ties are kept (sampling.py:35)
        const int code = a.sampled_ws[b * a.N + i];
        const int64_t nid = masking ? (int64_t)a.mask_token_id : (int64_t)code + a.image_offset;
        a.ids[(int64_t)b * a.ids_stride + a.ids_pos0 + i] = nid;
        if (a.ids2) a.ids2[(int64_t)b * a.ids2_stride + a.ids_pos0 + i] = nid;
        if (a.sampled_out) a.sampled_out[(int64_t)b * a.N + i] = code;
        if (a.masking_out) a.masking_out[b * a.N + i] = masking ? 1 : 0;
    }
}

int t2i_sampler_step(const SamplerArgs& a, cudaStream_t st) {
    SHOWO_CHECK(a.C % 4 == 0 && a.C <= 8 * 4 * kSampThreads, "sampler: codebook size must be a multiple of 4 and <= 8192");
    SHOWO_CHECK((a.ld % 4) == 0, "sampler: logits row stride must be a multiple of 4 floats");
    SHOWO_CHECK(a.N * (int)sizeof(float) <= 48 * 1024, "sampler: N too large");
    if (a.B == 0) return 0;
    sampler_token_kernel<<<a.B * a.N, kSampThreads, 0, st>>>(a);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    const int thr = a.N >= 1024 ? 1024 : ((a.N + 31) / 32) * 32;
    sampler_row_kernel<<<a.B, thr, a.N * sizeof(float), st>>>(a);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- greedy next token (mmu_generate with top_k=1)
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int64_t ld, int V,
                                                           int64_t* __restrict__ out) {
    __shared__ ArgMax red_am[32];
    const float* row = logits + (int64_t)blockIdx.x * ld;
    ArgMax best{-FLT_MAX, 0x7fffffff, 0.f};
    for (int i = threadIdx.x; i < V; i += blockDim.x) best = better(best, ArgMax{row[i], i, 0.f});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        other.p = 0.f;
        best = better(best, other);
    }
    if ((threadIdx.x & 31) == 0) red_am[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        ArgMax r = red_am[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = better(r, red_am[i]);
        out[blockIdx.x] = r.i;
    }
}
int argmax_rows(const float* logits, int64_t ld, int B, int V, int64_t* out, cudaStream_t st) {
    if (B == 0) return 0;
    argmax_rows_kernel<<<B, 1024, 0, st>>>(logits, ld, V, out);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------- mmu_generate next-token draw
// modeling_showo.py:219-228: logits[:, -1] / temperature -> optional top-k filter (everything below the k-th largest
// value becomes -inf, ties kept) -> softmax -> multinomial(probs, 1) (== argmax(p / Exp(1))).  One CTA per row; the row
// (V = 58498 floats) stays in L2 across the passes.  The k-th largest value comes from a 4 x 8-bit radix select over
// order-preserving integer keys, not a sort.
__device__ __forceinline__ uint32_t float_order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024) mmu_sample_kernel(MmuSampleArgs a) {
    __shared__ float red[32];
    __shared__ ArgMax red_am[32];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining;
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* lr = a.logits + (int64_t)row * a.ld;
    float mx = -FLT_MAX;
    for (int i = tid; i < a.V; i += 1024) mx = fmaxf(mx, __fdiv_rn(lr[i], a.temperature));
    mx = block_reduce_max<1024>(mx, red);

    uint32_t kth = 0u;                            // keep keys >= kth (0 keeps everything)
    if (a.top_k > 0 && a.top_k < a.V) {
        if (tid == 0) { s_prefix = 0u; s_remaining = (uint32_t)a.top_k; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += 1024) hist[i] = 0u;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t hi_mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (int i = tid; i < a.V; i += 1024) {
                const uint32_t key = float_order_key(__fdiv_rn(lr[i], a.temperature));
                if ((key & hi_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t cum = 0u, rem = s_remaining;
                for (int b = 255; b >= 0; --b) {
                    if (cum + hist[b] >= rem) { s_prefix = prefix | ((uint32_t)b << shift); s_remaining = rem - cum; break; }
                    cum += hist[b];
                }
            }
            __syncthreads();
        }
        kth = s_prefix;
    }
    float sum = 0.f;
    for (int i = tid; i < a.V; i += 1024) {
        const float l = __fdiv_rn(lr[i], a.temperature);
        if (float_order_key(l) >= kth) sum += expf(l - mx);
    }
    sum = block_reduce_sum<1024>(sum, red);

    ArgMax best{-1.f, 0x7fffffff, 0.f};
    const float* ex = a.noise_expo ? a.noise_expo + (int64_t)row * a.V : nullptr;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    for (int i = tid; i < a.V; i += 1024) {
        const float l = __fdiv_rn(lr[i], a.temperature);
        if (float_order_key(l) < kth) continue;   // filtered: p = 0 never wins against a kept entry
        float q;
        if (ex) q = ex[i];
        else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(row + a.row_base), a.step, 0xAu), key);
            q = -logf(u01(r.x));
        }
        const float pr = __fdiv_rn(expf(l - mx), sum);
        best = better(best, ArgMax{__fdiv_rn(pr, q), i, pr});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        other.p = 0.f;
        best = better(best, other);
    }
    if ((tid & 31) == 0) red_am[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        ArgMax r = red_am[0];
        for (int i = 1; i < 32; ++i) r = better(r, red_am[i]);
        a.out[(int64_t)row * a.out_stride] = r.i;
        if (a.out_next) a.out_next[row] = r.i;
    }
}
int mmu_sample(const MmuSampleArgs& a, cudaStream_t st) {
    if (a.B == 0) return 0;
    SHOWO_CHECK(a.temperature > 0.f && a.V > 0, "mmu_sample: temperature must be positive");
    mmu_sample_kernel<<<a.B, 1024, 0, st>>>(a);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace showo
