// Omni-mask attention (causal over text, bidirectional inside image spans, always-visible windows, left-pad removal).
//
// The mask is never materialised: each (query, key) pair evaluates the closed-form predicate of showo_seq_mask_t in
// registers, and whole 64-key tiles that the predicate rules out for the CTA's query range are skipped.
//
// Prefill / denoise-step kernel: flash-style, one CTA = 64 query rows of one (sequence, head); K tiles [64 keys][64]
// and V^T tiles [64 dims][64 keys] stream through a double-buffered cp.async pipeline; QK^T and PV run on the tensor
// cores (mma.sync m16n8k16 bf16, fp32 accumulate), softmax in fp32 with exp2.  head_dim is fixed at 64.
// Attention is ~3 % of the step FLOPs at L=387, the tcgen05 GEMMs carry the rest.
//
// Decode kernel: one query per sequence against the KV cache (mmu_generate), HBM-bound, CUDA cores.
#include "attn_common.cuh"

namespace showo {

__global__ void __launch_bounds__(128) omni_attention_kernel(AttnArgs a) {
    __shared__ __align__(16) bf16 Ks[2][kTileK][kPad];
    __shared__ __align__(16) bf16 Vs[2][64][kPad];
    pdl_trigger();
    pdl_wait();

    const int seq = blockIdx.z, h = blockIdx.y;
    const int q0 = a.row_begin + blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const showo_seq_mask_t msk = a.masks[seq];

    // ---- Q fragments (rows r0 = q0 + warp*16 + g and r0 + 8), straight from global
    const int r0 = q0 + warp * 16 + g;
    const int r1 = r0 + 8;
    const bool r0_ok = r0 < a.rows_per_seq, r1_ok = r1 < a.rows_per_seq;
    bf16* qrow0 = a.q + ((int64_t)seq * a.rows_per_seq + r0) * a.ld + h * 64;
    bf16* qrow1 = a.q + ((int64_t)seq * a.rows_per_seq + r1) * a.ld + h * 64;
    uint32_t qf[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int c = kk * 16 + t4 * 2;
        qf[kk][0] = r0_ok ? *reinterpret_cast<const uint32_t*>(qrow0 + c) : 0u;
        qf[kk][1] = r1_ok ? *reinterpret_cast<const uint32_t*>(qrow1 + c) : 0u;
        qf[kk][2] = r0_ok ? *reinterpret_cast<const uint32_t*>(qrow0 + c + 8) : 0u;
        qf[kk][3] = r1_ok ? *reinterpret_cast<const uint32_t*>(qrow1 + c + 8) : 0u;
    }
    const int qpos0 = a.pos0 + r0, qpos1 = a.pos0 + r1;
    // warps whose 16 rows all lie past the sequence's rows (the tail tile of 258 = 4*64 + 2) only help with the loads
    const bool warp_active = (q0 + warp * 16) < a.rows_per_seq;
    const int wq_lo = a.pos0 + q0 + warp * 16;
    const int wq_hi = a.pos0 + min(q0 + warp * 16 + 15, a.rows_per_seq - 1);
    const int cta_q_lo = a.pos0 + q0;
    const int cta_q_hi = a.pos0 + min(q0 + 63, a.rows_per_seq - 1);

    const bf16* kbase = a.kcache + ((int64_t)seq * a.H + h) * (int64_t)a.Lmax * 64;
    const bf16* vbase = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax;
    const int n_tiles = (a.n_keys + kTileK - 1) / kTileK;

    auto next_tile = [&](int kt) {
        while (kt < n_tiles &&
               !omni_tile_possible(msk, cta_q_lo, cta_q_hi, kt * kTileK, min((kt + 1) * kTileK, a.n_keys)))
            ++kt;
        return kt;
    };
    auto load_tile = [&](int kt, int buf) {
        const int k0 = kt * kTileK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = threadIdx.x + i * 128;   // 0..511
            const int row = idx >> 3, ch = idx & 7;
            cp_async16(&Ks[buf][row][ch * 8], kbase + (int64_t)(k0 + row) * 64 + ch * 8, 16);
            cp_async16(&Vs[buf][row][ch * 8], vbase + (int64_t)row * a.Lmax + k0 + ch * 8, 16);
        }
        cp_async_commit();
    };

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = kNegBig, m1 = kNegBig, l0 = 0.f, l1 = 0.f;
    const float sc = a.scale * 1.4426950408889634f;

    int kt = next_tile(0);
    int buf = 0;
    if (kt < n_tiles) load_tile(kt, 0);
    while (kt < n_tiles) {
        const int kt_next = next_tile(kt + 1);
        if (kt_next < n_tiles) {
            load_tile(kt_next, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // ---- S = Q K^T
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
        if (warp_active) {
        // one ldmatrix.x4 feeds two MMAs: matrices = (key block nb, d 0..7), (nb, d 8..15), (nb+1, d 0..7), (nb+1, d 8..15)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nb = 0; nb < 8; nb += 2) {
                uint32_t b[4];
                ldmatrix_x4(b, &Ks[buf][(nb + (lane >> 4)) * 8 + (lane & 7)][kk * 16 + ((lane >> 3) & 1) * 8]);
                mma_bf16_16816(s[nb], qf[kk], b[0], b[1]);
                mma_bf16_16816(s[nb + 1], qf[kk], b[2], b[3]);
            }
        }
        // ---- mask (only on tiles the predicate does not fully allow for this warp's 16 rows) + online softmax.
        //      Maxima are tracked on the raw scores; p = exp2(s*c - m*c) is one FFMA + one MUFU per element.
        const int k0 = kt * kTileK;
        if (!omni_tile_all_allowed(msk, wq_lo, wq_hi, k0, k0 + kTileK, a.n_keys)) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = k0 + nb * 8 + t4 * 2 + (e & 1);
                    const int qp = (e < 2) ? qpos0 : qpos1;
                    const bool ok = (col < a.n_keys) && omni_allowed(msk, qp, col);
                    s[nb][e] = ok ? s[nb][e] : kNegBig;
                }
            }
        }
        float tm0 = kNegBig, tm1 = kNegBig;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            tm0 = fmaxf(tm0, fmaxf(s[nb][0], s[nb][1]));
            tm1 = fmaxf(tm1, fmaxf(s[nb][2], s[nb][3]));
        }
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1));
        tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1));
        tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
        const float mn0 = fmaxf(m0, tm0), mn1 = fmaxf(m1, tm1);
        const float al0 = ex2_approx((m0 - mn0) * sc), al1 = ex2_approx((m1 - mn1) * sc);
        m0 = mn0; m1 = mn1;
        // rows that have not met an allowed key yet keep offset 0: fma(-1e30, c, +1e30*c) would NOT cancel exactly
        // (exact product vs rounded addend) and could overflow exp2; with offset 0 their p is exactly 0
        const float ms0 = (m0 == kNegBig) ? 0.f : -m0 * sc, ms1 = (m1 == kNegBig) ? 0.f : -m1 * sc;
        l0 *= al0; l1 *= al1;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) { o[nd][0] *= al0; o[nd][1] *= al0; o[nd][2] *= al1; o[nd][3] *= al1; }
        uint32_t pf[4][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = ex2_approx(fmaf(s[nb][0], sc, ms0)), p1 = ex2_approx(fmaf(s[nb][1], sc, ms0));
            const float p2 = ex2_approx(fmaf(s[nb][2], sc, ms1)), p3 = ex2_approx(fmaf(s[nb][3], sc, ms1));
            l0 += p0 + p1; l1 += p2 + p3;
            const int kk = nb >> 1;
            if ((nb & 1) == 0) { pf[kk][0] = pack_bf16(p0, p1); pf[kk][1] = pack_bf16(p2, p3); }
            else               { pf[kk][2] = pack_bf16(p0, p1); pf[kk][3] = pack_bf16(p2, p3); }
        }
        // ---- O += P V   (B operand = V^T tile: [dim][key], keys contiguous; same ldmatrix pattern)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int nd = 0; nd < 8; nd += 2) {
                uint32_t b[4];
                ldmatrix_x4(b, &Vs[buf][(nd + (lane >> 4)) * 8 + (lane & 7)][kk * 16 + ((lane >> 3) & 1) * 8]);
                mma_bf16_16816(o[nd], pf[kk], b[0], b[1]);
                mma_bf16_16816(o[nd + 1], pf[kk], b[2], b[3]);
            }
        }
        }  // warp_active
        __syncthreads();
        buf ^= 1;
        kt = kt_next;
    }

    // ---- finalise: row sums across the quad, normalise, write over q
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    bf16* orow0 = a.out ? a.out + ((int64_t)seq * a.rows_per_seq + r0) * a.out_ld + h * 64 : qrow0;
    bf16* orow1 = a.out ? a.out + ((int64_t)seq * a.rows_per_seq + r1) * a.out_ld + h * 64 : qrow1;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        const int c = nd * 8 + t4 * 2;
        if (r0_ok) *reinterpret_cast<uint32_t*>(orow0 + c) = pack_bf16(o[nd][0] * i0, o[nd][1] * i0);
        if (r1_ok) *reinterpret_cast<uint32_t*>(orow1 + c) = pack_bf16(o[nd][2] * i1, o[nd][3] * i1);
    }
    if (a.lse != nullptr && t4 == 0) {          // exp2-domain log-sum-exp; a row without any allowed key gets +big (p = 0)
        if (r0_ok) a.lse[((int64_t)seq * a.rows_per_seq + r0) * a.H + h] = l0 > 0.f ? m0 * sc + log2f(l0) : 1.0e30f;
        if (r1_ok) a.lse[((int64_t)seq * a.rows_per_seq + r1) * a.H + h] = l1 > 0.f ? m1 * sc + log2f(l1) : 1.0e30f;
    }
}


int omni_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.n_seq == 0 || a.rows_per_seq == 0) return 0;
    SHOWO_CHECK(a.Lmax % 64 == 0, "attention: Lmax must be a multiple of 64");
    SHOWO_CHECK(a.n_keys <= a.Lmax, "attention: n_keys exceeds the cache length");
    // full 128-row tiles of every sequence on tcgen05 / TMEM; the ragged tail (and everything when the switch is off) on mma.sync
    AttnArgs b = a;
    b.row_begin = attention_tc_rows(a);
    const int n_tail = a.rows_per_seq - b.row_begin;
    if (b.row_begin > 0 && attention_tc_tail_rows(a) == n_tail) return omni_attention_tc(a, st);     // the kernel's tail phase takes them
    if (b.row_begin > 0) SHOWO_TRY(omni_attention_tc(a, st));
    if (b.row_begin >= a.rows_per_seq) return 0;
    dim3 grid(cdiv(a.rows_per_seq - b.row_begin, 64), a.H, a.n_seq);
    SHOWO_CUDA_OK(launch_kernel(omni_attention_kernel, grid, dim3(128), 0, st, 1, b));
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ decode (1 query / seq)
__global__ void __launch_bounds__(128) omni_attention_decode_kernel(AttnArgs a) {
    extern __shared__ float sc_s[];             // [n_keys]
    __shared__ float red[8];
    const int h = blockIdx.x, seq = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const showo_seq_mask_t msk = a.masks[seq];
    const int qpos = a.pos0;
    bf16* qrow = a.q + (int64_t)seq * a.rows_per_seq * a.ld + h * 64;
    float q[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 u = *reinterpret_cast<const uint4*>(qrow + i * 8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            q[i * 8 + 2 * j] = f.x; q[i * 8 + 2 * j + 1] = f.y;
        }
    }
    const bf16* kbase = a.kcache + ((int64_t)seq * a.H + h) * (int64_t)a.Lmax * 64;
    const bf16* vbase = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax;
    const float sc = a.scale * 1.4426950408889634f;
    float mx = kNegBig;
    for (int k = tid; k < a.n_keys; k += 128) {
        float s = kNegBig;
        if (omni_allowed(msk, qpos, k)) {
            float acc = 0.f;
            const uint4* kr = reinterpret_cast<const uint4*>(kbase + (int64_t)k * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 u = __ldg(kr + i);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    acc += q[i * 8 + 2 * j] * f.x + q[i * 8 + 2 * j + 1] * f.y;
                }
            }
            s = acc * sc;
        }
        sc_s[k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int k = tid; k < a.n_keys; k += 128) {
        const float p = exp2f(sc_s[k] - mx);
        sc_s[k] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + warp] = sum;
    __syncthreads();
    sum = red[4] + red[5] + red[6] + red[7];
    // o[d] = sum_k p[k] V^T[d][k]; two threads per d split the keys
    const int d = tid >> 1, half = tid & 1;
    const bf16* vr = vbase + (int64_t)d * a.Lmax;
    float acc = 0.f;
    const int n8 = a.n_keys >> 3;
    for (int c = half; c < n8; c += 2) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(vr) + c);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            acc += sc_s[c * 8 + 2 * j] * f.x + sc_s[c * 8 + 2 * j + 1] * f.y;
        }
    }
    if (half == 0)
        for (int k = n8 * 8; k < a.n_keys; ++k) acc += sc_s[k] * __bfloat162float(vr[k]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    __syncthreads();   // everyone has read q before it is overwritten
    if (half == 0) qrow[d] = __float2bfloat16(acc / sum);
}

// Bulk-copy variant (default): the whole K block of a (sequence, head) is contiguous in the cache ([n_keys][64] bf16) and
// each V^T row is a contiguous run, so ONE cp.async.bulk brings K and 64 row copies bring V^T -- all bytes of the CTA are
// in flight at once (70 KB at 276 keys, 2-3 CTAs per SM) instead of a few 16 B loads per thread, and the math then
// runs out of shared memory: 8 lanes per key (one 128 B row per quarter-warp, conflict-free), two threads per output dim.

__global__ void __launch_bounds__(128) omni_attention_decode_bulk_kernel(AttnArgs a, int n_pad, int vstride) {
    extern __shared__ __align__(128) uint8_t dec_smem[];
    bf16* Ks = reinterpret_cast<bf16*>(dec_smem);                                        // [n_pad][64]
    uint8_t* Vs = dec_smem + (size_t)n_pad * 128;                                        // [64][vstride bytes]
    float* sc_s = reinterpret_cast<float*>(Vs + (size_t)64 * vstride);                   // [n_pad]
    __shared__ float red[8];
    __shared__ __align__(8) uint64_t bars[2];
    const int h = blockIdx.x, seq = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    pdl_trigger();
    if (a.l2_prefetch_bytes)         // the successor GEMM's weights -> L2 while this kernel streams the cache
        l2_prefetch_slice(a.l2_prefetch, a.l2_prefetch_bytes, (int)(blockIdx.y * gridDim.x + blockIdx.x) * 128 + tid, (int)(gridDim.x * gridDim.y) * 128);
    // Only the row / column of the current token (position n_keys - 1) comes from the predecessor GEMM; everything older was
    // written by earlier decode steps, so those bytes are requested BEFORE griddepcontrol.wait and stream in while the
    // predecessor drains.
    const bf16* kbase = a.kcache + ((int64_t)seq * a.H + h) * (int64_t)a.Lmax * 64;
    const bf16* vbase = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax;
    const int old_keys = a.n_keys - 1, old_cols = n_pad - 8;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], (uint32_t)a.n_keys * 128u);
        if (old_keys > 0) bulk_g2s(Ks, kbase, (uint32_t)old_keys * 128u, &bars[0]);
        mbar_arrive_expect_tx(&bars[1], 64u * (uint32_t)n_pad * 2u);
    }
    __syncthreads();
    if (tid < 64 && old_cols > 0) bulk_g2s(Vs + (size_t)tid * vstride, vbase + (int64_t)tid * a.Lmax, (uint32_t)old_cols * 2u, &bars[1]);
    pdl_wait();
    if (tid == 0) bulk_g2s(Ks + (size_t)old_keys * 64, kbase + (int64_t)old_keys * 64, 128u, &bars[0]);
    if (tid < 64) bulk_g2s(Vs + (size_t)tid * vstride + (size_t)old_cols * 2, vbase + (int64_t)tid * a.Lmax + old_cols, 16u, &bars[1]);

    const showo_seq_mask_t msk = a.masks[seq];
    const int qpos = a.pos0;
    bf16* qrow = a.q + (int64_t)seq * a.rows_per_seq * a.ld + h * 64;
    const int sub = lane & 7, kq = lane >> 3;
    float q[8];
    {
        const uint4 u = *reinterpret_cast<const uint4*>(qrow + sub * 8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h2[j]); q[2 * j] = f.x; q[2 * j + 1] = f.y; }
    }
    const float sc = a.scale * 1.4426950408889634f;
    float mx = kNegBig;
    mbar_wait(&bars[0], 0);
    for (int k0 = warp * 4; k0 < n_pad; k0 += 16) {           // warp-uniform trip count: the shuffles stay convergent
        const int k = k0 + kq;
        float acc = 0.f;
        if (k < a.n_keys) {
            const uint4 u = *reinterpret_cast<const uint4*>(Ks + (size_t)k * 64 + sub * 8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h2[j]);
                acc += q[2 * j] * f.x + q[2 * j + 1] * f.y;
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (k < n_pad) {
            const float s = (k < a.n_keys && omni_allowed(msk, qpos, k)) ? acc * sc : kNegBig;
            if (sub == 0) sc_s[k] = s;
            mx = fmaxf(mx, s);
        }
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int k = tid; k < n_pad; k += 128) {
        const float p = exp2f(sc_s[k] - mx);     // masked / pad keys: exp2(-1e30) = 0
        sc_s[k] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + warp] = sum;
    __syncthreads();
    sum = red[4] + red[5] + red[6] + red[7];
    // o[d] = sum_k p[k] V^T[d][k]; two threads per d split the 8-key chunks
    const int d = tid >> 1, half = tid & 1;
    const uint8_t* vr = Vs + (size_t)d * vstride;
    float acc = 0.f;
    const int n8 = a.n_keys >> 3;
    mbar_wait(&bars[1], 0);
    for (int c = half; c < n8; c += 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(vr + c * 16);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(h2[j]);
            acc += sc_s[c * 8 + 2 * j] * f.x + sc_s[c * 8 + 2 * j + 1] * f.y;
        }
    }
    if (half == 0)                               // ragged tail: the pad columns of the copy are never multiplied
        for (int k = n8 * 8; k < a.n_keys; ++k) acc += sc_s[k] * __bfloat162float(reinterpret_cast<const bf16*>(vr)[k]);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    __syncthreads();   // everyone has read q before it is overwritten
    if (half == 0) qrow[d] = __float2bfloat16(acc / sum);
}

int omni_attention_decode(const AttnArgs& a, cudaStream_t st) {
    if (a.n_seq == 0) return 0;
    SHOWO_CHECK(a.n_keys <= a.Lmax && a.n_keys * 4 <= 48 * 1024, "attention decode: n_keys too large");
    dim3 grid(a.H, a.n_seq);
    static int variant = -1;                     // SHOWO_DECODE_ATTN=1 selects the per-thread-load kernel (A/B switch)
    if (variant < 0) { const char* e = getenv("SHOWO_DECODE_ATTN"); variant = e ? atoi(e) : 2; }
    const int n_pad = (a.n_keys + 7) & ~7;
    const int vstride = n_pad * 2 + 16;
    const size_t smem = (size_t)n_pad * 128 + (size_t)64 * vstride + (size_t)n_pad * 4;
    if (variant == 2 && a.Lmax % 8 == 0 && n_pad <= a.Lmax && smem <= 200 * 1024 && a.ld % 8 == 0 && a.pos0 == a.n_keys - 1 &&
        a.rows_per_seq == 1) {
        static PerDeviceOnce once;
        if (once.need()) SHOWO_CUDA_OK(cudaFuncSetAttribute(omni_attention_decode_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SHOWO_CUDA_OK(launch_kernel(omni_attention_decode_bulk_kernel, grid, dim3(128), smem, st, 1, a, n_pad, vstride));
        note_launch();
        return 0;
    }
    omni_attention_decode_kernel<<<grid, 128, a.n_keys * sizeof(float), st>>>(a);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace showo
