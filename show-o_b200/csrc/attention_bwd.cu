// Backward of the omni-mask attention (training step; phi.py:696-722 differentiated), head_dim 64.
//
// Flash-style recomputation from what the forward saved: with sc = scale * log2(e) and lse the exp2-domain log-sum-exp,
//   P = exp2(S * sc - lse) on the pairs the mask predicate allows (0 elsewhere),   delta = rowsum(dO o O),
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),   dQ = scale * dS K,   dK = scale * dS^T Q.
// Two passes, each recomputing S with mma.sync m16n8k16 (bf16 in, fp32 accumulate), no atomics, deterministic:
//   attn_bwd_dkdv_kernel : one CTA = 64 keys of one (sequence, head); K / V fragments live in registers, Q / dO tiles of 64
//                          queries stream through a double-buffered cp.async pipeline; the CTA works on the TRANSPOSED
//                          problem S^T = K Q^T so that P^T / dS^T come out of the accumulators in A-fragment layout.
//   attn_bwd_dq_kernel   : one CTA = 64 queries; Q / dO fragments in registers, K / V tiles stream.
// Tiles the mask rules out for the CTA are skipped; the per-element predicate only runs on tiles that are not fully allowed.
// The same tile is read through ldmatrix (as [n][k]) and ldmatrix.trans (as [k][n]), so nothing is transposed in memory.
#include "attn_common.cuh"

namespace showo {

// delta[r, h] = sum_d dO[r, 64h + d] * O[r, 64h + d]
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(AttnBwdArgs a, int n_rows) {
    const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    for (int h = warp; h < a.H; h += 8) {
        const float2 x = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(a.o + (int64_t)r * a.o_ld + h * 64)[lane]);
        const float2 y = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(a.d_o + (int64_t)r * a.do_ld + h * 64)[lane]);
        const float s = warp_sum(x.x * y.x + x.y * y.y);
        if (lane == 0) a.delta[(int64_t)r * a.H + h] = s;
    }
}

// A-operand fragment (16 rows x 64 dims) of rows r0 / r0 + 8 straight from global (zeros for rows past the sequence)
__device__ __forceinline__ void load_frag_rows(uint32_t (&f)[4][4], const bf16* row0, const bf16* row1, bool ok0, bool ok1, int t4) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int c = kk * 16 + t4 * 2;
        f[kk][0] = ok0 ? *reinterpret_cast<const uint32_t*>(row0 + c) : 0u;
        f[kk][1] = ok1 ? *reinterpret_cast<const uint32_t*>(row1 + c) : 0u;
        f[kk][2] = ok0 ? *reinterpret_cast<const uint32_t*>(row0 + c + 8) : 0u;
        f[kk][3] = ok1 ? *reinterpret_cast<const uint32_t*>(row1 + c + 8) : 0u;
    }
}
// accumulator tile (16 x 64, fp32) -> A-operand fragments of the next MMA (bf16)
__device__ __forceinline__ void acc_to_afrag(uint32_t (&pf)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        pf[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pf[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pf[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pf[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
}
// acc[16 x 64] += A[16 x 64] * T^T  with the tile T stored [64 rows n][64 cols k] (k contiguous): ldmatrix, no transpose
__device__ __forceinline__ void mma_tile_nk(float (&acc)[8][4], const uint32_t (&af)[4][4], const bf16 (*T)[kPad], int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nb = 0; nb < 8; nb += 2) {
            uint32_t b[4];
            ldmatrix_x4(b, &T[(nb + (lane >> 4)) * 8 + (lane & 7)][kk * 16 + ((lane >> 3) & 1) * 8]);
            mma_bf16_16816(acc[nb], af[kk], b[0], b[1]);
            mma_bf16_16816(acc[nb + 1], af[kk], b[2], b[3]);
        }
    }
}
// acc[16 x 64] += A[16 x 64] * T  with the tile T stored [64 rows k][64 cols n] (n contiguous): ldmatrix.trans
__device__ __forceinline__ void mma_tile_kn(float (&acc)[8][4], const uint32_t (&af)[4][4], const bf16 (*T)[kPad], int lane) {
    const int mi = lane >> 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int nd = 0; nd < 8; nd += 2) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, &T[kk * 16 + (mi & 1) * 8 + (lane & 7)][(nd + (mi >> 1)) * 8]);
            mma_bf16_16816(acc[nd], af[kk], b[0], b[1]);
            mma_bf16_16816(acc[nd + 1], af[kk], b[2], b[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(128, 3) attn_bwd_dkdv_kernel(AttnBwdArgs a) {
    __shared__ __align__(16) bf16 Qs[2][64][kPad];
    __shared__ __align__(16) bf16 Ds[2][64][kPad];
    __shared__ float Ls[2][64], Es[2][64];

    const int seq = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const showo_seq_mask_t msk = a.masks[seq];
    const int L = a.L;
    const int64_t rb = (int64_t)seq * L;

    const int kr0 = k0 + warp * 16 + g, kr1 = kr0 + 8;
    const bool k0_ok = kr0 < L, k1_ok = kr1 < L;
    uint32_t kf[4][4], vf[4][4];
    load_frag_rows(kf, a.k + (rb + kr0) * a.k_ld + h * 64, a.k + (rb + kr1) * a.k_ld + h * 64, k0_ok, k1_ok, t4);
    load_frag_rows(vf, a.v + (rb + kr0) * a.v_ld + h * 64, a.v + (rb + kr1) * a.v_ld + h * 64, k0_ok, k1_ok, t4);
    const bool warp_active = (k0 + warp * 16) < L;
    const int wk_lo = k0 + warp * 16;
    const int cta_k_hi = min(k0 + 64, L);
    const int n_qt = (L + 63) / 64;

    auto next_tile = [&](int qt) {
        while (qt < n_qt && !omni_tile_possible(msk, qt * 64, min(qt * 64 + 63, L - 1), k0, cta_k_hi)) ++qt;
        return qt;
    };
    auto load_tile = [&](int qt, int buf) {
        const int q0 = qt * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = threadIdx.x + i * 128;
            const int row = idx >> 3, ch = idx & 7;
            const bool ok = q0 + row < L;
            const int64_t r = rb + (ok ? q0 + row : L - 1);
            cp_async16(&Qs[buf][row][ch * 8], a.q + r * a.q_ld + h * 64 + ch * 8, ok ? 16 : 0);
            cp_async16(&Ds[buf][row][ch * 8], a.d_o + r * a.do_ld + h * 64 + ch * 8, ok ? 16 : 0);
        }
        if (threadIdx.x < 64) {
            const bool ok = q0 + (int)threadIdx.x < L;
            Ls[buf][threadIdx.x] = ok ? a.lse[(rb + q0 + threadIdx.x) * a.H + h] : 1.0e30f;     // p = 0 past the sequence
            Es[buf][threadIdx.x] = ok ? a.delta[(rb + q0 + threadIdx.x) * a.H + h] : 0.f;
        }
        cp_async_commit();
    };

    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    const float sc = a.scale * 1.4426950408889634f;

    int qt = next_tile(0);
    int buf = 0;
    if (qt < n_qt) load_tile(qt, 0);
    while (qt < n_qt) {
        const int qt_next = next_tile(qt + 1);
        if (qt_next < n_qt) { load_tile(qt_next, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        if (warp_active) {
            const int q0 = qt * 64;
            // S^T = K Q^T  (rows = this warp's 16 keys, columns = the tile's 64 queries)
            float s[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
            mma_tile_nk(s, kf, Qs[buf], lane);
            const bool all_ok = (q0 + 64 <= L) && omni_tile_all_allowed(msk, q0, q0 + 63, wk_lo, wk_lo + 16, L);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qc = nb * 8 + t4 * 2 + (e & 1);
                    float p = ex2_approx(fmaf(s[nb][e], sc, -Ls[buf][qc]));
                    if (!all_ok) {
                        const int kr = (e < 2) ? kr0 : kr1;
                        const bool ok = (q0 + qc < L) && (kr < L) && omni_allowed(msk, q0 + qc, kr);
                        p = ok ? p : 0.f;
                    }
                    s[nb][e] = p;
                }
            }
            uint32_t pf[4][4];
            acc_to_afrag(pf, s);
            mma_tile_kn(dv, pf, Ds[buf], lane);                    // dV += P^T dO
            float dp[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) { dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f; }
            mma_tile_nk(dp, vf, Ds[buf], lane);                    // dP^T = V dO^T
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s[nb][e] *= dp[nb][e] - Es[buf][nb * 8 + t4 * 2 + (e & 1)];
            }
            acc_to_afrag(pf, s);
            mma_tile_kn(dk, pf, Qs[buf], lane);                    // dK += dS^T Q
        }
        __syncthreads();
        buf ^= 1;
        qt = qt_next;
    }
    bf16* dk0 = a.dk + (rb + kr0) * a.dk_ld + h * 64;
    bf16* dk1 = a.dk + (rb + kr1) * a.dk_ld + h * 64;
    bf16* dv0 = a.dv + (rb + kr0) * a.dv_ld + h * 64;
    bf16* dv1 = a.dv + (rb + kr1) * a.dv_ld + h * 64;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        const int c = nd * 8 + t4 * 2;
        if (k0_ok) {
            *reinterpret_cast<uint32_t*>(dk0 + c) = pack_bf16(dk[nd][0] * a.scale, dk[nd][1] * a.scale);
            *reinterpret_cast<uint32_t*>(dv0 + c) = pack_bf16(dv[nd][0], dv[nd][1]);
        }
        if (k1_ok) {
            *reinterpret_cast<uint32_t*>(dk1 + c) = pack_bf16(dk[nd][2] * a.scale, dk[nd][3] * a.scale);
            *reinterpret_cast<uint32_t*>(dv1 + c) = pack_bf16(dv[nd][2], dv[nd][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ dQ
__global__ void __launch_bounds__(128, 3) attn_bwd_dq_kernel(AttnBwdArgs a) {
    __shared__ __align__(16) bf16 Ks[2][64][kPad];
    __shared__ __align__(16) bf16 Vs[2][64][kPad];

    const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const showo_seq_mask_t msk = a.masks[seq];
    const int L = a.L;
    const int64_t rb = (int64_t)seq * L;

    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const bool r0_ok = r0 < L, r1_ok = r1 < L;
    uint32_t qf[4][4], df[4][4];
    load_frag_rows(qf, a.q + (rb + r0) * a.q_ld + h * 64, a.q + (rb + r1) * a.q_ld + h * 64, r0_ok, r1_ok, t4);
    load_frag_rows(df, a.d_o + (rb + r0) * a.do_ld + h * 64, a.d_o + (rb + r1) * a.do_ld + h * 64, r0_ok, r1_ok, t4);
    const float lse0 = r0_ok ? a.lse[(rb + r0) * a.H + h] : 1.0e30f, lse1 = r1_ok ? a.lse[(rb + r1) * a.H + h] : 1.0e30f;
    const float del0 = r0_ok ? a.delta[(rb + r0) * a.H + h] : 0.f, del1 = r1_ok ? a.delta[(rb + r1) * a.H + h] : 0.f;
    const bool warp_active = (q0 + warp * 16) < L;
    const int wq_lo = q0 + warp * 16, wq_hi = min(q0 + warp * 16 + 15, L - 1);
    const int cta_q_hi = min(q0 + 63, L - 1);
    const int n_kt = (L + 63) / 64;

    auto next_tile = [&](int kt) {
        while (kt < n_kt && !omni_tile_possible(msk, q0, cta_q_hi, kt * 64, min(kt * 64 + 64, L))) ++kt;
        return kt;
    };
    auto load_tile = [&](int kt, int buf) {
        const int k0 = kt * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = threadIdx.x + i * 128;
            const int row = idx >> 3, ch = idx & 7;
            const bool ok = k0 + row < L;
            const int64_t r = rb + (ok ? k0 + row : L - 1);
            cp_async16(&Ks[buf][row][ch * 8], a.k + r * a.k_ld + h * 64 + ch * 8, ok ? 16 : 0);
            cp_async16(&Vs[buf][row][ch * 8], a.v + r * a.v_ld + h * 64 + ch * 8, ok ? 16 : 0);
        }
        cp_async_commit();
    };

    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const float sc = a.scale * 1.4426950408889634f;

    int kt = next_tile(0);
    int buf = 0;
    if (kt < n_kt) load_tile(kt, 0);
    while (kt < n_kt) {
        const int kt_next = next_tile(kt + 1);
        if (kt_next < n_kt) { load_tile(kt_next, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        if (warp_active) {
            const int k0 = kt * 64;
            float s[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
            mma_tile_nk(s, qf, Ks[buf], lane);                     // S = Q K^T
            const bool all_ok = omni_tile_all_allowed(msk, wq_lo, wq_hi, k0, k0 + 64, L);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float p = ex2_approx(fmaf(s[nb][e], sc, (e < 2) ? -lse0 : -lse1));
                    if (!all_ok) {
                        const int kc = k0 + nb * 8 + t4 * 2 + (e & 1);
                        const bool ok = (kc < L) && omni_allowed(msk, (e < 2) ? r0 : r1, kc);
                        p = ok ? p : 0.f;
                    }
                    s[nb][e] = p;
                }
            }
            float dp[8][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) { dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f; }
            mma_tile_nk(dp, df, Vs[buf], lane);                    // dP = dO V^T
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                s[nb][0] *= dp[nb][0] - del0; s[nb][1] *= dp[nb][1] - del0;
                s[nb][2] *= dp[nb][2] - del1; s[nb][3] *= dp[nb][3] - del1;
            }
            uint32_t pf[4][4];
            acc_to_afrag(pf, s);
            mma_tile_kn(dq, pf, Ks[buf], lane);                    // dQ += dS K
        }
        __syncthreads();
        buf ^= 1;
        kt = kt_next;
    }
    bf16* o0 = a.dq + (rb + r0) * a.dq_ld + h * 64;
    bf16* o1 = a.dq + (rb + r1) * a.dq_ld + h * 64;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        const int c = nd * 8 + t4 * 2;
        if (r0_ok) *reinterpret_cast<uint32_t*>(o0 + c) = pack_bf16(dq[nd][0] * a.scale, dq[nd][1] * a.scale);
        if (r1_ok) *reinterpret_cast<uint32_t*>(o1 + c) = pack_bf16(dq[nd][2] * a.scale, dq[nd][3] * a.scale);
    }
}

int omni_attention_backward(const AttnBwdArgs& a, cudaStream_t st) {
    if (a.n_seq == 0 || a.L == 0) return 0;
    SHOWO_CHECK((a.q_ld | a.k_ld | a.v_ld | a.o_ld | a.do_ld | a.dq_ld | a.dk_ld | a.dv_ld) % 8 == 0,
                "attention backward: row strides must be multiples of 8 elements");
    const int rows = a.n_seq * a.L;
    attn_bwd_delta_kernel<<<rows, 256, 0, st>>>(a, rows);
    dim3 grid(cdiv(a.L, 64), a.H, a.n_seq);
    attn_bwd_dkdv_kernel<<<grid, 128, 0, st>>>(a);
    attn_bwd_dq_kernel<<<grid, 128, 0, st>>>(a);
    note_launch(3);
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace showo
