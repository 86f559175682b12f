// Omni-mask attention on tcgen05 / TMEM / TMA (flash-style, online softmax), head_dim 64, any sequence length.
//
// One CTA = 128 query rows of one (sequence, head); it walks the keys in blocks of 128 (blocks that the mask predicate
// rules out for the CTA's rows are skipped by every role).  Two CTAs fit per SM (112 KB smem, 256 TMEM columns each), so
// the TMA loads / MMAs / softmax of neighbouring tiles overlap.
//   warp 0 (1 lane) : TMA producer -- Q tile once, then K block [128 keys x 64] and V^T block (2 chunks [64 dims x 64 keys])
//                     per step through 2-stage rings (128B-swizzled smem, mbarrier full/empty).
//   warp 1 (1 lane) : tcgen05.mma issuer -- S = Q K^T (M=128, N<=128, K=64) into TMEM columns [0,128), then, once the
//                     softmax warps have published P, O += P V (M=128, N=64, K=128) into TMEM columns [128,192).
//                     Warp 1 also owns the TMEM allocation.
//   warps 2..5      : softmax, one query row per thread (TMEM lane = row): pass 1 row max of the block (mask predicate of
//                     showo_seq_mask_t in registers, only for blocks classified "mixed" for the warp's 32 rows), online
//                     rescale of the running sum and of O (tcgen05.ld / tcgen05.st), pass 2 p = exp2((s-m)*c) written as
//                     bf16 into the 128B-swizzled A-operand layout, fence.proxy.async, mbarrier arrive.  Final: O / l.
// (phi.py:696-722 SDPA with the additive mask == softmax over the allowed keys only.)
//
// STATUS (round 1, measured on B200 with bench.py, SHOWO_ATTN_TC=0/1 A/B in the same run): parity-green on every attention
// test (L = 64 .. 1155, all mask kinds, step-style with cached prefix), but SLOWER than the mma.sync kernel of
// attention.cu: ~164 us vs 78 us per layer at 16 seqs x 32 heads x 258 x 387 (an earlier single-pass variant with the whole
// 400-column score row resident in TMEM measured ~130 us).  Why: every score has to cross TMEM -> registers
// (tcgen05.ld, ~64 B/clk/SM), and this kernel reads each S block twice (max pass, exp pass) plus the O rescale, on top of
// three mbarrier round trips per 128-key block; with head_dim 64 the MMAs are tiny, so the tensor pipe never was the
// limiter -- instruction issue / TMEM read bandwidth is (floors: ~20 us for one TMEM read of all scores, ~20 us for the
// exps on MUFU).  mma.sync keeps S in registers for free.  Round-2 plan: FA4 structure -- read S ONCE (two threads per
// row holding 64 scores each), lazy O rescale, softmax of block j overlapping the QK^T of block j+1 through a second S
// buffer, polynomial exp2 on the FMA pipe for part of the elements.  Until then this kernel is opt-in (SHOWO_ATTN_TC=1).
#include "common.cuh"
#include "kernels.h"

namespace showo {

constexpr int kTcThreads = 192;
constexpr float kTcNeg = -1.0e30f;

__device__ __forceinline__ bool tc_allowed(const showo_seq_mask_t& m, int q, int k) {
    const bool ok = (k <= q) | ((q >= m.full_begin) & (q < m.full_end)) | ((k >= m.win_begin) & (k < m.win_end));
    return ok & !((k < m.pad_end) & (q >= m.pad_end));
}
__device__ __forceinline__ bool tc_all_allowed(const showo_seq_mask_t& m, int q_lo, int q_hi, int k_lo, int k_hi, int n_keys) {
    if (k_hi > n_keys) return false;
    if (k_lo < m.pad_end && q_hi >= m.pad_end) return false;
    return ((k_hi - 1) <= q_lo) || ((q_lo >= m.full_begin) && (q_hi < m.full_end)) || ((k_lo >= m.win_begin) && (k_hi <= m.win_end));
}
__device__ __forceinline__ bool tc_none_allowed(const showo_seq_mask_t& m, int q_lo, int q_hi, int k_lo, int k_hi, int n_keys) {
    if (k_lo >= n_keys) return true;
    if (k_hi <= m.pad_end && q_lo >= m.pad_end) return true;
    const bool causal = k_lo <= q_hi;
    const bool full = (q_hi >= m.full_begin) && (q_lo < m.full_end);
    const bool win = (k_lo < m.win_end) && (k_hi > m.win_begin);
    return !(causal || full || win);
}
__device__ __forceinline__ float tc_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kTcThreads, 2)
omni_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                       // 16 KB      [128 rows][128 B]
    uint8_t* sK = sQ + 16384;                 // 2 x 16 KB  [128 keys][128 B]
    uint8_t* sV = sK + 32768;                 // 2 x 16 KB  2 chunks of [64 dims][128 B (64 keys)]
    uint8_t* sP = sV + 32768;                 // 32 KB      2 sub-tiles of [128 rows][128 B (64 keys)]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;              // [2]
    uint64_t* k_empty = bars + 3;             // [2]
    uint64_t* v_full = bars + 5;              // [2]
    uint64_t* v_empty = bars + 7;             // [2]
    uint64_t* s_full = bars + 9;
    uint64_t* p_full = bars + 10;
    uint64_t* pv_done = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_keys = a.n_keys;
    const int nkb = (n_keys + 127) >> 7;
    const showo_seq_mask_t msk = a.masks[seq];
    const int cta_q_lo = a.pos0 + q0, cta_q_hi = a.pos0 + min(q0 + 127, a.rows_per_seq - 1);
    // every role walks the same list of key blocks: those that can hold an allowed key for some row of this CTA
    auto block_live = [&](int j) { return !tc_none_allowed(msk, cta_q_lo, cta_q_hi, j * 128, j * 128 + 128, n_keys); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {          // ================================================================= TMA producer
            const int kv_row0 = seq * a.H + h;
            mbar_arrive_expect_tx(q_full, 16384);
            tma_load_2d(sQ, &tmap_q, q_full, h * 64, seq * a.rows_per_seq + q0);
            int it = 0;
            for (int j = 0; j < nkb; ++j) {
                if (!block_live(j)) continue;
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], 16384);
                tma_load_2d(sK + st * 16384, &tmap_k, &k_full[st], 0, kv_row0 * a.Lmax + j * 128);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], 16384);
                tma_load_2d(sV + st * 16384, &tmap_v, &v_full[st], j * 128, kv_row0 * 64);
                tma_load_2d(sV + st * 16384 + 8192, &tmap_v, &v_full[st], j * 128 + 64, kv_row0 * 64);
                ++it;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {          // ================================================================= MMA issuer
            mbar_wait(q_full, 0);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64);
            int it = 0;
            for (int j = 0; j < nkb; ++j) {
                if (!block_live(j)) continue;
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                // S = Q K_j^T  (the previous block's softmax has finished reading S before it published P, and this thread
                // waited for that P before getting here)
                mbar_wait(&k_full[st], ph);
                tc_fence_after();
                const int nj = min(128, ((n_keys - j * 128) + 15) & ~15);
                const uint32_t idesc_s = umma_idesc_bf16(128, nj);
                const uint32_t k_addr = smem_u32(sK + st * 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_desc_k128(q_addr + k * 32), umma_desc_k128(k_addr + k * 32), idesc_s, k != 0);
                umma_commit(s_full);
                umma_commit(&k_empty[st]);
                // O += P_j V_j
                mbar_wait(&v_full[st], ph);
                mbar_wait(p_full, it & 1);
                tc_fence_after();
                const uint32_t p_addr = smem_u32(sP), v_addr = smem_u32(sV + st * 16384);
                const int nsub = (j * 128 + 64 < n_keys) ? 2 : 1;          // second 64-key half entirely past n_keys: skip it
                for (int hh = 0; hh < nsub; ++hh) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + 128, umma_desc_k128(p_addr + hh * 16384 + k * 32),
                                  umma_desc_k128(v_addr + hh * 8192 + k * 32), idesc_pv, (it | hh | k) != 0);
                }
                umma_commit(pv_done);
                umma_commit(&v_empty[st]);
                ++it;
            }
        }
    } else {                      // ================================================================= softmax + epilogue
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int r = q0 + row;
        const bool row_ok = r < a.rows_per_seq;
        const int qpos = a.pos0 + r;
        const int wq_lo = a.pos0 + q0 + quarter * 32;
        const int wq_hi = a.pos0 + min(q0 + quarter * 32 + 31, a.rows_per_seq - 1);
        const float sc = a.scale * 1.4426950408889634f;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint8_t* prow = sP + row * 128;
        float m = kTcNeg, l = 0.f;
        int it = 0;
        for (int j = 0; j < nkb; ++j) {
            if (!block_live(j)) continue;
            const int k0 = j * 128;
            const bool none = tc_none_allowed(msk, wq_lo, wq_hi, k0, k0 + 128, n_keys);      // warp-uniform
            const bool all_ok = !none && tc_all_allowed(msk, wq_lo, wq_hi, k0, k0 + 128, n_keys);
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            // ---- pass 1: block maximum of this row
            float bm = kTcNeg;
            if (!none) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (k0 + cc * 32 >= n_keys) break;
                    uint32_t v[32];
                    tmem_ld32(t_row + cc * 32, v);
                    tmem_ld_wait();
                    if (all_ok) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) bm = fmaxf(bm, __uint_as_float(v[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int col = k0 + cc * 32 + i;
                            const bool ok = (col < n_keys) && tc_allowed(msk, qpos, col);
                            bm = fmaxf(bm, ok ? __uint_as_float(v[i]) : kTcNeg);
                        }
                    }
                }
            }
            const float m_new = fmaxf(m, bm);
            const float alpha = tc_ex2((m - m_new) * sc);          // m == m_new -> 1 ; m == -1e30 -> 0
            m = m_new;
            const float ms = (m == kTcNeg) ? 0.f : -m * sc;
            l *= alpha;
            // ---- the previous block's PV must have retired before O is rescaled and before P is overwritten
            if (it > 0) {
                mbar_wait(pv_done, (it - 1) & 1);
                tc_fence_after();
                if (!__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t v[32];
                        tmem_ld32(t_row + 128 + cc * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st32(t_row + 128 + cc * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            // ---- pass 2: p = exp2(s*c - m*c) -> bf16 into the swizzled A-operand layout (2 sub-tiles of 64 keys)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t pk[16];
                if (none || k0 + cc * 32 >= n_keys) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;
                } else {
                    uint32_t v[32];
                    tmem_ld32(t_row + cc * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                        if (!all_ok) {
                            const int col = k0 + cc * 32 + i;
                            s0 = ((col < n_keys) && tc_allowed(msk, qpos, col)) ? s0 : kTcNeg;
                            s1 = ((col + 1 < n_keys) && tc_allowed(msk, qpos, col + 1)) ? s1 : kTcNeg;
                        }
                        const float p0 = tc_ex2(fmaf(s0, sc, ms)), p1 = tc_ex2(fmaf(s1, sc, ms));
                        l += p0 + p1;
                        pk[i >> 1] = pack_bf16(p0, p1);
                    }
                }
                uint8_t* pb = prow + (cc >> 1) * 16384;               // sub-tile = 64-key half of the block
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int jchunk = ((cc & 1) * 4 + i) ^ (row & 7);    // 16-byte chunk index, XOR-swizzled with the row
                    *reinterpret_cast<uint4*>(pb + jchunk * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
            tc_fence_before();                 // order my tcgen05.ld/st before the MMA thread's next instructions
            fence_async_smem();                // generic-proxy smem stores -> visible to the tensor core
            mbar_arrive(p_full);
            ++it;
        }
        // ---- epilogue: O / l -> bf16 over the q rows
        if (it > 0) {
            mbar_wait(pv_done, (it - 1) & 1);
            tc_fence_after();
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        bf16* orow = a.q + ((int64_t)seq * a.rows_per_seq + r) * a.ld + h * 64;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            uint32_t v[32];
            if (it > 0) { tmem_ld32(t_row + 128 + cc * 32, v); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
            }
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 pk;
                    pk.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                    pk.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                    pk.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                    pk.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + cc * 32 + i) = pk;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

bool attention_tc_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_ATTN_TC"); v = (e && atoi(e) == 1) ? 1 : 0; }     // opt-in, see STATUS above
    return v == 1;
}
bool attention_tc_supported(const AttnArgs& a) {
    return attention_tc_enabled() && a.rows_per_seq >= 32 && (a.ld % 8) == 0 && a.Lmax % 64 == 0 && a.n_keys >= 1;
}

int omni_attention_tc(const AttnArgs& a, cudaStream_t st) {
    constexpr int kSmem = 16384 + 32768 + 32768 + 32768 + 256;
    static bool attr = false;
    if (!attr) {
        SHOWO_CUDA_OK(cudaFuncSetAttribute(omni_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr = true;
    }
    CUtensorMap mq, mk, mv;
    const int D = a.H * 64;
    const uint64_t q_rows = (uint64_t)a.n_seq * a.rows_per_seq;
    SHOWO_TRY(make_tmap_2d(&mq, a.q, (uint64_t)D, q_rows, (uint64_t)a.ld * 2, 64, 128));
    SHOWO_TRY(make_tmap_2d(&mk, a.kcache, 64, (uint64_t)a.n_seq * a.H * a.Lmax, 128, 64, 128));
    SHOWO_TRY(make_tmap_2d(&mv, a.vtcache, (uint64_t)a.Lmax, (uint64_t)a.n_seq * a.H * 64, (uint64_t)a.Lmax * 2, 64, 64));
    dim3 grid(cdiv(a.rows_per_seq, 128), a.H, a.n_seq);
    SHOWO_CUDA_OK(launch_kernel(omni_attention_tc_kernel, grid, dim3(kTcThreads), kSmem, st, 1, mq, mk, mv, a));
    note_launch();
    return 0;
}

}  // namespace showo
