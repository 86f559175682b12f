// Omni-mask attention on tcgen05 / TMEM / TMA (flash-style, online softmax with LAZY rescaling), head_dim 64, any length.
// (phi.py:696-722 SDPA with the additive mask == softmax over the allowed keys only.)
//
// One CTA = one FULL tile of 128 query rows of one (sequence, head); it walks the keys in blocks of 64 (blocks the mask
// predicate rules out for the CTA's rows are skipped by every role).  Two CTAs per SM (96 KB smem, 256 TMEM columns each).
//   warp 0 (1 lane) : TMA producer -- Q tile once, then K block [64 keys x 64 dims] and V^T block [64 dims x 64 keys] per step
//                     through 3-stage rings (128B-swizzled smem, mbarrier full / empty).
//   warp 1 (1 lane) : tcgen05.mma issuer; owns TMEM: S buffers at columns [0,64) and [64,128) (DOUBLE-buffered), O at [128,192).
//                     Issue order QK(0), then per block i: QK(i+1), PV(i) -- the score MMA of the next block runs while the
//                     softmax warps are busy with the current one.
//   warps 2..5      : softmax, one query row per thread (TMEM lane = row).  Each S block is read from TMEM exactly ONCE
//                     (64 columns -> 64 registers): row maximum (mask predicate of showo_seq_mask_t in registers, only for
//                     blocks classified "mixed" for the warp's 32 rows), p = exp2(s*c - m*c) as 1 FFMA + 1 MUFU, running sum,
//                     bf16 P written into the 128B-swizzled A-operand layout of a double-buffered smem tile.
//                     LAZY rescale (FlashAttention-4): the running maximum m is only raised -- and O / l rescaled through
//                     tcgen05.ld / tcgen05.st -- when a block's maximum exceeds it by more than 2^8 in the exp2 domain; otherwise
//                     the stale m is kept (p <= 256, harmless in fp32 / bf16), so in the common case the softmax warps never wait
//                     for the PV MMA.  Final: O / l -> bf16 rows.
// Rows past the last full tile (rows_per_seq % 128, e.g. the 2 rows of 258 = 2 * 128 + 2 in the t2i step) are handled by the
// mma.sync kernel of attention.cu (omni_attention() launches both).
//
// History: the round-1 version of this file (one S buffer of 128 keys read twice, O rescaled every block, 128-row tiles
// also for the ragged tail -> 3 tiles for 258 rows) measured 164 us per layer at 16 x 32 x 258 x 387 against 78 us for mma.sync.
#include "common.cuh"
#include "kernels.h"

namespace showo {

constexpr int kTcThreads = 192;
constexpr int kTcStages = 3;
constexpr float kTcNeg = -1.0e30f;
constexpr float kTcLazy = 8.0f;          // exp2-domain slack before the running maximum is raised

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
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                              // 16 KB     [128 rows][128 B]
    uint8_t* sK = sQ + 16384;                        // 3 x 8 KB  [64 keys][128 B]
    uint8_t* sV = sK + kTcStages * 8192;             // 3 x 8 KB  [64 dims][128 B (64 keys)]
    uint8_t* sP = sV + kTcStages * 8192;             // 2 x 16 KB [128 rows][128 B (64 keys)]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                     // [3]
    uint64_t* k_empty = bars + 4;                    // [3]
    uint64_t* v_full = bars + 7;                     // [3]
    uint64_t* v_empty = bars + 10;                   // [3]
    uint64_t* s_full = bars + 13;                    // [2]
    uint64_t* p_full = bars + 15;                    // [2]
    uint64_t* pv_done = bars + 17;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_keys = a.n_keys;
    const int nkb = (n_keys + 63) >> 6;
    const showo_seq_mask_t msk = a.masks[seq];
    const int cta_q_lo = a.pos0 + q0, cta_q_hi = a.pos0 + q0 + 127;
    // every role walks the same list of key blocks: those that can hold an allowed key for some row of this CTA
    auto block_live = [&](int j) { return !tc_none_allowed(msk, cta_q_lo, cta_q_hi, j * 64, j * 64 + 64, n_keys); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1); }
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
                const int st = it % kTcStages;
                const uint32_t ph = (it / kTcStages) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], 8192);
                tma_load_2d(sK + st * 8192, &tmap_k, &k_full[st], 0, kv_row0 * a.Lmax + j * 64);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], 8192);
                tma_load_2d(sV + st * 8192, &tmap_v, &v_full[st], j * 64, kv_row0 * 64);
                ++it;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {          // ================================================================= MMA issuer
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64);
            auto issue_qk = [&](int i, int j) {        // S[i & 1] = Q K_j^T ; the softmax of block i - 2 (same buffer) has
                const int st = i % kTcStages;           // published its P, and this thread has waited for that P already
                mbar_wait(&k_full[st], (i / kTcStages) & 1);
                tc_fence_after();
                const int nj = min(64, ((n_keys - j * 64) + 15) & ~15);
                const uint32_t idesc_s = umma_idesc_bf16(128, nj);
                const uint32_t k_addr = smem_u32(sK + st * 8192);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + (i & 1) * 64, umma_desc_k128(q_addr + k * 32), umma_desc_k128(k_addr + k * 32), idesc_s, k != 0);
                umma_commit(&s_full[i & 1]);
                umma_commit(&k_empty[st]);
            };
            mbar_wait(q_full, 0);
            // the list of live blocks is short (<= 32 at L = 2048): walk it with two cursors (QK runs one block ahead of PV)
            int jq = 0;
            while (jq < nkb && !block_live(jq)) ++jq;
            int i = 0;
            if (jq < nkb) issue_qk(0, jq);
            int jp = jq;                                // block of PV(i)
            while (jp < nkb) {
                int jn = jp + 1;
                while (jn < nkb && !block_live(jn)) ++jn;
                if (jn < nkb) issue_qk(i + 1, jn);
                const int st = i % kTcStages;
                mbar_wait(&v_full[st], (i / kTcStages) & 1);
                mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                tc_fence_after();
                const uint32_t p_addr = smem_u32(sP + (i & 1) * 16384), v_addr = smem_u32(sV + st * 8192);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + 128, umma_desc_k128(p_addr + k * 32), umma_desc_k128(v_addr + k * 32), idesc_pv, (i | k) != 0);
                umma_commit(&pv_done[i & 1]);
                umma_commit(&v_empty[st]);
                ++i;
                jp = jn;
            }
        }
    } else {                      // ================================================================= softmax + epilogue
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int r = q0 + row;
        const int qpos = a.pos0 + r;
        const int wq_lo = a.pos0 + q0 + quarter * 32, wq_hi = wq_lo + 31;
        const float sc = a.scale * 1.4426950408889634f;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float m = kTcNeg, l = 0.f;                     // m: the maximum the stored exponentials refer to (raw score units)
        int i = 0;
        for (int j = 0; j < nkb; ++j) {
            if (!block_live(j)) continue;
            const int k0 = j * 64;
            const bool none = tc_none_allowed(msk, wq_lo, wq_hi, k0, k0 + 64, n_keys);      // warp-uniform
            const bool all_ok = !none && tc_all_allowed(msk, wq_lo, wq_hi, k0, k0 + 64, n_keys);
            mbar_wait(&s_full[i & 1], (i >> 1) & 1);
            tc_fence_after();
            float s[64];
            if (!none) {
                uint32_t v0[32], v1[32];
                tmem_ld32(t_row + (i & 1) * 64, v0);
                if (k0 + 32 < n_keys) tmem_ld32(t_row + (i & 1) * 64 + 32, v1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) { s[c] = __uint_as_float(v0[c]); s[32 + c] = (k0 + 32 < n_keys) ? __uint_as_float(v1[c]) : kTcNeg; }
                if (!all_ok) {
#pragma unroll
                    for (int c = 0; c < 64; ++c) {
                        const int col = k0 + c;
                        s[c] = ((col < n_keys) && tc_allowed(msk, qpos, col)) ? s[c] : kTcNeg;
                    }
                }
            }
            float bm = kTcNeg;
            if (!none) {
#pragma unroll
                for (int c = 0; c < 64; ++c) bm = fmaxf(bm, s[c]);
            }
            // lazy maximum: raise m only when the block exceeds it by more than the slack (always on the first real block)
            const bool raise = (bm - m) * sc > kTcLazy;
            if (__any_sync(0xffffffffu, raise)) {
                const float m_new = raise ? bm : m;
                const float alpha = raise ? tc_ex2((m - m_new) * sc) : 1.f;      // m = -1e30 -> 0
                m = m_new;
                l *= alpha;
                if (i > 0) {      // O holds the previous blocks: wait for PV(i-1), then scale this thread's row in TMEM
                    mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t v[32];
                        tmem_ld32(t_row + 128 + cc * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) * alpha);
                        tmem_st32(t_row + 128 + cc * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            const float ms = (m == kTcNeg) ? 0.f : -m * sc;
            uint32_t pk[32];
            if (none) {
#pragma unroll
                for (int c = 0; c < 32; ++c) pk[c] = 0u;
            } else {
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    const float p0 = tc_ex2(fmaf(s[c], sc, ms)), p1 = tc_ex2(fmaf(s[c + 1], sc, ms));
                    l += p0 + p1;
                    pk[c >> 1] = pack_bf16(p0, p1);
                }
            }
            // the P buffer was last read by PV(i-2)
            if (i >= 2) mbar_wait(&pv_done[i & 1], ((i - 2) >> 1) & 1);
            uint8_t* prow = sP + (i & 1) * 16384 + row * 128;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) * 16)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
            tc_fence_before();                 // order my tcgen05.ld/st before the MMA thread's next instructions
            fence_async_smem();                // generic-proxy smem stores -> visible to the tensor core
            mbar_arrive(&p_full[i & 1]);
            ++i;
        }
        // ---- epilogue: O / l -> bf16
        if (i > 0) {
            mbar_wait(&pv_done[(i - 1) & 1], ((i - 1) >> 1) & 1);
            tc_fence_after();
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        bf16* orow = a.out ? a.out + ((int64_t)seq * a.rows_per_seq + r) * a.out_ld + h * 64
                           : a.q + ((int64_t)seq * a.rows_per_seq + r) * a.ld + h * 64;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            uint32_t v[32];
            if (i > 0) { tmem_ld32(t_row + 128 + cc * 32, v); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0u;
            }
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                uint4 o;
                o.x = pack_bf16(__uint_as_float(v[c]) * inv, __uint_as_float(v[c + 1]) * inv);
                o.y = pack_bf16(__uint_as_float(v[c + 2]) * inv, __uint_as_float(v[c + 3]) * inv);
                o.z = pack_bf16(__uint_as_float(v[c + 4]) * inv, __uint_as_float(v[c + 5]) * inv);
                o.w = pack_bf16(__uint_as_float(v[c + 6]) * inv, __uint_as_float(v[c + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + cc * 32 + c) = o;
            }
        }
        if (a.lse != nullptr) a.lse[((int64_t)seq * a.rows_per_seq + r) * a.H + h] = l > 0.f ? m * sc + log2f(l) : 1.0e30f;
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
    if (v < 0) { const char* e = getenv("SHOWO_ATTN_TC"); v = (e && atoi(e) == 0) ? 0 : 1; }     // default on; SHOWO_ATTN_TC=0: mma.sync only
    return v == 1;
}
// number of leading rows of every sequence the tcgen05 kernel takes (full 128-row tiles); the rest goes to the mma.sync kernel
int attention_tc_rows(const AttnArgs& a) {
    if (!attention_tc_enabled() || (a.ld % 8) != 0 || a.Lmax % 64 != 0 || a.n_keys < 1) return 0;
    if (a.out != nullptr && (a.out_ld % 8) != 0) return 0;
    return (a.rows_per_seq / 128) * 128;
}

int omni_attention_tc(const AttnArgs& a, cudaStream_t st) {
    constexpr int kSmem = 16384 + 2 * kTcStages * 8192 + 32768 + 256 + 1024;
    static PerDeviceOnce once;
    if (once.need()) SHOWO_CUDA_OK(cudaFuncSetAttribute(omni_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CUtensorMap mq, mk, mv;
    const int D = a.H * 64;
    const uint64_t q_rows = (uint64_t)a.n_seq * a.rows_per_seq;
    SHOWO_TRY(make_tmap_2d(&mq, a.q, (uint64_t)D, q_rows, (uint64_t)a.ld * 2, 64, 128));
    SHOWO_TRY(make_tmap_2d(&mk, a.kcache, 64, (uint64_t)a.n_seq * a.H * a.Lmax, 128, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mv, a.vtcache, (uint64_t)a.Lmax, (uint64_t)a.n_seq * a.H * 64, (uint64_t)a.Lmax * 2, 64, 64));
    dim3 grid(a.rows_per_seq / 128, a.H, a.n_seq);
    SHOWO_CUDA_OK(launch_kernel(omni_attention_tc_kernel, grid, dim3(kTcThreads), kSmem, st, 1, mq, mk, mv, a));
    note_launch();
    return 0;
}

}  // namespace showo
