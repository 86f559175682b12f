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
#include "attn_common.cuh"

namespace showo {

constexpr int kTcThreads = 320;
constexpr int kTcStages = 2;
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

// In-kernel timeline of CTA 0 (SHOWO_TC_PROF=1, read back with showo_debug_tc_prof): clock64 stamps per role and block.
__device__ unsigned long long g_tc_prof[3 * 64 * 6];
#define TC_STAMP(role, g, f) do { if (prof && blockIdx.x == 0 && (g) < 64) g_tc_prof[((role) * 64 + (g)) * 6 + (f)] = clock64(); } while (0)

// work unit u -> (sequence, head, 128-row query tile); tile index fastest
struct TcUnit { int seq, h, q0; };
__device__ __forceinline__ TcUnit tc_unit(int u, int tiles, int H) {
    const int t = u % tiles, sh = u / tiles;
    return TcUnit{sh / H, sh % H, t * 128};
}
// the units of one CTA (u = blockIdx.x, + gridDim.x, ...) decoded incrementally: the stride is split once into its (tile, head,
// sequence) digits, so a step is three adds with carries instead of four integer divisions
struct TcUnitWalk {
    int t, h, seq, dt, dh, ds, tiles, H;
    __device__ __forceinline__ TcUnitWalk(int u0, int stride, int tiles_, int H_) : tiles(tiles_), H(H_) {
        t = u0 % tiles_; const int sh = u0 / tiles_; h = sh % H_; seq = sh / H_;
        dt = stride % tiles_; const int s2 = stride / tiles_; dh = s2 % H_; ds = s2 / H_;
    }
    __device__ __forceinline__ TcUnit get() const { return TcUnit{seq, h, t * 128}; }
    __device__ __forceinline__ void step() {
        t += dt; if (t >= tiles) { t -= tiles; ++h; }
        h += dh; if (h >= H) { h -= H; ++seq; }
        seq += ds;
    }
};
// first key a tile may attend to, rounded down to the 16-byte granule of the V^T box: rows at or behind the left padding never
// see a pad column, so their key loop starts at pad_end instead of 0 (t2i: 60-126 of the 387 keys are pads)
__device__ __forceinline__ int tc_key_begin(const showo_seq_mask_t& m, int q_lo) { return q_lo >= m.pad_end ? (m.pad_end & ~7) : 0; }

// ------------------------------------------------------------------------------------------------ tail phase
// The rows of a sequence that do not fill a 128-row tile (2 of the 258 rows of the t2i denoise step, 3 of 1155 in the training
// step) are handled by the SAME kernel once a CTA has finished its tensor-core units: all 320 threads, CUDA cores, out of the
// shared memory the pipeline no longer needs.  Items = (sequence, head) pairs, handed out through a self-resetting global counter,
// so the CTAs that got one unit fewer in the round-robin (1024 units on 296 CTAs) pick up most of them: the tail work fills the
// units' load-imbalance bubble instead of a separate launch (a stand-alone kernel cost 18 us per layer at L = 387).
// Per item the visible keys are walked in chunks of 128: one bulk copy brings the chunk's K rows (contiguous in the cache), 64 row
// copies its V^T columns; two threads per key for q.K, a warp per row for the online-softmax statistics, (output dim, key slice)
// per thread for P.V; all tail rows of the item share the chunk.
constexpr int kTailMax = 4;                                    // rows per (sequence, head) this phase takes (more: mma.sync tail kernel)
constexpr int kTailChunk = 128;
constexpr int kTailVStride = kTailChunk * 2 + 16;              // bytes per V^T row in smem
constexpr int kTailParts = kTcThreads / 64;                    // key slices of the P.V step
struct TailSmem {
    static constexpr size_t k_off = 0;                                                   // [128][128 B]
    static constexpr size_t v_off = (size_t)kTailChunk * 128;                            // [64][kTailVStride]
    static constexpr size_t buf_bytes = v_off + (size_t)64 * kTailVStride;               // one K + V^T chunk; two of them (double buffer)
    static constexpr size_t sc_off = 2 * buf_bytes;                                      // float [kTailMax][128]
    static constexpr size_t q_off = sc_off + (size_t)kTailMax * kTailChunk * 4;          // float [kTailMax][64]
    static constexpr size_t part_off = q_off + (size_t)kTailMax * 64 * 4;                // float [kTailParts][kTailMax][64]
    static constexpr size_t stat_off = part_off + (size_t)kTailParts * kTailMax * 64 * 4;   // float m[4], l[4], alpha[4]; int item
    static constexpr size_t total = stat_off + 64;
};
static_assert(TailSmem::total <= 32768 + 2 * kTcStages * 8192 + 32768, "the tail phase reuses the pipeline's tile buffers");

__device__ __forceinline__ void tc_tail_phase(const AttnArgs& a, uint8_t* smem, uint64_t* bar, int n_tail, int row_begin, int* work_ctr) {
    float* sc_s = reinterpret_cast<float*>(smem + TailSmem::sc_off);
    float* q_s = reinterpret_cast<float*>(smem + TailSmem::q_off);
    float* part = reinterpret_cast<float*>(smem + TailSmem::part_off);
    float* st_m = reinterpret_cast<float*>(smem + TailSmem::stat_off);
    float* st_l = st_m + 4;
    float* st_alpha = st_m + 8;
    volatile int* s_item = reinterpret_cast<volatile int*>(st_m + 12);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_items = a.n_seq * a.H;
    const float scl = a.scale * 1.4426950408889634f;
    const int d = tid & 63, slice = tid >> 6;
    uint32_t phase[2] = {0u, 0u};
    // a chunk's copies: threads 0..63 bring one V^T row each, thread 64 the K rows; every issuer arrives with its own byte count
    // (barrier count 65), so no CTA barrier is needed between the announcement and the copies
    auto issue_chunk = [&](int buf, const bf16* kbase, const bf16* vbase, int c0, int nk) {
        uint8_t* Kb = smem + (size_t)buf * TailSmem::buf_bytes + TailSmem::k_off;
        uint8_t* Vb = smem + (size_t)buf * TailSmem::buf_bytes + TailSmem::v_off;
        if (tid < 64) {
            mbar_arrive_expect_tx(&bar[buf], (uint32_t)nk * 2u);
            bulk_g2s(Vb + (size_t)tid * kTailVStride, vbase + (int64_t)tid * a.Lmax + c0, (uint32_t)nk * 2u, &bar[buf]);
        } else if (tid == 64) {
            mbar_arrive_expect_tx(&bar[buf], (uint32_t)nk * 128u);
            bulk_g2s(Kb, kbase + (int64_t)c0 * 64, (uint32_t)nk * 128u, &bar[buf]);
        }
    };
    for (;;) {
        __syncthreads();                         // the previous item's smem (and s_item) is no longer read
        if (tid == 0) {
            const int t = atomicAdd(work_ctr, 1);
            if (t == n_items + (int)gridDim.x - 1) *work_ctr = 0;      // the launch's last grab: every CTA has had its failing one
            *s_item = t;
        }
        if (tid < 12) st_m[tid] = tid < 4 ? kNegBig : 0.f;
        __syncthreads();
        const int item = *s_item;
        if (item >= n_items) break;
        const int seq = item / a.H, h = item % a.H;
        const showo_seq_mask_t msk = a.masks[seq];
        int k_begin = a.n_keys, k_end = 0;       // the union of the rows' key ranges
        for (int r = 0; r < n_tail; ++r) {
            const RowKeys rk = omni_row_keys(msk, a.pos0 + row_begin + r, a.n_keys);
            k_begin = min(k_begin, rk.b1 > rk.b0 ? min(rk.a0, rk.b0) : rk.a0);
            k_end = max(k_end, rk.b1 > rk.b0 ? max(rk.a1, rk.b1) : rk.a1);
        }
        const int kb8 = k_begin & ~7;
        const bf16* kbase = a.kcache + ((int64_t)seq * a.H + h) * (int64_t)a.Lmax * 64;
        const bf16* vbase = a.vtcache + ((int64_t)seq * a.H + h) * 64 * (int64_t)a.Lmax;
        bf16* qbase = a.q + ((int64_t)seq * a.rows_per_seq + row_begin) * a.ld + h * 64;
        for (int i = tid; i < n_tail * 64; i += kTcThreads) q_s[i] = __bfloat162float(qbase[(int64_t)(i >> 6) * a.ld + (i & 63)]);
        float acc[kTailMax];
#pragma unroll
        for (int r = 0; r < kTailMax; ++r) acc[r] = 0.f;
        auto chunk_keys = [&](int c0) { return min(kTailChunk, ((k_end - c0) + 7) & ~7); };     // multiple of 8, within Lmax
        if (kb8 < k_end) issue_chunk(0, kbase, vbase, kb8, chunk_keys(kb8));
        int ci = 0;
        for (int c0 = kb8; c0 < k_end; c0 += kTailChunk, ++ci) {
            const int nk = chunk_keys(c0);
            const int buf = ci & 1;
            __syncthreads();                     // chunk ci - 1 (the other buffer) is no longer read; q_s / the statistics are written
            if (c0 + kTailChunk < k_end) issue_chunk(buf ^ 1, kbase, vbase, c0 + kTailChunk, chunk_keys(c0 + kTailChunk));   // next chunk in flight
            const uint8_t* Ks = smem + (size_t)buf * TailSmem::buf_bytes + TailSmem::k_off;
            const uint8_t* Vs = smem + (size_t)buf * TailSmem::buf_bytes + TailSmem::v_off;
            mbar_wait(&bar[buf], phase[buf]);
            phase[buf] ^= 1u;
            // ---- scores: two threads per key (32 dims each); 16-byte piece (c ^ key): consecutive keys hit different bank groups
            if (tid < 2 * kTailChunk) {
                const int kk = tid >> 1, hf = tid & 1;
                const int k = c0 + kk;
                float s[kTailMax];
#pragma unroll
                for (int r = 0; r < kTailMax; ++r) s[r] = 0.f;
                if (kk < nk) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cc = (hf * 4 + c) ^ (kk & 7);
                        const uint4 u = *reinterpret_cast<const uint4*>(Ks + (size_t)kk * 128 + cc * 16);
                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
                        float kf[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h2[j]); kf[2 * j] = f.x; kf[2 * j + 1] = f.y; }
#pragma unroll
                        for (int r = 0; r < kTailMax; ++r) {
                            if (r < n_tail) {
                                const float4 qa = *reinterpret_cast<const float4*>(q_s + r * 64 + cc * 8), qb = *reinterpret_cast<const float4*>(q_s + r * 64 + cc * 8 + 4);
                                s[r] = fmaf(kf[0], qa.x, s[r]); s[r] = fmaf(kf[1], qa.y, s[r]); s[r] = fmaf(kf[2], qa.z, s[r]); s[r] = fmaf(kf[3], qa.w, s[r]);
                                s[r] = fmaf(kf[4], qb.x, s[r]); s[r] = fmaf(kf[5], qb.y, s[r]); s[r] = fmaf(kf[6], qb.z, s[r]); s[r] = fmaf(kf[7], qb.w, s[r]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kTailMax; ++r) {
                    if (r < n_tail) {
                        const float tot = s[r] + __shfl_xor_sync(0xffffffffu, s[r], 1);
                        if (hf == 0 && kk < nk)
                            sc_s[r * kTailChunk + kk] = omni_allowed(msk, a.pos0 + row_begin + r, k) && k < a.n_keys ? tot * scl : kNegBig;
                    }
                }
            }
            __syncthreads();
            // ---- one warp per row: chunk maximum, running maximum / sum, probabilities in place
            if (warp < n_tail) {
                float* sr = sc_s + warp * kTailChunk;
                float mx = kNegBig;
                for (int i = lane; i < nk; i += 32) mx = fmaxf(mx, sr[i]);
                mx = warp_max(mx);
                const float m_old = st_m[warp], m_new = fmaxf(m_old, mx);
                float sum = 0.f;
                for (int i = lane; i < nk; i += 32) {
                    const float p = m_new > kNegBig ? ex2_approx(sr[i] - m_new) : 0.f;
                    sr[i] = p;
                    sum += p;
                }
                sum = warp_sum(sum);
                if (lane == 0) {
                    const float alpha = m_old > kNegBig ? ex2_approx(m_old - m_new) : 0.f;
                    st_alpha[warp] = alpha;
                    st_l[warp] = st_l[warp] * alpha + sum;
                    st_m[warp] = m_new;
                }
            }
            __syncthreads();
            // ---- P.V: thread = (output dim d, slice of the chunk's 8-key pieces)
            {
                const uint8_t* vr = Vs + (size_t)d * kTailVStride;
#pragma unroll
                for (int r = 0; r < kTailMax; ++r) if (r < n_tail) acc[r] *= st_alpha[r];
                const int n8 = nk >> 3;
                for (int c = slice; c < n8; c += kTailParts) {
                    const uint4 u = *reinterpret_cast<const uint4*>(vr + c * 16);
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
                    float vf[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h2[j]); vf[2 * j] = f.x; vf[2 * j + 1] = f.y; }
#pragma unroll
                    for (int r = 0; r < kTailMax; ++r) {
                        if (r < n_tail) {
                            const float4 pa = *reinterpret_cast<const float4*>(sc_s + r * kTailChunk + c * 8), pb = *reinterpret_cast<const float4*>(sc_s + r * kTailChunk + c * 8 + 4);
                            const float pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (pp[j] != 0.f) acc[r] = fmaf(pp[j], vf[j], acc[r]);    // 0 x stale cache bytes stays 0
                        }
                    }
                }
            }
        }
        // ---- merge the key slices, normalise, write
#pragma unroll
        for (int r = 0; r < kTailMax; ++r) if (r < n_tail) part[(slice * kTailMax + r) * 64 + d] = acc[r];
        __syncthreads();
        for (int i = tid; i < n_tail * 64; i += kTcThreads) {
            const int r = i >> 6, dd = i & 63;
            float o = 0.f;
#pragma unroll
            for (int p = 0; p < kTailParts; ++p) o += part[(p * kTailMax + r) * 64 + dd];
            const float l = st_l[r];
            const int row = row_begin + r;
            bf16* orow = a.out ? a.out + ((int64_t)seq * a.rows_per_seq + row) * a.out_ld + h * 64 : qbase + (int64_t)r * a.ld;
            orow[dd] = __float2bfloat16(l > 0.f ? o / l : 0.f);
            if (a.lse != nullptr && dd == 0) a.lse[((int64_t)seq * a.rows_per_seq + row) * a.H + h] = l > 0.f ? st_m[r] + log2f(l) : 1.0e30f;
        }
    }
}

constexpr int kTcMaskSlots = 128;      // mask descriptors cached in shared memory (more sequences than this: read from global)

// live 64-key blocks of a tile as a bit set (block j = keys [k_lo + 64 j, k_lo + 64 j + 64)); at most 32 blocks at L = 2048
__device__ __forceinline__ uint32_t tc_live_blocks(const showo_seq_mask_t& m, int q_lo, int q_hi, int k_lo, int nkb, int n_keys) {
    uint32_t live = 0;
    for (int j = 0; j < nkb; ++j)
        live |= tc_none_allowed(m, q_lo, q_hi, k_lo + j * 64, k_lo + j * 64 + 64, n_keys) ? 0u : (1u << j);
    return live;
}

__global__ void __launch_bounds__(kTcThreads, 2)
omni_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const AttnArgs a, const int n_units, const int tiles, const int prof, const uint32_t sleep_ns,
                         const int n_tail, int* const work_ctr) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                              // 2 x 16 KB [128 rows][128 B], double-buffered across units
    uint8_t* sK = sQ + 32768;                        // 2 x 8 KB  [64 keys][128 B]
    uint8_t* sV = sK + kTcStages * 8192;             // 2 x 8 KB  [64 dims][128 B (64 keys)]
    uint8_t* sP = sV + kTcStages * 8192;             // 2 x 16 KB [128 rows][128 B (64 keys)]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
    uint64_t* q_full = bars;                         // [2]
    uint64_t* q_empty = bars + 2;                    // [2]
    uint64_t* k_full = bars + 4;                     // [2]
    uint64_t* k_empty = bars + 6;                    // [2]
    uint64_t* v_full = bars + 8;                     // [2]
    uint64_t* v_empty = bars + 10;                   // [2]
    uint64_t* s_full = bars + 12;                    // [2]
    uint64_t* p_full = bars + 14;                    // [2]
    uint64_t* pv_done = bars + 16;                   // [2]
    uint64_t* o_free = bars + 18;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    showo_seq_mask_t* s_masks = reinterpret_cast<showo_seq_mask_t*>(bars + 24);      // [kTcMaskSlots]
    float* x_bm = reinterpret_cast<float*>(s_masks + kTcMaskSlots);                  // [2 block parities][2 halves][128 rows] block maxima
    float* x_l = x_bm + 4 * 128;                                                     // [2 unit parities][2 halves][128 rows] partial row sums
    volatile int* x_need = reinterpret_cast<volatile int*>(x_l + 4 * 128);           // [2 block parities][2 halves][4 quarters]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_keys = a.n_keys;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
        for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 256); mbar_init(&pv_done[s], 1); mbar_init(&o_free[s], 256); }
        mbar_init(bars + 22, 65); mbar_init(bars + 23, 65);   // tail phase: the bulk copies of a K / V^T chunk (64 row copies + 1), two buffers
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    // the mask descriptors are host data uploaded before the predecessor kernel started: safe to read ahead of the PDL wait
    for (int i = threadIdx.x; i < min(a.n_seq, kTcMaskSlots); i += kTcThreads) s_masks[i] = a.masks[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();
    auto mask_of = [&](int seq) -> showo_seq_mask_t { return seq < kTcMaskSlots ? s_masks[seq] : a.masks[seq]; };

    // Every role walks the same units (this CTA's share, round robin) and, inside a unit, the same set of key blocks: the 64-key
    // blocks from the tile's first visible key on that can hold an allowed key for some row of the tile (a bit set per unit).
    if (warp == 0) {
        if (lane == 0) {          // ================================================================= TMA producer
            uint32_t kv_it = 0, u_it = 0;
            TcUnitWalk walk(blockIdx.x, gridDim.x, tiles, a.H);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++u_it, walk.step()) {
                const TcUnit un = walk.get();
                const showo_seq_mask_t msk = mask_of(un.seq);
                const int q_lo = a.pos0 + un.q0, q_hi = q_lo + 127;
                const int k_lo = tc_key_begin(msk, q_lo);
                const int nkb = (n_keys - k_lo + 63) >> 6;
                const int kv_row0 = un.seq * a.H + un.h;
                uint32_t rem = tc_live_blocks(msk, q_lo, q_hi, k_lo, nkb, n_keys);
                const int qb = u_it & 1;                          // the unit two back has finished reading this Q buffer
                mbar_wait_sleep(&q_empty[qb], ((u_it >> 1) & 1) ^ 1, sleep_ns);
                mbar_arrive_expect_tx(&q_full[qb], 16384);
                tma_load_2d(sQ + qb * 16384, &tmap_q, &q_full[qb], un.h * 64, un.seq * a.rows_per_seq + un.q0);
                while (rem) {
                    const int j = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const int k0 = k_lo + j * 64;
                    const int st = kv_it % kTcStages;
                    const uint32_t ph = (kv_it / kTcStages) & 1;
                    TC_STAMP(2, kv_it, 0);
                    mbar_wait_sleep(&k_empty[st], ph ^ 1, sleep_ns);
                    TC_STAMP(2, kv_it, 1);
                    mbar_arrive_expect_tx(&k_full[st], 8192);
                    tma_load_2d(sK + st * 8192, &tmap_k, &k_full[st], 0, kv_row0 * a.Lmax + k0);
                    mbar_wait_sleep(&v_empty[st], ph ^ 1, sleep_ns);
                    mbar_arrive_expect_tx(&v_full[st], 8192);
                    tma_load_2d(sV + st * 8192, &tmap_v, &v_full[st], k0, kv_row0 * 64);
                    TC_STAMP(2, kv_it, 2);
                    ++kv_it;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {          // ================================================================= MMA issuer
            // operand descriptors: the 14-bit start-address field advances by 2 per 32-byte K step
            const uint64_t dq0 = umma_desc_k128(smem_u32(sQ));
            const uint64_t dk0 = umma_desc_k128(smem_u32(sK)), dv0 = umma_desc_k128(smem_u32(sV)), dp0 = umma_desc_k128(smem_u32(sP));
            const uint32_t idesc_pv = umma_idesc_bf16(128, 64);
            // The blocks of all of this CTA's units form ONE stream: QK of block g + 1 is issued before PV of block g even when
            // g + 1 opens the next unit, so the tensor core has the new unit's first scores ready while the softmax warps finish
            // the old one.  A cursor walks (unit, block); `cur` is the block whose QK is in flight, `nxt` the one after it.
            struct Cursor { int u; uint32_t u_it, rem; int k_lo; bool valid; };
            TcUnitWalk walk(blockIdx.x, gridDim.x, tiles, a.H);          // always positioned on unit c.u of the cursor below
            auto load_unit = [&](Cursor& c) {           // position c on the first live block of unit c.u (skipping empty units)
                for (;;) {
                    if (c.u >= n_units) { c.valid = false; return; }
                    const TcUnit un = walk.get();
                    const showo_seq_mask_t msk = mask_of(un.seq);
                    const int q_lo = a.pos0 + un.q0;
                    c.k_lo = tc_key_begin(msk, q_lo);
                    c.rem = tc_live_blocks(msk, q_lo, q_lo + 127, c.k_lo, (n_keys - c.k_lo + 63) >> 6, n_keys);
                    mbar_wait_sleep(&q_full[c.u_it & 1], (c.u_it >> 1) & 1, sleep_ns);
                    if (c.rem != 0) { c.valid = true; return; }
                    mbar_arrive(&q_empty[c.u_it & 1]);   // nothing visible: the softmax warps write zeros
                    c.u += gridDim.x; ++c.u_it; walk.step();
                }
            };
            uint32_t g = 0;
            Cursor cur{(int)blockIdx.x, 0u, 0u, 0, false};
            load_unit(cur);
            auto issue_qk = [&](uint32_t gi, const Cursor& c, int j) {   // S[gi & 1] = Q K_j^T ; the softmax of block gi - 2 (same
                const int st = gi % kTcStages;                            // buffer) has published its P, and this thread has waited for it
                TC_STAMP(1, gi, 0);
                mbar_wait_sleep(&k_full[st], (gi / kTcStages) & 1, sleep_ns);
                TC_STAMP(1, gi, 1);
                tc_fence_after();
                const int nj = min(64, ((n_keys - (c.k_lo + j * 64)) + 15) & ~15);
                const uint32_t idesc_s = umma_idesc_bf16(128, nj);
                const uint64_t dq = dq0 + (uint64_t)((c.u_it & 1) * (16384 >> 4));
                const uint64_t dk = dk0 + (uint64_t)(st * (8192 >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + (gi & 1) * 64, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[gi & 1]);
                umma_commit(&k_empty[st]);
                TC_STAMP(1, gi, 2);
            };
            if (cur.valid) {
                // block g: unit cur.u_it, first-of-unit flag, last-of-unit flag
                bool g_first = true;
                {
                    const int j = __ffs(cur.rem) - 1;
                    cur.rem &= cur.rem - 1;
                    issue_qk(g, cur, j);
                }
                uint32_t g_uit = cur.u_it;
                for (;;) {
                    // ---- look ahead: the block after g (same unit, or the first live block of the next one)
                    const bool g_last = cur.rem == 0;
                    if (g_last) {
                        umma_commit(&q_empty[cur.u_it & 1]);   // the unit's last QK is in flight: its Q buffer is free once it completes
                        cur.u += gridDim.x; ++cur.u_it; walk.step();
                        load_unit(cur);
                    }
                    const bool have_next = cur.valid;
                    uint32_t n_uit = cur.u_it;
                    if (have_next) {
                        const int j = __ffs(cur.rem) - 1;
                        cur.rem &= cur.rem - 1;
                        issue_qk(g + 1, cur, j);
                    }
                    // ---- PV(g)
                    const int st = g % kTcStages;
                    TC_STAMP(1, g, 3);
                    mbar_wait_sleep(&v_full[st], (g / kTcStages) & 1, sleep_ns);
                    mbar_wait_sleep(&p_full[g & 1], (g >> 1) & 1, sleep_ns);
                    TC_STAMP(1, g, 4);
                    if (g_first) mbar_wait_sleep(&o_free[g_uit & 1], ((g_uit >> 1) & 1) ^ 1, sleep_ns);   // the epilogue two units back has read this O buffer
                    tc_fence_after();
                    const uint32_t o_col = tmem_base + 128 + (g_uit & 1) * 64;
                    const uint64_t dp = dp0 + (uint64_t)((g & 1) * (16384 >> 4)), dv = dv0 + (uint64_t)(st * (8192 >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(o_col, dp + 2 * k, dv + 2 * k, idesc_pv, !g_first || k != 0);
                    umma_commit(&pv_done[g & 1]);
                    umma_commit(&v_empty[st]);
                    TC_STAMP(1, g, 5);
                    ++g;
                    if (!have_next) break;
                    g_first = g_last;                    // block g + 1 opens a unit iff block g closed one
                    g_uit = n_uit;
                }
            }
        }
    } else {                      // ================================================================= softmax + epilogue
        // Two threads per query row: warps 2-5 take key columns [0,32) of every 64-key block, warps 6-9 columns [32,64) (same TMEM
        // lane quarter = warp & 3).  Half the per-thread state of a one-thread-per-row layout, so twice the softmax warps fit
        // the register file: the phases of a block (TMEM read, max, 32 exponentials, P store) are latency-bound per warp, and the
        // extra warps are what hides that latency.  The two halves of a row agree on the running maximum through shared memory
        // and a 64-thread named barrier per block; each half keeps its own partial row sum and its own 32 columns of O.
        const int half = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const float sc = a.scale * 1.4426950408889634f;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t bar_id = 1 + quarter;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
        uint32_t g = 0, u_it = 0;
        // the finished unit's epilogue is deferred until the next unit's first P block is published: the wait for the last PV
        // and the O read-out then overlap the tensor core working on the next unit
        struct Pending { bool on; uint32_t g_last, ob; float l, m; bf16* orow; float* lse; } pend{false, 0, 0, 0.f, 0.f, nullptr, nullptr};
        auto flush = [&]() {       // (the partner's partial sum was stored before a pair barrier this thread has passed since)
            if (!pend.on) return;
            pend.on = false;
            uint32_t o0[32];
            const uint32_t t_o = t_row + 128 + pend.ob * 64 + half * 32;
            mbar_wait(&pv_done[pend.g_last & 1], (pend.g_last >> 1) & 1);
            tc_fence_after();
            tmem_ld32(t_o, o0);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_free[pend.ob]);     // this O buffer may be overwritten by the unit after the next
            const float l = pend.l + x_l[(pend.ob * 2 + (half ^ 1)) * 128 + row];
            const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
                uint4 o;
                o.x = pack_bf16(__uint_as_float(o0[c]) * inv, __uint_as_float(o0[c + 1]) * inv);
                o.y = pack_bf16(__uint_as_float(o0[c + 2]) * inv, __uint_as_float(o0[c + 3]) * inv);
                o.z = pack_bf16(__uint_as_float(o0[c + 4]) * inv, __uint_as_float(o0[c + 5]) * inv);
                o.w = pack_bf16(__uint_as_float(o0[c + 6]) * inv, __uint_as_float(o0[c + 7]) * inv);
                *reinterpret_cast<uint4*>(pend.orow + half * 32 + c) = o;
            }
            if (half == 0 && pend.lse != nullptr) *pend.lse = l > 0.f ? pend.m * sc + log2f(l) : 1.0e30f;
        };
        TcUnitWalk walk(blockIdx.x, gridDim.x, tiles, a.H);
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++u_it, walk.step()) {
            const TcUnit un = walk.get();
            const showo_seq_mask_t msk = mask_of(un.seq);
            const int q_lo = a.pos0 + un.q0, q_hi = q_lo + 127;
            const int k_lo = tc_key_begin(msk, q_lo);
            const int nkb = (n_keys - k_lo + 63) >> 6;
            const int r = un.q0 + row;
            const RowKeys rk = omni_row_keys(msk, a.pos0 + r, n_keys);
            const int wq_lo = q_lo + quarter * 32, wq_hi = wq_lo + 31;
            // block classification, one block per lane: live for the tile / nothing for this warp's 32 rows / everything for them
            const int kl = k_lo + lane * 64;
            const bool in = lane < nkb;
            uint32_t rem = __ballot_sync(0xffffffffu, in && !tc_none_allowed(msk, q_lo, q_hi, kl, kl + 64, n_keys));
            const uint32_t none_set = __ballot_sync(0xffffffffu, in && tc_none_allowed(msk, wq_lo, wq_hi, kl, kl + 64, n_keys));
            const uint32_t all_set = __ballot_sync(0xffffffffu, in && tc_all_allowed(msk, wq_lo, wq_hi, kl, kl + 64, n_keys));
            bf16* orow = a.out ? a.out + ((int64_t)un.seq * a.rows_per_seq + r) * a.out_ld + un.h * 64
                               : a.q + ((int64_t)un.seq * a.rows_per_seq + r) * a.ld + un.h * 64;
            float* lse_p = a.lse != nullptr ? a.lse + ((int64_t)un.seq * a.rows_per_seq + r) * a.H + un.h : nullptr;
            if (rem == 0) {       // nothing visible for the whole tile: zeros (keeps the O-buffer hand-shake in step)
                pair_sync();      // the partner's partial sum of the pending unit is visible
                flush();
                mbar_arrive(&o_free[u_it & 1]);
#pragma unroll
                for (int c = 0; c < 32; c += 8) *reinterpret_cast<uint4*>(orow + half * 32 + c) = make_uint4(0, 0, 0, 0);
                if (half == 0 && lse_p) *lse_p = 1.0e30f;
                continue;
            }
            const uint32_t t_o = t_row + 128 + (u_it & 1) * 64 + half * 32;
            float m = kTcNeg, l = 0.f;                 // m: the maximum the stored exponentials refer to (raw score units), same in both halves
            int i = 0;
            while (rem) {
                const int j = __ffs(rem) - 1;
                rem &= rem - 1;
                const int k0 = k_lo + j * 64;
                const bool none = ((none_set >> j) & 1u) || (k0 + half * 32 >= n_keys);       // warp-uniform
                const bool all_ok = (all_set >> j) & 1u;
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 0);
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 1);
                float s[32];
                float bm = kTcNeg;
                if (!none) {
                    uint32_t v0[32];
                    tmem_ld32(t_row + (g & 1) * 64 + half * 32, v0);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) s[c] = __uint_as_float(v0[c]);
                    if (!all_ok) {       // mixed block: the row's allowed keys are two intervals -> a column mask (this half's 32 bits)
                        const uint64_t mb = omni_range_bits(rk.a0, rk.a1, k0) | omni_range_bits(rk.b0, rk.b1, k0);
                        const uint32_t mw = (uint32_t)(mb >> (half * 32));
#pragma unroll
                        for (int c = 0; c < 32; ++c) s[c] = (mw >> c) & 1u ? s[c] : kTcNeg;
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) bm = fmaxf(bm, s[c]);
                }
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 2);
                // lazy maximum, agreed between the two halves of the row: each warp posts "one of my rows exceeds the running
                // maximum by more than the slack" plus its block maxima; if neither warp of the pair raised the flag (the common
                // case) nothing else is exchanged
                const bool need = (bm - m) * sc > kTcLazy;
                const uint32_t slot = (g & 1) * 2;
                const bool w_need = __any_sync(0xffffffffu, need);
                x_bm[(slot + half) * 128 + row] = bm;
                if (lane == 0) x_need[(slot + half) * 4 + quarter] = w_need ? 1 : 0;
                pair_sync();
                if (w_need || x_need[(slot + (half ^ 1)) * 4 + quarter] != 0) {
                    const float bmx = fmaxf(bm, x_bm[(slot + (half ^ 1)) * 128 + row]);
                    const bool raise = (bmx - m) * sc > kTcLazy;
                    const float m_new = raise ? bmx : m;
                    const float alpha = raise ? tc_ex2((m - m_new) * sc) : 1.f;      // m = -1e30 -> 0
                    m = m_new;
                    l *= alpha;
                    if (i > 0) {      // O holds the previous blocks: wait for PV(g-1), then scale this thread's half row in TMEM
                        mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
                        tc_fence_after();
                        uint32_t v[32];
                        tmem_ld32(t_o, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) * alpha);
                        tmem_st32(t_o, v);
                        tmem_st_wait();
                    }
                }
                const float ms = (m == kTcNeg) ? 0.f : -m * sc;
                uint32_t pk[16];
                if (none) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) pk[c] = 0u;
                } else {
                    float l0 = 0.f, l1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        const float p0 = tc_ex2(fmaf(s[c], sc, ms)), p1 = tc_ex2(fmaf(s[c + 1], sc, ms));
                        l0 += p0; l1 += p1;
                        pk[c >> 1] = pack_bf16(p0, p1);
                    }
                    l += l0 + l1;
                }
                // the P buffer was last read by PV(g-2)
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 3);
                if (g >= 2) mbar_wait(&pv_done[g & 1], ((g - 2) >> 1) & 1);
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 4);
                uint8_t* prow = sP + (g & 1) * 16384 + row * 128;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<uint4*>(prow + (((half * 4 + ch) ^ (row & 7)) * 16)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
                tc_fence_before();                 // order my tcgen05.ld/st before the MMA thread's next instructions
                fence_async_smem();                // generic-proxy smem stores -> visible to the tensor core
                mbar_arrive(&p_full[g & 1]);
                if (warp == 2 && lane == 0) TC_STAMP(0, g, 5);
                if (i == 0) flush();               // the previous unit's epilogue, behind this unit's first published block
                ++g;
                ++i;
            }
            x_l[((u_it & 1) * 2 + half) * 128 + row] = l;         // read by the partner in its (deferred) epilogue of this unit
            pend = Pending{true, g - 1, u_it & 1, l, m, orow, lse_p};
        }
        pair_sync();                               // the partner's last partial sum is visible
        flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
    // every MMA and TMA of this CTA has completed (the softmax warps waited for the last PV before the barrier above)
    if (n_tail > 0) tc_tail_phase(a, smem, bars + 22, n_tail, tiles * 128, work_ctr);
}

bool attention_tc_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_ATTN_TC"); v = (e && atoi(e) == 0) ? 0 : 1; }     // default on; SHOWO_ATTN_TC=0: mma.sync only
    return v == 1;
}
// number of leading rows of every sequence the tcgen05 kernel takes (full 128-row tiles); the rest goes to the mma.sync kernel
int attention_tc_rows(const AttnArgs& a) {
    if (!attention_tc_enabled() || (a.ld % 8) != 0 || a.Lmax % 64 != 0 || a.n_keys < 1 || a.n_keys > 2048) return 0;   // <= 32 key blocks per tile
    if (a.out != nullptr && (a.out_ld % 8) != 0) return 0;
    return (a.rows_per_seq / 128) * 128;
}

// number of trailing rows (of every sequence) the tcgen05 kernel's tail phase takes when the leading rows fill >= 1 tile
int attention_tc_tail_rows(const AttnArgs& a) {
    const int rows = attention_tc_rows(a);
    const int n_tail = a.rows_per_seq - rows;
    return (rows > 0 && n_tail > 0 && n_tail <= kTailMax && a.Lmax % 8 == 0) ? n_tail : 0;
}

// work counter of the tail phase: one per device for launches without an engine (the test entry points), the engine passes its own
static int* tail_counter() {
    static int* ctr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int*& c = ctr[dev & 63];
    if (!c) {
        if (cudaMalloc(&c, 64) != cudaSuccess) return nullptr;
        cudaMemset(c, 0, 64);
    }
    return c;
}

int omni_attention_tc(const AttnArgs& a, cudaStream_t st) {
    constexpr int kSmem = 32768 + 2 * kTcStages * 8192 + 32768 + 256 + kTcMaskSlots * 20 + 8 * 128 * 4 + 64 + 1024;
    static PerDeviceOnce once;
    if (once.need()) SHOWO_CUDA_OK(cudaFuncSetAttribute(omni_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    CUtensorMap mq, mk, mv;
    const int D = a.H * 64;
    const uint64_t q_rows = (uint64_t)a.n_seq * a.rows_per_seq;
    SHOWO_TRY(make_tmap_2d(&mq, a.q, (uint64_t)D, q_rows, (uint64_t)a.ld * 2, 64, 128));
    SHOWO_TRY(make_tmap_2d(&mk, a.kcache, 64, (uint64_t)a.n_seq * a.H * a.Lmax, 128, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mv, a.vtcache, (uint64_t)a.Lmax, (uint64_t)a.n_seq * a.H * 64, (uint64_t)a.Lmax * 2, 64, 64));
    // persistent: two CTAs per SM, each walking its share of the (sequence, head, tile) units round robin
    const int tiles = a.rows_per_seq / 128;
    const int n_units = a.n_seq * a.H * tiles;
    dim3 grid(std::min(n_units, 2 * gemm_num_sms()));
    static int prof = -1;
    if (prof < 0) { const char* e = getenv("SHOWO_TC_PROF"); prof = e ? atoi(e) : 0; }
    static int sleep_ns = -1;
    if (sleep_ns < 0) { const char* e = getenv("SHOWO_TC_SLEEP"); sleep_ns = e ? atoi(e) : 32; }
    int n_tail = attention_tc_tail_rows(a);
    {   // timing probes only (tests/attn_trace.py): leave the tail rows out to see the tensor-core part on its own
        static int skip = -1;
        if (skip < 0) { const char* e = getenv("SHOWO_ATTN_SKIP_TAIL"); skip = e ? atoi(e) : 0; }
        if (skip) n_tail = 0;
    }
    int* ctr = a.work_ctr ? a.work_ctr : tail_counter();
    SHOWO_CHECK(n_tail == 0 || ctr != nullptr, "attention: could not allocate the tail-phase work counter");
    SHOWO_CUDA_OK(launch_kernel(omni_attention_tc_kernel, grid, dim3(kTcThreads), kSmem, st, 1, mq, mk, mv, a, n_units, tiles, prof, (uint32_t)sleep_ns,
                                n_tail, ctr));
    note_launch();
    return 0;
}

}  // namespace showo

// debug: the clock64 timeline CTA 0 of the last SHOWO_TC_PROF=1 launch wrote: [role 0 softmax warp 2 / 1 MMA / 2 TMA][block < 64][6 stamps]
extern "C" __attribute__((visibility("default"))) int showo_debug_tc_prof(unsigned long long* out_host, int n) {
    if (n > 3 * 64 * 6) n = 3 * 64 * 6;
    return cudaMemcpyFromSymbol(out_host, showo::g_tc_prof, (size_t)n * 8) == cudaSuccess ? 0 : -1;
}
