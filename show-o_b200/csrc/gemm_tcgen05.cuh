// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
//
//   A, B : bf16, K-contiguous ("K-major"); loaded by TMA into 128B-swizzled smem tiles (BLOCK_K = 64 elements).
//   acc  : fp32 in TMEM, two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//   roles: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane; also owns TMEM
//          alloc/dealloc), warps 2..5 = epilogue (TMEM -> registers -> global, one output row per thread).
//   grid : persistent, min(#tiles, #SMs) CTAs, 1 CTA / SM; tiles are walked m-fastest so concurrently running
//          CTAs share the same B (weight) columns in L2.
//
// A-operand modes:
//   A_PLAIN : 2-D tensor map over [M, K] (row stride lda).
//   A_CONV3 : implicit-GEMM 3x3 / 1x1 convolution over an NHWC activation: a 4-D tensor map [C, W, H, N]; an M tile is
//             a TH x TW patch of output pixels of one image, the K loop runs over (tap, channel-chunk) and each tap is
//             the same box shifted by (dy, dx) -- out-of-bounds rows/cols are zero-filled by TMA (= zero padding).
//
// Epilogues (fused, applied on the fp32 accumulator):
//   EPI_BIAS_BF16  : out_bf16 = acc + bias[n], gelu_new on columns >= gelu_from
//   EPI_RESID_F32  : out_f32  = resid[m,n] + acc + bias[n]                    (in place allowed)
//   EPI_BIAS_F32   : out_f32  = acc + bias[n]                                 (logits)
//   EPI_CONV_BF16  : out_bf16 = acc + bias[n] (+ resid_bf16[m,n]); rows are NHWC pixels of the conv tile
//   EPI_QKV_BF16   : the layer's fused k|v|q|fc1 projection: per 64-column head slice, in registers (one row per thread):
//                    k,q: + bias, LayerNorm(64) (weights shared across heads), partial rotary on dims [0,32) at the row's
//                    position; k -> KV cache [seq][H][Lmax][64], q -> out (bf16, row-major); v: + bias -> V^T cache
//                    [seq][H][64][Lmax] (warp = 32 consecutive positions -> coalesced); fc1: + bias, gelu_new -> out.
//                    (phi.py:657-694 q/k/v proj + q/k layernorm + rotary, phi.py:208-211 fc1 + gelu_new)
#pragma once
#include "common.cuh"

namespace showo {

enum { EPI_BIAS_BF16 = 0, EPI_RESID_F32 = 1, EPI_BIAS_F32 = 2, EPI_CONV_BF16 = 3, EPI_QKV_BF16 = 4 };
// A_MN: BOTH operands are [k][features] row-major (features contiguous) -- the weight-gradient GEMM dW = dY^T X contracting over
// the tokens reads dY and X as they lie, through MN-major shared-memory descriptors (no transposed copies)
enum { A_PLAIN = 0, A_CONV3 = 1, A_MN = 2 };

struct GemmParams {
    int M, N, K;
    // outputs
    void* out;              // bf16 or f32, row-major, leading dimension ldc (elements)
    int64_t ldc;
    const float* bias;      // [N] or nullptr
    const void* resid;      // EPI_RESID_F32: f32 [M, ldr]; EPI_CONV_BF16: bf16 [M, ldr] or nullptr
    int64_t ldr;
    int gelu_from;          // EPI_BIAS_BF16: columns >= gelu_from get gelu_new; pass N for none
    int gelu_mode; __nv_bfloat16* gelu_out; int64_t gelu_out_ld; const __nv_bfloat16* gelu_pre; int64_t gelu_pre_ld;   // kernels.h GemmArgs
    // conv mode
    int conv_H, conv_W;     // output spatial size (== input spatial size; stride-1, 'same' padding)
    int conv_TH, conv_TW;   // patch: TH*TW == 128
    int conv_taps;          // 9 (3x3) or 1 (1x1)
    int conv_cin;           // channels per tap; K == taps * cin
    int conv_pad;           // 1 for 3x3 same-padding, 0 for 1x1
    // EPI_QKV_BF16 (fused k|v|q|fc1 projection epilogue): columns [0,D) = k, [D,2D) = v, [2D,3D) = q, [3D,N) = fc1
    int qkv_D, qkv_H, qkv_rows_per_seq, qkv_pos0, qkv_Lmax;
    const float* q_gamma; const float* q_beta; const float* k_gamma; const float* k_beta; float qk_eps;
    const float* cos_tab; const float* sin_tab;     // [max_pos][32]
    __nv_bfloat16* kcache; __nv_bfloat16* vtcache;  // [seq][H][Lmax][64], [seq][H][64][Lmax]
    // stream-K (CTA-pair path, EPI_RESID_F32): the (tile, k block) pairs are linearised and every cluster takes an equal contiguous
    // range, so that 136 tiles on 74 clusters cost 1.84 tile times instead of 2.  A tile cut between two clusters: the cluster holding
    // its TAIL k blocks meets it FIRST in its range and parks the raw fp32 accumulators in sk_ws[unit][cta rank][col][row]; the
    // cluster holding its HEAD k blocks meets it LAST, adds the parked partial (fixed order: deterministic) and runs the epilogue.
    // LayerNorm folding (kernels.h QkvFuse / GemmArgs): producer side (EPI_RESID_F32) and consumer side (EPI_QKV_BF16)
    __nv_bfloat16* ln_xb; int64_t ln_xb_ld; float* ln_part_out;
    const float* ln_part_in; const float* ln_c; float ln_eps;
    float* sk_ws;           // nullptr: classic tile loop
    // Tile order of the persistent loop.  0: consecutive units sweep M under one B tile (B is read from DRAM once; A once per group of
    // concurrently processed B tiles unless it stays in L2) -- right when A is the smaller operand.  1: consecutive units sweep N over
    // one A row block (A read once, B re-read per wave, from L2 when it fits) -- right when A is the larger operand (M > N: the
    // layer's second GEMM, the weight gradients of wide layers, dgrad into a narrow layer).  Measured cause: dense|fc2 at
    // 4128 x 2048 x 10240 moved 267 MB for 194 MB algorithmic in order 0 (A = 85 MB crossed DRAM once per wave).
    int n_fast;
    // L2 eviction priority of the two operand streams (CTA-pair path): the operand every wave reads again (A when M sweeps first, B
    // when N sweeps first) is loaded evict_last so that the streaming operand does not push it out of the L2; 0 = no preference
    uint64_t hint_a, hint_b;
    int* sk_flags;          // [units][2] one flag per (tile, CTA rank), zero between launches
};

template <int BN, int BK_ = 64, int CG = 1>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int BK = BK_;                          // 64 -> 128B-swizzled rows, 32 -> 64B-swizzled rows (more, smaller stages)
    static constexpr int kAB = BM * BK * 2;                 // 16 KB
    static constexpr int kBB = BN * BK * 2 / CG;          // cta_group::2: each CTA of the pair holds half of the B tile
    static constexpr int kStageBytes = kAB + kBB;
    static constexpr int kStages = (CG == 2) ? (BK_ == 128 ? 3 : 6) : (BK_ == 64) ? ((BN == 256) ? 4 : (BN == 128 ? 6 : 8)) : ((BN == 256) ? 9 : (BN == 128 ? 13 : 16));
    static constexpr int kTmemCols = 2 * BN;                // 2 accumulator stages (power of two for BN in {64,128,256})
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
    static constexpr int kThreads = 192;
};

// The work of one CTA (cluster) as a list of segments = (tile, k block range): classic = whole tiles unit0, unit0 + stride, ...;
// stream-K = the cluster's contiguous share of the linearised (tile, k block) space, cut at tile boundaries.
struct GemmSegCursor {
    bool sk; int num_kb, num_units, stride, unit; long long c, c_end;
    __device__ __forceinline__ GemmSegCursor(bool sk_, int unit0, int stride_, int num_units_, int num_kb_)
        : sk(sk_), num_kb(num_kb_), num_units(num_units_), stride(stride_), unit(unit0) {
        const long long total = (long long)num_units_ * num_kb_;
        c = sk_ ? (long long)unit0 * total / stride_ : 0;
        c_end = sk_ ? (long long)(unit0 + 1) * total / stride_ : 0;
    }
    __device__ __forceinline__ bool next(int& u, int& kb0, int& kb1) {
        if (sk) {
            if (c >= c_end) return false;
            u = (int)(c / num_kb); kb0 = (int)(c % num_kb);
            const long long left = c_end - c;
            kb1 = (long long)(num_kb - kb0) <= left ? num_kb : kb0 + (int)left;
            c += kb1 - kb0;
            return true;
        }
        if (unit >= num_units) return false;
        u = unit; kb0 = 0; kb1 = num_kb; unit += stride;
        return true;
    }
};

template <int BN, int EPI, int AMODE, int BK_ = 64, int CL = 1, int CG = 1>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
    using Cfg = GemmCfg<BN, BK_, CG>;
    static_assert(CG == 1 || (CG == 2 && CL == 2 && BN == 256 && (BK_ == 64 || BK_ == 128)), "cta_group::2 needs a CTA pair");
    static_assert(BK_ != 128 || CG == 2, "BK = 128 (two 64-wide sub-boxes per stage) is only wired for the CTA-pair path");
    constexpr int BM = Cfg::BM, BK = Cfg::BK, kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + kStages;            // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;        // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    int tiles_m;
    if constexpr (AMODE == A_CONV3) {
        tiles_m = (p.M / (p.conv_H * p.conv_W)) * cdiv_dev(p.conv_H, p.conv_TH) * cdiv_dev(p.conv_W, p.conv_TW);
    } else {
        tiles_m = (p.M + BM - 1) / BM;
    }
    const int tiles_n = (p.N + BN - 1) / BN;
    // Work unit = CL m-tiles (one per CTA of the cluster) that share one B tile: each CTA TMA-loads 1/CL of the B tile and
    // multicasts it to the whole cluster, which cuts the L2->SM traffic per flop (the measured limiter at CL = 1).
    static_assert(CL == 1 || ((AMODE == A_PLAIN || AMODE == A_MN) && (BN % CL) == 0), "clusters only for the plain GEMM");
    static_assert(AMODE != A_MN || (CG == 2 && BK_ == 128), "token-major operands are wired for the CTA-pair, BK = 128 configuration");
    const int crank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int tiles_mc = (tiles_m + CL - 1) / CL;
    const int num_units = tiles_mc * tiles_n;
    const int unit0 = blockIdx.x / CL, unit_stride = gridDim.x / CL;
    const int num_kb = (AMODE == A_CONV3) ? p.conv_taps * ((p.conv_cin + BK - 1) / BK) : (p.K + BK - 1) / BK;
    const bool sk = (CG == 2 && EPI == EPI_RESID_F32 && AMODE == A_PLAIN) && p.sk_ws != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], CG == 2 ? 1 : CL);   // CL>1 multicast: one arrival per CTA writing into this smem
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(&tmem_full[a], 1);
                mbar_init(&tmem_empty[a], 128 * CG);          // cta_group::2: the leader waits for both CTAs epilogues
            }
            mbar_fence_init();
        }
        __syncwarp();
        if constexpr (CG == 2) tmem_alloc_cg2<Cfg::kTmemCols>(tmem_slot);
        else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();      // peer barriers must be initialised before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // the next kernel may start scheduling its CTAs ...
    pdl_wait();         // ... and this one must not read what its predecessor wrote before that predecessor has completed

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            GemmSegCursor cur(sk, unit0, unit_stride, num_units, num_kb);
            int unit, kb_begin, kb_end;
            const uint64_t hint_a = p.hint_a ? p.hint_a : kL2EvictNormal, hint_b = p.hint_b ? p.hint_b : kL2EvictNormal;
            (void)hint_a; (void)hint_b;
            while (cur.next(unit, kb_begin, kb_end)) {
                const int tm = (p.n_fast ? unit / tiles_n : unit % tiles_mc) * CL + crank, tn = p.n_fast ? unit % tiles_n : unit / tiles_mc;
                int img = 0, y0 = 0, x0 = 0;
                if constexpr (AMODE == A_CONV3) {
                    const int tw = cdiv_dev(p.conv_W, p.conv_TW), th = cdiv_dev(p.conv_H, p.conv_TH);
                    img = tm / (tw * th);
                    const int r = tm % (tw * th);
                    y0 = (r / tw) * p.conv_TH;
                    x0 = (r % tw) * p.conv_TW;
                }
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        // both CTAs load their A tile and their half of B; all bytes are credited to the LEADER barrier
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        if constexpr (AMODE == A_MN) {
                            // 64 k x 64 feature boxes (8 KB, rows of 128 swizzled bytes): box (h, q) = k half h, feature atom q
#pragma unroll
                            for (int h = 0; h < BK / 64; ++h)
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    tma_load_2d_cg2_hint(smem_a + stage * Cfg::kAB + (h * 2 + q) * 8192, &tmap_a, &full_bar[stage], tm * BM + q * 64,
                                                         kb * BK + h * 64, hint_a);
                                    tma_load_2d_cg2_hint(smem_b + stage * Cfg::kBB + (h * 2 + q) * 8192, &tmap_b, &full_bar[stage],
                                                         tn * BN + crank * (BN / 2) + q * 64, kb * BK + h * 64, hint_b);
                                }
                        } else {
#pragma unroll
                        for (int h = 0; h < BK / 64; ++h) {      // a stage = BK/64 sub-tiles of 64 k (one 128B swizzle atom wide)
                            tma_load_2d_cg2_hint(smem_a + stage * Cfg::kAB + h * (BM * 128), &tmap_a, &full_bar[stage], kb * BK + h * 64, tm * BM, hint_a);
                            tma_load_2d_cg2_hint(smem_b + stage * Cfg::kBB + h * ((BN / 2) * 128), &tmap_b, &full_bar[stage], kb * BK + h * 64,
                                                 tn * BN + crank * (BN / 2), hint_b);
                        }
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    if constexpr (AMODE == A_CONV3) {
                        const int cchunks = (p.conv_cin + BK - 1) / BK;
                        const int tap = kb / cchunks, cc = kb % cchunks;
                        // 9 taps: 3x3 'same' (shift -1..1); 4 taps: 2x2 forward window (shift 0..1, the
                        // space-to-depth form of the stride-2 Downsample); 1 tap: 1x1
                        const int dy = (p.conv_taps == 9) ? tap / 3 - 1 : (p.conv_taps == 4 ? tap >> 1 : 0);
                        const int dx = (p.conv_taps == 9) ? tap % 3 - 1 : (p.conv_taps == 4 ? tap & 1 : 0);
                        tma_load_4d(smem_a + stage * Cfg::kAB, &tmap_a, &full_bar[stage], cc * BK, x0 + dx, y0 + dy, img);
                        tma_load_2d(smem_b + stage * Cfg::kBB, &tmap_b, &full_bar[stage], tap * p.conv_cin + cc * BK,
                                    tn * BN);
                    } else {
                        tma_load_2d(smem_a + stage * Cfg::kAB, &tmap_a, &full_bar[stage], kb * BK, tm * BM);
                        if constexpr (CL == 1) {
                            tma_load_2d(smem_b + stage * Cfg::kBB, &tmap_b, &full_bar[stage], kb * BK, tn * BN);
                        } else {   // my 1/CL slice of the B tile, delivered to every CTA of the cluster (tmap_b box = BN/CL rows)
                            tma_load_2d_multicast(smem_b + stage * Cfg::kBB + crank * (Cfg::kBB / CL), &tmap_b, &full_bar[stage],
                                                  kb * BK, tn * BN + crank * (BN / CL), (uint16_t)((1u << CL) - 1));
                        }
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = (AMODE == A_MN) ? umma_idesc_bf16_mn(BM * CG, BN)
                                                   : umma_idesc_bf16(BM * CG, BN);     // cta_group::2: one M = 256 instruction for the pair
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const bool issuer = (CG == 1) || (crank == 0);               // only the pair leader issues 2-SM MMAs
        GemmSegCursor cur(sk, unit0, unit_stride, num_units, num_kb);
        int unit, kb_begin, kb_end;
        for (; issuer && cur.next(unit, kb_begin, kb_end); ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kAB);
                    const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBB);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        constexpr int kRowB = (BK >= 64) ? 128 : BK * 2;            // bytes per smem row (swizzle width)
                        uint64_t da, db;
                        if constexpr (AMODE == A_MN) {
                            // 16 k = two 8-row swizzle atoms (SBO 1024) of the 64-row box k >> 2; the second feature atom is the next box (LBO)
                            da = umma_desc_mnmajor(a_addr + (k >> 2) * 16384 + (k & 3) * 2048, 8192, 1024);
                            db = umma_desc_mnmajor(b_addr + (k >> 2) * 16384 + (k & 3) * 2048, 8192, 1024);
                        } else {
                            da = umma_desc_kmajor<kRowB>(a_addr + (k >> 2) * (BM * 128) + (k & 3) * 32);
                            db = umma_desc_kmajor<kRowB>(b_addr + (k >> 2) * ((BN / CG) * 128) + (k & 3) * 32);
                        }
                        if constexpr (CG == 2) umma_bf16_cg2(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0);
                        else umma_bf16(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0);
                    }
                    if constexpr (CG == 2) {                          // frees the slot / publishes the accumulator in BOTH CTAs
                        umma_commit_cg2(&empty_bar[stage], 3);
                        if (kb == kb_end - 1) umma_commit_cg2(&tmem_full[acc], 3);
                    } else {
                        if constexpr (CL == 1) umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                        else umma_commit_multicast(&empty_bar[stage], (uint16_t)((1u << CL) - 1));   // ... in every CTA of the cluster
                        if (kb == kb_end - 1) umma_commit(&tmem_full[acc]);
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        // 16-byte vector stores / loads need 16 B-aligned rows (odd ldc such as the 58498-wide logits take the scalar path)
        constexpr int kOutElem = (EPI == EPI_BIAS_BF16 || EPI == EPI_CONV_BF16 || EPI == EPI_QKV_BF16) ? 2 : 4;
        bool out_vec_ok = ((p.ldc * kOutElem) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
        if (EPI == EPI_RESID_F32 || (EPI == EPI_CONV_BF16 && p.resid != nullptr))
            out_vec_ok = out_vec_ok && ((p.ldr * kOutElem) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.resid) & 15) == 0);
        int it = 0;
        GemmSegCursor cur(sk, unit0, unit_stride, num_units, num_kb);
        int unit, kb_begin, kb_end;
        for (; cur.next(unit, kb_begin, kb_end); ++it) {
            const int tm = (p.n_fast ? unit / tiles_n : unit % tiles_mc) * CL + crank, tn = p.n_fast ? unit % tiles_n : unit / tiles_mc;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            // stream-K: a segment without the tile's first k block parks its accumulators; one with the first but not the last
            // block finishes the tile with the parked partial of the cluster that met it first
            const bool sk_park = kb_begin > 0;
            const bool sk_join = kb_begin == 0 && kb_end < num_kb;
            // output row of this thread
            int64_t m;
            bool row_ok;
            if constexpr (AMODE == A_CONV3) {
                const int tw = cdiv_dev(p.conv_W, p.conv_TW), th = cdiv_dev(p.conv_H, p.conv_TH);
                const int img = tm / (tw * th);
                const int r = tm % (tw * th);
                const int y = (r / tw) * p.conv_TH + row_in_tile / p.conv_TW;
                const int x = (r % tw) * p.conv_TW + row_in_tile % p.conv_TW;
                row_ok = (y < p.conv_H) && (x < p.conv_W);
                m = ((int64_t)img * p.conv_H + y) * p.conv_W + x;
            } else {
                m = (int64_t)tm * BM + row_in_tile;
                row_ok = m < p.M;
            }
            // folded LayerNorm (EPI_QKV_BF16): this row's mean / rstd from the slot statistics the previous residual GEMM wrote
            // ([slot][row], lanes contiguous), read while the tile's MMAs are still running; one pass, sums shifted by slot 0's mean
            float ln_rstd = 1.f, ln_mrs = 0.f;
            if constexpr (EPI == EPI_QKV_BF16) {
                if (p.ln_part_in != nullptr && row_ok) {
                    const int slots = p.K >> 6;
                    const float2* ps = reinterpret_cast<const float2*>(p.ln_part_in) + m;
                    const float k0 = __ldcg(ps).x;
                    float s1 = 0.f, s2 = 0.f, sm = 0.f;
#pragma unroll 8
                    for (int t = 0; t < slots; ++t) {
                        const float2 v = __ldcg(ps + (int64_t)t * p.M);
                        const float dm = v.x - k0;
                        s1 += dm; s2 = fmaf(dm, dm, s2); sm += v.y;
                    }
                    const float inv = 1.f / (float)slots;
                    const float mu = k0 + s1 * inv;
                    const float var = (sm + 64.f * (s2 - s1 * s1 * inv)) / (float)p.K;
                    ln_rstd = rsqrtf(fmaxf(var, 0.f) + p.ln_eps);
                    ln_mrs = mu * ln_rstd;
                }
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            float* sk_part = nullptr;                 // this thread's row of the parked partial: [col][128 rows] (lanes contiguous)
            if constexpr (CG == 2 && EPI == EPI_RESID_F32 && AMODE == A_PLAIN) {
                if (sk_park || sk_join) sk_part = p.sk_ws + ((size_t)unit * 2 + crank) * (size_t)(BN * 128) + row_in_tile;
                if (sk_park) {
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        uint32_t v[32];
                        __syncwarp();
                        tmem_ld32(taddr0 + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) sk_part[(size_t)(c + j) * 128] = __uint_as_float(v[j]);
                    }
                    __threadfence();
                    asm volatile("bar.sync 2, 128;" ::: "memory");          // the four epilogue warps: all rows of this CTA are stored
                    if (warp == 2 && lane == 0) atomicExch(p.sk_flags + unit * 2 + crank, 1);
                    __syncwarp();
                    tc_fence_before();
                    if (crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);
                    else mbar_arrive(&tmem_empty[acc]);
                    continue;
                }
                if (sk_join) {
                    volatile int* fl = p.sk_flags + unit * 2 + crank;
                    while (*fl == 0) { }
                    __threadfence();
                    asm volatile("bar.sync 2, 128;" ::: "memory");          // everyone has seen the flag before it is cleared
                    if (warp == 2 && lane == 0) *fl = 0;                     // self-resetting for the next launch
                }
            }
            if constexpr (EPI == EPI_QKV_BF16) {
                const int D = p.qkv_D;
                const int seq = row_ok ? (int)(m / p.qkv_rows_per_seq) : 0;
                const int pos = row_ok ? p.qkv_pos0 + (int)(m % p.qkv_rows_per_seq) : 0;
                const int region0 = (tn * BN) / D;          // a BN tile never straddles regions (D % BN == 0)
                float cs[16], sn[16];
                if (region0 == 0 || region0 == 2) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.cos_tab + (int64_t)pos * 32 + j));
                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.sin_tab + (int64_t)pos * 32 + j));
                        cs[j] = c4.x; cs[j + 1] = c4.y; cs[j + 2] = c4.z; cs[j + 3] = c4.w;
                        sn[j] = s4.x; sn[j + 1] = s4.y; sn[j + 2] = s4.z; sn[j + 3] = s4.w;
                    }
                }
#pragma unroll 1
                for (int c = 0; c < BN; c += 64) {
                    uint32_t v0[32], v1[32];
                    __syncwarp();
                    tmem_ld32(taddr0 + c, v0);
                    tmem_ld32(taddr0 + c + 32, v1);
                    tmem_ld_wait();
                    const int n0 = tn * BN + c;
                    if (!(row_ok && n0 < p.N)) continue;
                    float f[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(v0[j]); f[32 + j] = __uint_as_float(v1[j]); }
                    if (p.ln_part_in != nullptr) {        // y = rstd acc - (mu rstd) c_n   (+ d_n below, passed as the bias)
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ln_c + n0 + j));
                            f[j] = f[j] * ln_rstd - ln_mrs * c4.x; f[j + 1] = f[j + 1] * ln_rstd - ln_mrs * c4.y;
                            f[j + 2] = f[j + 2] * ln_rstd - ln_mrs * c4.z; f[j + 3] = f[j + 3] * ln_rstd - ln_mrs * c4.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 64; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                    const int region = n0 / D;
                    if (region >= 3) {                       // fc1 + gelu_new
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] = gelu_new_f(f[j]);
                    } else if (region == 1) {                // v -> transposed cache, nothing to the activation buffer
                        const int h = (n0 - D) >> 6;
                        __nv_bfloat16* vt = p.vtcache + ((int64_t)seq * p.qkv_H + h) * 64 * (int64_t)p.qkv_Lmax + pos;
#pragma unroll
                        for (int j = 0; j < 64; ++j) vt[(int64_t)j * p.qkv_Lmax] = __float2bfloat16(f[j]);
                        continue;
                    } else {                                 // k or q: LayerNorm(64) + partial rotary
                        const float* gam = region == 0 ? p.k_gamma : p.q_gamma;
                        const float* bet = region == 0 ? p.k_beta : p.q_beta;
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 64; ++j) s += f[j];
                        const float mean = s * (1.f / 64.f);
                        float q = 0.f;
#pragma unroll
                        for (int j = 0; j < 64; ++j) { f[j] -= mean; q += f[j] * f[j]; }
                        const float rstd = rsqrtf(q * (1.f / 64.f) + p.qk_eps);
#pragma unroll
                        for (int j = 0; j < 64; ++j) f[j] = f[j] * rstd * __ldg(gam + j) + __ldg(bet + j);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {       // rotate_half pairing (j, j+16), emb = cat(freqs, freqs)
                            const float a = f[j], b = f[j + 16];
                            f[j] = a * cs[j] - b * sn[j];
                            f[j + 16] = b * cs[j] + a * sn[j];
                        }
                        if (region == 0) {                   // k -> cache row [pos][64]
                            const int h = n0 >> 6;
                            __nv_bfloat16* kd = p.kcache + (((int64_t)seq * p.qkv_H + h) * p.qkv_Lmax + pos) * 64;
#pragma unroll
                            for (int j = 0; j < 64; j += 8) {
                                uint4 pk;
                                pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
                                pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
                                *reinterpret_cast<uint4*>(kd + j) = pk;
                            }
                            continue;
                        }
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n0;
#pragma unroll
                    for (int j = 0; j < 64; j += 8) {
                        uint4 pk;
                        pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
                        pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
                        *reinterpret_cast<uint4*>(o + j) = pk;
                    }
                }
            } else {
            float ln_mean_lo = 0.f, ln_m2_lo = 0.f;        // statistics of the first 32-column half of the current 64-column slot
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(taddr0 + c, v);
                tmem_ld_wait();
                const int n0 = tn * BN + c;
                if (row_ok && n0 < p.N) {
                const bool full = (n0 + 32 <= p.N);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if constexpr (CG == 2 && EPI == EPI_RESID_F32 && AMODE == A_PLAIN) {
                    if (sk_join) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += __ldcg(sk_part + (size_t)(c + j) * 128);
                    }
                }
                if (p.bias != nullptr) {
                    if (full && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) f[j] += __ldg(p.bias + n0 + j);
                    }
                }
                if constexpr (EPI == EPI_BIAS_BF16) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n0;
                    if (n0 >= p.gelu_from) {
                        if (p.gelu_mode == 0) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_new_f(f[j]);
                        } else if (full && out_vec_ok) {
                            // training step (kernels.h GemmArgs::gelu_mode); both variants work on the bf16-rounded value, which is what
                            // the stand-alone gelu_fwd / gelu_bwd passes they replace read back from memory
                            __nv_bfloat16* go = p.gelu_out + m * p.gelu_out_ld + (n0 - p.gelu_from);
                            float gv[32];
                            if (p.gelu_mode == 1) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) gv[j] = gelu_new_f(__bfloat162float(__float2bfloat16(f[j])));
                            } else {
                                const __nv_bfloat16* gp = p.gelu_pre + m * p.gelu_pre_ld + (n0 - p.gelu_from);
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    const uint4 pr = *reinterpret_cast<const uint4*>(gp + j);
                                    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pr);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 x = __bfloat1622float2(p2[q]);
                                        gv[j + 2 * q] = __bfloat162float(__float2bfloat16(f[j + 2 * q])) * gelu_new_grad(x.x);
                                        gv[j + 2 * q + 1] = __bfloat162float(__float2bfloat16(f[j + 2 * q + 1])) * gelu_new_grad(x.y);
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 pk;
                                pk.x = pack_bf16(gv[j], gv[j + 1]); pk.y = pack_bf16(gv[j + 2], gv[j + 3]);
                                pk.z = pack_bf16(gv[j + 4], gv[j + 5]); pk.w = pack_bf16(gv[j + 6], gv[j + 7]);
                                *reinterpret_cast<uint4*>(go + j) = pk;
                            }
                            if (p.gelu_mode == 2) continue;            // the raw d act is not kept
                        }
                    }
                    if (full && out_vec_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
                            pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
                            *reinterpret_cast<uint4*>(o + j) = pk;
                        }
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) o[j] = __float2bfloat16(f[j]);
                    }
                } else if constexpr (EPI == EPI_CONV_BF16) {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n0;
                    const __nv_bfloat16* r = p.resid ? reinterpret_cast<const __nv_bfloat16*>(p.resid) + m * p.ldr + n0
                                                     : nullptr;
                    if (full && out_vec_ok) {
                        if (r) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 rr = *reinterpret_cast<const uint4*>(r + j);
                                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 t = __bfloat1622float2(r2[q]);
                                    f[j + 2 * q] += t.x; f[j + 2 * q + 1] += t.y;
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
                            pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
                            *reinterpret_cast<uint4*>(o + j) = pk;
                        }
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) {
                                float t = f[j];
                                if (r) t += __bfloat162float(r[j]);
                                o[j] = __float2bfloat16(t);
                            }
                    }
                } else if constexpr (EPI == EPI_RESID_F32) {
                    float* o = reinterpret_cast<float*>(p.out) + m * p.ldc + n0;
                    const float* r = reinterpret_cast<const float*>(p.resid) + m * p.ldr + n0;
                    if (full && out_vec_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(r + j);
                            f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                            *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        }
                        if (p.ln_part_out != nullptr) {
                            // the next layer's LayerNorm is folded into its projection GEMM: it reads these rows raw, in bf16, and
                            // needs their statistics -- (mean, M2) of every 64-column slot, two 32-column chunks merged (shifted sums)
                            __nv_bfloat16* xb = p.ln_xb + m * p.ln_xb_ld + n0;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 pk;
                                pk.x = pack_bf16(f[j], f[j + 1]); pk.y = pack_bf16(f[j + 2], f[j + 3]);
                                pk.z = pack_bf16(f[j + 4], f[j + 5]); pk.w = pack_bf16(f[j + 6], f[j + 7]);
                                *reinterpret_cast<uint4*>(xb + j) = pk;
                            }
                            const float kk = f[0];
                            float s1 = 0.f, s2 = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) { const float dd = f[j] - kk; s1 += dd; s2 = fmaf(dd, dd, s2); }
                            const float mean32 = kk + s1 * (1.f / 32.f), m2_32 = s2 - s1 * s1 * (1.f / 32.f);
                            if ((c & 32) == 0) { ln_mean_lo = mean32; ln_m2_lo = m2_32; }
                            else {
                                const float dm = mean32 - ln_mean_lo;
                                reinterpret_cast<float2*>(p.ln_part_out)[(int64_t)(n0 >> 6) * p.M + m] =
                                    make_float2(0.5f * (mean32 + ln_mean_lo), ln_m2_lo + m2_32 + 16.f * dm * dm);
                            }
                        }
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) o[j] = f[j] + r[j];
                    }
                } else {  // EPI_BIAS_F32
                    float* o = reinterpret_cast<float*>(p.out) + m * p.ldc + n0;
                    if (full && out_vec_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (n0 + j < p.N) o[j] = f[j];
                    }
                }
                }  // row_ok
            }
            }  // generic epilogues
            __syncwarp();
            tc_fence_before();
            if (CG == 2 && crank != 0) mbar_arrive_remote(&tmem_empty[acc], 0);   // the leader owns the accumulator hand-off
            else mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();      // no CTA may exit while a peer can still multicast into it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_cg2<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace showo
