// Weight-streaming "skinny" GEMM for the decode path (M <= 16 token rows: one new token per sequence):
//     C[M,N] = X[M,K] * W[N,K]^T (+ fused epilogue)
// HBM-bound: every weight is read exactly once, 16 B per thread per load, 8 rows x 128 B per warp-load (100 % sector
// efficiency), several loads in flight per thread.  The 16-row X slab lives in shared memory; the contraction runs on
// tensor cores (mma.sync m16n8k16, M = 16 fits exactly) -- CUDA cores would be compute-bound at 16 rows.
//
// Fragment trick: a thread's 16 B weight load covers 8 consecutive k of ONE feature row, whereas the mma B fragment
// wants k = {2t, 2t+1, 2t+8, 2t+9}.  The contraction order is free, so k is permuted: within a 64-wide chunk thread t4
// owns physical k = t4*16 + 4j + {0,1,2,3} for k16-step j, and the A (X) fragments are gathered from smem with the same
// permutation.
//
// Parallelism: one CTA = 64 output features x one K split (grid = N/64 x splits, >= ~1 CTA wave over 148 SMs even for
// the N = 2048, K = 10240 projection).  Split-K partials go to a workspace; the LAST CTA of a feature tile (atomic ticket)
// sums them in split order -- deterministic -- and runs the epilogue on the complete [16 x 64] fp32 tile.
// Epilogues: bias(+gelu_new) -> bf16, residual + bias -> fp32, bias -> fp32, and the fused k|v|q|fc1 one (a 64-feature
// tile is exactly one attention head: LayerNorm(64) + partial rotary + KV-cache scatter, phi.py:657-694).
#include "skinny.cuh"

namespace showo {

constexpr int kSkThreads = 128;
constexpr int kSkTileN = 64;

template <int EPI>
__global__ void __launch_bounds__(kSkThreads) skinny_gemm_kernel(SkinnyParams p) {
    extern __shared__ __align__(16) uint8_t sk_smem[];
    const int xs_stride = p.kc + 2;                                   // +1 word: conflict-free permuted A-fragment loads
    bf16* xs = reinterpret_cast<bf16*>(sk_smem);                      // [16][kc + 2]
    float* tile = reinterpret_cast<float*>(sk_smem + (size_t)16 * xs_stride * 2 + 16);   // [16][64] fp32
    tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tile) + 15) & ~uintptr_t(15));
    __shared__ int s_last;

    const int tn = blockIdx.x, split = blockIdx.y;
    const int n0 = tn * kSkTileN, k0 = split * p.kc;
    const int kc = min(p.kc, p.K - k0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;

    // ---- start streaming the weights first (they do not depend on X): warp w owns features n0 + w*16 + {0..7} (block 0) and + {8..15} (block 1)
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int f0 = n0 + warp * 16 + g, f1 = f0 + 8;
    const bool f0_ok = f0 < p.N, f1_ok = f1 < p.N;
    const bf16* w0 = p.W + (int64_t)(f0_ok ? f0 : 0) * p.ldb + k0 + t4 * 16;
    const bf16* w1 = p.W + (int64_t)(f1_ok ? f1 : 0) * p.ldb + k0 + t4 * 16;
    const bf16* xr0 = xs + g * xs_stride + t4 * 16;
    const bf16* xr1 = xs + (g + 8) * xs_stride + t4 * 16;
    const int nchunk = kc / 64;
    constexpr int kDepth = 4;                      // chunks in flight per thread (4 x 64 B)
    uint4 wb[kDepth][2][2];
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
        if (d < nchunk) {
            wb[d][0][0] = ldg_stream(w0 + d * 64); wb[d][0][1] = ldg_stream(w0 + d * 64 + 8);
            wb[d][1][0] = ldg_stream(w1 + d * 64); wb[d][1][1] = ldg_stream(w1 + d * 64 + 8);
        }
    }
    // ---- X slab -> smem (rows >= M are zero) while the first weight chunks are in flight
    for (int i = tid; i < 16 * (p.kc / 8); i += kSkThreads) {
        const int r = i / (p.kc / 8), c = (i % (p.kc / 8)) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < p.M && c < kc) v = *reinterpret_cast<const uint4*>(p.X + (int64_t)r * p.lda + k0 + c);
        uint32_t* d = reinterpret_cast<uint32_t*>(xs + r * xs_stride + c);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();

    for (int c0 = 0; c0 < nchunk; c0 += kDepth) {
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int c = c0 + d;
            if (c >= nchunk) break;
            const uint4 a0 = wb[d][0][0], a1 = wb[d][0][1], b0 = wb[d][1][0], b1 = wb[d][1][1];
            if (c + kDepth < nchunk) {             // refill this slot for chunk c + kDepth
                wb[d][0][0] = ldg_stream(w0 + (c + kDepth) * 64); wb[d][0][1] = ldg_stream(w0 + (c + kDepth) * 64 + 8);
                wb[d][1][0] = ldg_stream(w1 + (c + kDepth) * 64); wb[d][1][1] = ldg_stream(w1 + (c + kDepth) * 64 + 8);
            }
            const uint32_t wq0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};   // feature f0: k = t4*16 + 0..15
            const uint32_t wq1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const uint32_t* x0 = reinterpret_cast<const uint32_t*>(xr0 + c * 64);
            const uint32_t* x1 = reinterpret_cast<const uint32_t*>(xr1 + c * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // k16-step j uses physical k = t4*16 + 4j + {0,1 | 2,3}
                uint32_t af[4] = {x0[2 * j], x1[2 * j], x0[2 * j + 1], x1[2 * j + 1]};
                mma16816(acc[0], af, wq0[2 * j], wq0[2 * j + 1]);
                mma16816(acc[1], af, wq1[2 * j], wq1[2 * j + 1]);
            }
        }
    }
    // ---- partial tile -> smem [16][64]
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int col = warp * 16 + b * 8 + t4 * 2;
        tile[g * 64 + col] = acc[b][0]; tile[g * 64 + col + 1] = acc[b][1];
        tile[(g + 8) * 64 + col] = acc[b][2]; tile[(g + 8) * 64 + col + 1] = acc[b][3];
    }
    __syncthreads();
    if (p.splits > 1) {
        float* mine = p.partials + ((int64_t)tn * p.splits + split) * (16 * 64);
        for (int i = tid; i < 16 * 64 / 4; i += kSkThreads)
            reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(tile)[i];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int t = atomicAdd(&p.tickets[tn], 1);
            s_last = (t == p.splits - 1);
            if (s_last) p.tickets[tn] = 0;         // self-resetting for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const float* base = p.partials + (int64_t)tn * p.splits * (16 * 64);
        for (int i = tid; i < 16 * 64 / 4; i += kSkThreads) {       // fixed split order => deterministic sum
            float4 s = __ldcg(reinterpret_cast<const float4*>(base) + i);
            for (int sp = 1; sp < p.splits; ++sp) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)sp * 16 * 64) + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            reinterpret_cast<float4*>(tile)[i] = s;
        }
        __syncthreads();
    }

    // ---- epilogue on the complete tile: thread -> (row = tid/8, 8 columns at (tid%8)*8)
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = tile[(tid >> 3) * 64 + (tid & 7) * 8 + j];
    skinny_epilogue<EPI>(p, f, n0, tid);
}

// ================================================================================================ TMA-streamed variant
// The register-prefetch kernel above keeps only (warps resident) x 256 B of weights in flight per SM and measured
// 1.7 TB/s.  This one decouples the bytes in flight from occupancy: one persistent CTA per SM, a producer thread issuing
// TMA boxes into an 8 x 20 KB shared-memory ring (160 KB in flight per SM), four consumer warps running mma.sync from
// the 128B-swizzled tiles through ldmatrix.  Work is cut stream-K style: the (feature tile, 128-wide k chunk) pairs are
// linearised and every CTA streams an equal contiguous range, so all SMs pull the same number of bytes whatever N and K
// are.  A tile whose chunks span several CTAs is completed by the last arriver (atomic ticket), which sums the partial
// tiles in CTA order -- deterministic.  The 16-row X chunk rides along with each weight chunk (TMA zero-fills rows
// >= M; it is L2-resident), so there is no slab and no K limit.  Under programmatic dependent launch the producer
// starts the weight stream BEFORE griddepcontrol.wait (weights never depend on the predecessor); only the X boxes wait.
template <int EPI>
__global__ void __launch_bounds__(kSk2Threads, 1) skinny2_gemm_kernel(const __grid_constant__ CUtensorMap mw,
                                                                        const __grid_constant__ CUtensorMap mx,
                                                                        SkinnyParams p, Skinny2Sched sc) {
    extern __shared__ uint8_t sk2_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk2_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = base;
    float* part = reinterpret_cast<float*>(base + (size_t)sc.stages * kSk2StageBytes);           // [4][16][72]
    uint64_t* full = reinterpret_cast<uint64_t*>(part + 4 * 16 * kSk2PartStride);
    uint64_t* empty = full + kSk2MaxStages;
    int* s_flag = reinterpret_cast<int*>(empty + kSk2MaxStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kSk2Stages = sc.stages;               // ring depth (runtime: 8 = 160 KB in flight, 4 lets two kernels share an SM)
    const int c_begin = sk2_begin(sc, blockIdx.x), c_end = sk2_begin(sc, blockIdx.x + 1);
    if (tid == 0) {
        for (int i = 0; i < kSk2Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mw); tma_prefetch_desc(&mx);
    }
    __syncthreads();
    pdl_trigger();

    if (warp == 4) {
        // ------------------------------------------------------------------------------------------------ producer
        if (lane > 0 && c_end > c_begin && p.bias) {
            // idle lanes: pull the bias slices of this CTA's tiles into L2 now, so that the epilogue at the tail of the
            // stream does not start with a DRAM miss (weights: no need to wait for the predecessor)
            const int t0 = c_begin / sc.cpt, t1 = (c_end - 1) / sc.cpt;
            for (int t = t0 + (lane - 1) / 2; t <= t1; t += 16) {
                const int n = t * 64 + ((lane - 1) & 1) * 32;
                if (n < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bias + n));
            }
            if constexpr (EPI == SK_QKV) {        // q / k LayerNorm affine terms and the rotary row of this position
                const float* w4[4] = {p.qf.q_gamma, p.qf.q_beta, p.qf.k_gamma, p.qf.k_beta};
                if (lane <= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(w4[(lane - 1) >> 1] + ((lane - 1) & 1) * 32));
                else if (lane == 9) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.qf.cos_tab + (int64_t)p.qf.pos0 * 32));
                else if (lane == 10) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.qf.sin_tab + (int64_t)p.qf.pos0 * 32));
            }
        }
        if (lane > 0 && p.l2_prefetch_bytes)      // idle lanes: the successor kernel's weights -> L2, this CTA's 1/grid share
            l2_prefetch_slice(p.l2_prefetch, p.l2_prefetch_bytes, (int)blockIdx.x * 31 + (lane - 1), (int)gridDim.x * 31);
        if (lane == 0) {
            const int n = c_end - c_begin;
            const int pre = n < kSk2Stages ? n : kSk2Stages;
            int tile = c_begin / sc.cpt, kc = c_begin % sc.cpt;       // running (tile, k chunk) of the next weight box
            for (int it = 0; it < pre; ++it) {                       // weights first: independent of the predecessor kernel
                uint8_t* st = ring + (size_t)it * kSk2StageBytes;
                mbar_arrive_expect_tx(&full[it], kSk2StageBytes);
                tma_load_2d(st, &mw, &full[it], kc * kSk2ChunkK, tile * 64);
                tma_load_2d(st + 8192, &mw, &full[it], kc * kSk2ChunkK + 64, tile * 64);
                if (++kc == sc.cpt) { kc = 0; ++tile; }
            }
            pdl_wait();
            int xkc = c_begin % sc.cpt;
            for (int it = 0; it < pre; ++it) {
                uint8_t* st = ring + (size_t)it * kSk2StageBytes + kSk2WBytes;
                tma_load_2d(st, &mx, &full[it], xkc * kSk2ChunkK, 0);
                tma_load_2d(st + 2048, &mx, &full[it], xkc * kSk2ChunkK + 64, 0);
                if (++xkc == sc.cpt) xkc = 0;
            }
            RingPos rp{0, 0u};                                        // item `pre` reuses stage pre % stages
            for (int it = 0; it < pre; ++it) ring_advance(rp, kSk2Stages);
            for (int it = pre; it < n; ++it) {
                const int s = rp.stage;
                mbar_wait(&empty[s], rp.phase ^ 1u);
                uint8_t* st = ring + (size_t)s * kSk2StageBytes;
                mbar_arrive_expect_tx(&full[s], kSk2StageBytes);
                tma_load_2d(st, &mw, &full[s], kc * kSk2ChunkK, tile * 64);
                tma_load_2d(st + 8192, &mw, &full[s], kc * kSk2ChunkK + 64, tile * 64);
                tma_load_2d(st + kSk2WBytes, &mx, &full[s], kc * kSk2ChunkK, 0);
                tma_load_2d(st + kSk2WBytes + 2048, &mx, &full[s], kc * kSk2ChunkK + 64, 0);
                if (++kc == sc.cpt) { kc = 0; ++tile; }
                ring_advance(rp, kSk2Stages);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------------- consumers
    pdl_wait();                                   // the epilogue reads the residual / writes buffers of the predecessor
    Sk2Smem sm{ring, part, full, empty, s_flag, s_flag + 1, kSk2Stages};
    RingPos rp{0, 0u};
    sk2_consume<EPI>(p, sc, sm, blockIdx.x, c_begin, c_end, rp);
}



static int skinny_stages() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_SKINNY_STAGES"); v = e ? atoi(e) : kSk2MaxStages; if (v < 2 || v > kSk2MaxStages) v = kSk2MaxStages; }
    return v;
}
// SHOWO_SKINNY=1 selects the register-prefetch kernel (A/B switch while tuning)
static int skinny_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_SKINNY"); v = e ? atoi(e) : 2; }
    return v;
}

// Fix-up workspace, one per device (the library runs one stream of decode work per device; the tickets reset themselves)
struct SkinnyWs { float* partials = nullptr; size_t partials_cap = 0; int* tickets = nullptr; size_t tickets_cap = 0; };
static SkinnyWs g_ws[64];
static SkinnyWs& cur_ws() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_ws[dev & 63];
}
#define g_partials (cur_ws().partials)
#define g_partials_cap (cur_ws().partials_cap)
#define g_tickets (cur_ws().tickets)
#define g_tickets_cap (cur_ws().tickets_cap)

static int ensure_skinny_ws(size_t partial_floats, size_t tiles, cudaStream_t st) {
    if (partial_floats > g_partials_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (g_partials) cudaFree(g_partials);
        SHOWO_CUDA_OK(cudaMalloc(&g_partials, partial_floats * 4));
        g_partials_cap = partial_floats;
    }
    if (tiles > g_tickets_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (g_tickets) cudaFree(g_tickets);
        SHOWO_CUDA_OK(cudaMalloc(&g_tickets, tiles * 4));
        SHOWO_CUDA_OK(cudaMemset(g_tickets, 0, tiles * 4));
        g_tickets_cap = tiles;
    }
    return 0;
}

int skinny_workspace(int grid, int tiles, float** partials, int** tickets, cudaStream_t st) {
    SHOWO_TRY(ensure_skinny_ws((size_t)grid * 2 * 1024, (size_t)tiles, st));
    *partials = g_partials; *tickets = g_tickets;
    return 0;
}

static int gemm_skinny2(const GemmArgs& a, int epi, const QkvFuse* qf, int tiles, cudaStream_t st) {
    SHOWO_CHECK((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
                "gemm_skinny: operands must be 16-byte aligned");
    Skinny2Sched sc{};
    SHOWO_CHECK((long long)tiles * (a.K / kSk2ChunkK) < (1ll << 30), "gemm_skinny: problem too large");
    sc.tiles = tiles; sc.cpt = a.K / kSk2ChunkK; sc.total = tiles * sc.cpt;
    sc.grid = std::min(gemm_num_sms(), sc.total);
    sc.stages = skinny_stages();
    const size_t kSk2Smem = sk2_smem_bytes(sc.stages);
    SkinnyParams p{};
    p.X = a.A; p.lda = a.lda; p.W = a.B; p.ldb = a.ldb; p.M = a.M; p.N = a.N; p.K = a.K; p.splits = 1; p.kc = a.K;
    p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.resid = a.resid; p.ldr = a.ldr; p.gelu_from = a.gelu_from;
    if (qf) p.qf = *qf;
    p.argmax_keys = a.argmax_keys;
    p.l2_prefetch = a.l2_prefetch; p.l2_prefetch_bytes = a.l2_prefetch_bytes;
    p.ln_out = nullptr;
    if (epi == SK_RESID_F32 && a.ln_out != nullptr) {
        SHOWO_CHECK(a.N % 128 == 0 && a.N <= 2048 && a.ldc == a.N, "gemm_skinny: fused LayerNorm needs a contiguous hidden size <= 2048");
        p.ln_out = a.ln_out; p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_eps = a.ln_eps;
    }
    if (epi == SK_RESID_F32 && a.ln_part != nullptr) {
        SHOWO_CHECK(a.ln_xb != nullptr && a.N % 64 == 0 && a.ln_xb_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.ln_xb) & 15) == 0,
                    "gemm_skinny: the LayerNorm-statistics epilogue needs N % 64 == 0 and 16-byte aligned bf16 rows");
        p.ln_xb = a.ln_xb; p.ln_xb_ld = a.ln_xb_ld; p.ln_part_out = a.ln_part;
    }
    if (epi == SK_QKV && qf && qf->ln_part != nullptr) SHOWO_CHECK(qf->ln_c != nullptr, "gemm_skinny: folded LayerNorm needs c_n");
    if (epi == SK_ARGMAX) SHOWO_CHECK(a.argmax_keys != nullptr, "gemm_skinny: argmax epilogue needs a key buffer");
    SHOWO_TRY(ensure_skinny_ws((size_t)sc.grid * 2 * 1024, (size_t)tiles + 1, st));
    p.partials = g_partials; p.tickets = g_tickets;
    p.ln_ctr = g_tickets + tiles;                 // one more self-resetting counter behind the tile tickets
    CUtensorMap mw, mx;
    SHOWO_TRY(make_tmap_2d(&mw, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb * 2, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mx, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda * 2, 64, 16));
#define SK2_LAUNCH(E)                                                                                             \
    do {                                                                                                          \
        static PerDeviceOnce once;                                                                                \
        if (once.need())                                                                                          \
            SHOWO_CUDA_OK(cudaFuncSetAttribute(skinny2_gemm_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk2_smem_bytes(kSk2MaxStages))); \
        SHOWO_CUDA_OK(launch_kernel(skinny2_gemm_kernel<E>, dim3(sc.grid), dim3(kSk2Threads), kSk2Smem, st, 1, mw, mx, p, sc)); \
    } while (0)
    switch (epi) {
        case SK_BIAS_BF16: SK2_LAUNCH(SK_BIAS_BF16); break;
        case SK_RESID_F32: SK2_LAUNCH(SK_RESID_F32); break;
        case SK_BIAS_F32: SK2_LAUNCH(SK_BIAS_F32); break;
        case SK_ARGMAX: SK2_LAUNCH(SK_ARGMAX); break;
        case SK_QKV: SHOWO_CHECK(qf && qf->D % 64 == 0 && qf->pos0 + qf->rows_per_seq <= qf->Lmax, "gemm_skinny: bad qkv fuse args");
            SK2_LAUNCH(SK_QKV); break;
        default: SHOWO_CHECK(false, "gemm_skinny: bad epilogue");
    }
#undef SK2_LAUNCH
    note_launch();
    return 0;
}

bool skinny_ln_fold_ok(int K) { return skinny_variant() == 2 && K % kSk2ChunkK == 0; }

int gemm_skinny(const GemmArgs& a, int epi, const QkvFuse* qf, cudaStream_t st) {
    SHOWO_CHECK(a.M >= 1 && a.M <= 16, "gemm_skinny: M must be in [1,16]");
    SHOWO_CHECK(a.K % 64 == 0 && a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm_skinny: K must be a multiple of 64, lda/ldb of 8");
    const int tiles = cdiv(a.N, kSkTileN);
    if (skinny_variant() == 2 && a.K % kSk2ChunkK == 0)
        return gemm_skinny2(a, epi, qf, tiles, st);
    SHOWO_CHECK(a.ln_out == nullptr && a.ln_part == nullptr && !(qf && qf->ln_part),
                "gemm_skinny: the fused / folded LayerNorm needs the TMA-streamed kernel (K % 128 == 0)");
    // enough CTAs to cover the SMs a few times over, K per split a multiple of 64 and <= 2048 (X slab <= 64 KB of smem)
    int splits = 1;
    // at least one full wave of CTAs, X slab <= 64 KB of smem; fewer, longer-streaming CTAs beat many short ones
    while ((tiles * splits < gemm_num_sms() && (a.K / (splits * 2)) >= 512 && (a.K % (splits * 2 * 64)) == 0) ||
           a.K / splits > 2048)
        splits *= 2;
    SHOWO_CHECK(a.K % (splits * 64) == 0, "gemm_skinny: K not divisible into 64-aligned splits");
    SkinnyParams p{};
    p.X = a.A; p.lda = a.lda; p.W = a.B; p.ldb = a.ldb; p.M = a.M; p.N = a.N; p.K = a.K; p.splits = splits; p.kc = a.K / splits;
    p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.resid = a.resid; p.ldr = a.ldr; p.gelu_from = a.gelu_from;
    if (qf) p.qf = *qf;
    p.argmax_keys = a.argmax_keys;
    if (epi == SK_ARGMAX) SHOWO_CHECK(a.argmax_keys != nullptr, "gemm_skinny: argmax epilogue needs a key buffer");
    if (splits > 1) {
        const size_t need = (size_t)tiles * splits * 16 * 64;
        if (need > g_partials_cap) {
            SHOWO_CUDA_OK(cudaStreamSynchronize(st));
            if (g_partials) cudaFree(g_partials);
            SHOWO_CUDA_OK(cudaMalloc(&g_partials, need * 4));
            g_partials_cap = need;
        }
        if ((size_t)tiles > g_tickets_cap) {
            SHOWO_CUDA_OK(cudaStreamSynchronize(st));
            if (g_tickets) cudaFree(g_tickets);
            SHOWO_CUDA_OK(cudaMalloc(&g_tickets, (size_t)tiles * 4));
            SHOWO_CUDA_OK(cudaMemset(g_tickets, 0, (size_t)tiles * 4));
            g_tickets_cap = tiles;
        }
        p.partials = g_partials; p.tickets = g_tickets;
    }
    const size_t smem = (size_t)16 * (p.kc + 2) * 2 + 32 + 16 * 64 * 4;
    dim3 grid(tiles, splits);
#define SK_LAUNCH(E)                                                                                              \
    do {                                                                                                          \
        static PerDeviceOnce once;                                                                                \
        if (once.need())                                                                                          \
            SHOWO_CUDA_OK(cudaFuncSetAttribute(skinny_gemm_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); \
        skinny_gemm_kernel<E><<<grid, kSkThreads, smem, st>>>(p);                                                 \
    } while (0)
    switch (epi) {
        case SK_BIAS_BF16: SK_LAUNCH(SK_BIAS_BF16); break;
        case SK_RESID_F32: SK_LAUNCH(SK_RESID_F32); break;
        case SK_BIAS_F32: SK_LAUNCH(SK_BIAS_F32); break;
        case SK_ARGMAX: SK_LAUNCH(SK_ARGMAX); break;
        case SK_QKV: SHOWO_CHECK(qf && qf->D % 64 == 0 && qf->pos0 + qf->rows_per_seq <= qf->Lmax, "gemm_skinny: bad qkv fuse args");
            SK_LAUNCH(SK_QKV); break;
        default: SHOWO_CHECK(false, "gemm_skinny: bad epilogue");
    }
#undef SK_LAUNCH
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace showo
