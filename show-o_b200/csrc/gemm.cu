// Host launchers for the tcgen05 GEMM / implicit-GEMM convolution (gemm_tcgen05.cuh).
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "gemm_tcgen05.cuh"
#include "kernels.h"

namespace showo {

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time dependency on libcuda.so
// (it must dlopen on a CPU-only box for the symbol-export test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr; uint64_t d0, d1, d2, d3, s1, s2, s3; uint32_t b0, b1, b2, b3, rank;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && s1 == o.s1 && s2 == o.s2 &&
               s3 == o.s3 && b0 == o.b0 && b1 == o.b1 && b2 == o.b2 && b3 == o.b3 && rank == o.rank;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr);
        auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.d2); mix(k.d3); mix(k.s1); mix(k.s2); mix(k.s3);
        mix(k.b0); mix(k.b1); mix(k.b2); mix(k.b3); mix(k.rank);
        return (size_t)h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/strides innermost first; strides in bytes for dims 1..rank-1.
static int make_map(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box, int swizzle_bytes = 128) {
    MapKey key{ptr, dims[0], rank > 1 ? dims[1] : 0, rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0,
               rank > 1 ? strides[0] : 0, rank > 2 ? strides[1] : 0, rank > 3 ? strides[2] : 0,
               box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank | ((uint32_t)swizzle_bytes << 8)};
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn fn = get_encode_fn();
    SHOWO_CHECK(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    cuuint32_t estr[4] = {1, 1, 1, 1};
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4];
    for (uint32_t i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
    for (uint32_t i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gd, gs, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " rank " +
                       std::to_string(rank) + " dims " + std::to_string(dims[0]) + "," +
                       std::to_string(rank > 1 ? dims[1] : 0) + " stride " + std::to_string(rank > 1 ? strides[0] : 0));
        return -3;
    }
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps[key] = *out;
    return 0;
}

// A/B switch while tuning: SHOWO_GEMM_BK=32 selects 9 x 24 KB stages (64B swizzle) instead of 4 x 48 KB (128B swizzle)
static int gemm_bk() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_BK"); v = e ? atoi(e) : 128; }   // default: 3 x 64 KB stages, 8 MMAs per stage
    return v;
}
static bool gemm_bk32() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_BK"); v = (e && atoi(e) == 32) ? 1 : 0; }
    return v == 1;
}

// SHOWO_GEMM_CL=2: clusters of 2 CTAs along M with the B tile TMA-multicast (A/B switch while tuning)
static int gemm_cluster() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_CL"); v = e ? atoi(e) : 1; }
    return v;
}

// SHOWO_GEMM_CG=2: CTA pairs issuing cta_group::2 MMAs (M = 256 per pair, B tile split across the pair)
static int gemm_cta_group() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_CG"); v = e ? atoi(e) : 2; }   // default: CTA pairs
    return v;
}

// SHOWO_GEMM_ORDER=0: always sweep M first (the first version); default 1 = sweep N first when A is the larger operand and the tiles
// need more than one wave (GemmParams::n_fast)
static bool gemm_order_auto() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_ORDER"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}

// SHOWO_GEMM_HINT=1 (opt-in): the operand that every wave reads again is loaded with an L2 evict_last hint when it is small enough to live
// in the L2 next to the streaming operand (<= 64 MB).  Measured on one box, back to back: no effect -- DRAM bytes of the four probe shapes
// within 1 % (dense|fc2 222 vs 221 MB), times within noise, bench identical -- so the remaining 1.14x over the algorithmic bytes is not an
// eviction-order effect; the default leaves the policy at evict_normal.
static void gemm_l2_hints(GemmParams& p, int64_t a_bytes, int64_t b_bytes, int tiles, int clusters) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_HINT"); v = (e && atoi(e) == 1) ? 1 : 0; }
    p.hint_a = p.hint_b = 0;
    if (!v || tiles <= clusters) return;                       // one wave: nothing is read twice
    const int64_t kMax = (int64_t)64 << 20;
    if (p.n_fast) { if (b_bytes <= kMax) p.hint_b = kL2EvictLast; }
    else if (a_bytes <= kMax) p.hint_a = kL2EvictLast;
}

// 2-D bf16 tensor map with 128B swizzle for other TMA users (attention_tc.cu): dims = {inner, rows}, row stride in bytes
int make_tmap_2d(void* out_cutensormap, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows) {
    uint64_t d[2] = {inner, rows}, s[1] = {row_stride_bytes};
    uint32_t b[2] = {box_inner, box_rows};
    return make_map(reinterpret_cast<CUtensorMap*>(out_cutensormap), ptr, 2, d, s, b, 128);
}

int gemm_num_sms() {
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (n[dev & 63] == 0) cudaDeviceGetAttribute(&n[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    return n[dev & 63];
}

// SHOWO_GEMM_STREAMK=1 selects stream-K for the residual GEMM.  Off by default: with every cluster at a different k offset the
// operand blocks are no longer shared in L2 while they are hot (A + B of dense|fc2 = 127 MB, the size of the L2): measured
// 170 vs 137 us for 4128 x 2048 x 10240 although 147 instead of 160 k blocks are on the critical path.
static int gemm_streamk() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SHOWO_GEMM_STREAMK"); v = (e && atoi(e) == 1) ? 1 : 0; }
    return v;
}
// parked partial tiles + their flags, one set per device (the library runs one stream of GEMMs per device; flags reset themselves)
struct StreamKWs { float* ws = nullptr; size_t ws_cap = 0; int* flags = nullptr; size_t flags_cap = 0; };
static int streamk_workspace(size_t ws_floats, size_t n_flags, float** ws, int** flags, cudaStream_t st) {
    static StreamKWs g[64];
    int dev = 0;
    SHOWO_CUDA_OK(cudaGetDevice(&dev));
    StreamKWs& w = g[dev & 63];
    if (ws_floats > w.ws_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (w.ws) cudaFree(w.ws);
        w.ws = nullptr;
        SHOWO_CUDA_OK(cudaMalloc(&w.ws, ws_floats * sizeof(float)));
        w.ws_cap = ws_floats;
    }
    if (n_flags > w.flags_cap) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (w.flags) cudaFree(w.flags);
        w.flags = nullptr;
        SHOWO_CUDA_OK(cudaMalloc(&w.flags, n_flags * sizeof(int)));
        SHOWO_CUDA_OK(cudaMemset(w.flags, 0, n_flags * sizeof(int)));
        w.flags_cap = n_flags;
    }
    *ws = w.ws; *flags = w.flags;
    return 0;
}

template <int BN, int EPI, int AMODE, int BK = 64, int CL = 1, int CG = 1>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int num_tiles, cudaStream_t st) {
    using Cfg = GemmCfg<BN, BK, CG>;
    static PerDeviceOnce once;
    auto kern = gemm_tcgen05_kernel<BN, EPI, AMODE, BK, CL, CG>;
    if (once.need()) SHOWO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    int grid = num_tiles < gemm_num_sms() ? num_tiles : gemm_num_sms();
    if constexpr (CL > 1) {
        // num_tiles counts work units (CL m-tiles each); one cluster per unit, as many clusters as fit the SMs
        const int max_clusters = gemm_num_sms() / CL;
        const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
        SHOWO_CUDA_OK(launch_kernel(kern, dim3(clusters * CL), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, CL, ma, mb, p));
        note_launch();
        return 0;
    } else {
        SHOWO_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, 1, ma, mb, p));
        note_launch();
        SHOWO_CUDA_OK(cudaGetLastError());
        return 0;
    }
}

template <int BN, int BK = 64, int CL = 1, int CG = 1>
static int gemm_bn(const GemmArgs& a, GemmEpi epi, cudaStream_t st) {
    CUtensorMap ma, mb;
    uint64_t da[2] = {(uint64_t)a.K, (uint64_t)a.M}, sa[1] = {(uint64_t)a.lda * 2};
    constexpr int kBox = BK > 64 ? 64 : BK;       // TMA boxes are one 128B swizzle atom (64 bf16) wide at most
    uint32_t ba[2] = {kBox, 128};
    SHOWO_TRY(make_map(&ma, a.A, 2, da, sa, ba, kBox * 2));
    uint64_t db[2] = {(uint64_t)a.K, (uint64_t)a.N}, sb[1] = {(uint64_t)a.ldb * 2};
    uint32_t bb[2] = {kBox, (uint32_t)(BN / CL)};
    SHOWO_TRY(make_map(&mb, a.B, 2, db, sb, bb, kBox * 2));
    GemmParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.resid = a.resid; p.ldr = a.ldr; p.gelu_from = a.gelu_from;
    if (epi == GEMM_BIAS_BF16 && a.gelu_mode != 0) {
        SHOWO_CHECK(a.gelu_out != nullptr && (a.gelu_mode == 1 || a.gelu_pre != nullptr) && a.gelu_from % 32 == 0 && a.N % 32 == 0 &&
                    a.gelu_out_ld % 8 == 0 && a.gelu_pre_ld % 8 == 0 && a.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.gelu_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.gelu_pre) & 15) == 0,
                    "gemm: the split gelu epilogues need 32-column aligned regions and 16-byte aligned rows");
        p.gelu_mode = a.gelu_mode; p.gelu_out = a.gelu_out; p.gelu_out_ld = a.gelu_out_ld; p.gelu_pre = a.gelu_pre; p.gelu_pre_ld = a.gelu_pre_ld;
    }
    if (epi == GEMM_RESID_F32 && a.ln_part != nullptr) {
        SHOWO_CHECK(a.ln_xb != nullptr && a.N % 64 == 0 && (a.ln_xb_ld % 8) == 0 && (a.ldc * 4) % 16 == 0 && (a.ldr * 4) % 16 == 0,
                    "gemm: the LayerNorm-statistics epilogue needs N % 64 == 0 and 16-byte aligned rows");
        p.ln_xb = a.ln_xb; p.ln_xb_ld = a.ln_xb_ld; p.ln_part_out = a.ln_part;
    }
    const int tiles = cdiv(cdiv(a.M, 128), CL) * cdiv(a.N, BN);
    p.n_fast = (gemm_order_auto() && a.M > a.N && tiles > gemm_num_sms() / CL) ? 1 : 0;
    if constexpr (CG == 2) gemm_l2_hints(p, (int64_t)a.M * a.K * 2, (int64_t)a.N * a.K * 2, tiles, gemm_num_sms() / CL);
    if constexpr (CG == 2) {
        // stream-K for the residual GEMM when the tiles do not fill whole waves of clusters (dense|fc2: 136 tiles on 74 clusters)
        const int clusters = std::min(tiles, gemm_num_sms() / CL), num_kb = cdiv(a.K, BK);
        if (epi == GEMM_RESID_F32 && gemm_streamk() && tiles > clusters && tiles % clusters != 0 &&
            (long long)tiles * num_kb / clusters >= num_kb) {
            SHOWO_TRY(streamk_workspace((size_t)tiles * 2 * BN * 128, (size_t)tiles * 2, &p.sk_ws, &p.sk_flags, st));
        }
    }
    switch (epi) {
        case GEMM_BIAS_BF16: return launch<BN, EPI_BIAS_BF16, A_PLAIN, BK, CL, CG>(ma, mb, p, tiles, st);
        case GEMM_RESID_F32: return launch<BN, EPI_RESID_F32, A_PLAIN, BK, CL, CG>(ma, mb, p, tiles, st);
        case GEMM_BIAS_F32: return launch<BN, EPI_BIAS_F32, A_PLAIN, BK, CL, CG>(ma, mb, p, tiles, st);
    }
    SHOWO_CHECK(false, "bad epilogue");
}

int gemm_bf16_tn(const GemmArgs& a, cudaStream_t st) {
    SHOWO_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "gemm_tn: empty problem");
    SHOWO_CHECK((a.lda % 8) == 0 && (a.ldb % 8) == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
                "gemm_tn: lda/ldb must be multiples of 8 elements and A/B 16-byte aligned");
    constexpr int BN = 256, BK = 128, CL = 2, CG = 2;
    CUtensorMap ma, mb;
    uint64_t da[2] = {(uint64_t)a.M, (uint64_t)a.K}, sa[1] = {(uint64_t)a.lda * 2};
    uint64_t db[2] = {(uint64_t)a.N, (uint64_t)a.K}, sb[1] = {(uint64_t)a.ldb * 2};
    uint32_t box[2] = {64, 64};
    SHOWO_TRY(make_map(&ma, a.A, 2, da, sa, box, 128));
    SHOWO_TRY(make_map(&mb, a.B, 2, db, sb, box, 128));
    GemmParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K; p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.gelu_from = a.N;
    const int tiles = cdiv(cdiv(a.M, 128), CL) * cdiv(a.N, BN);
    p.n_fast = (gemm_order_auto() && a.M > a.N && tiles > gemm_num_sms() / CL) ? 1 : 0;
    gemm_l2_hints(p, (int64_t)a.M * a.K * 2, (int64_t)a.N * a.K * 2, tiles, gemm_num_sms() / CL);
    return launch<BN, EPI_BIAS_F32, A_MN, BK, CL, CG>(ma, mb, p, tiles, st);
}

int gemm_bf16(const GemmArgs& a, GemmEpi epi, cudaStream_t st) {
    SHOWO_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
    if (a.M <= 16 && a.K % 64 == 0 && a.block_n == 0) return gemm_skinny(a, (int)epi, nullptr, st);   // decode: HBM-bound weight streaming
    SHOWO_CHECK((a.lda % 8) == 0 && (a.ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (16 B)");
    SHOWO_CHECK((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
                "gemm: A/B must be 16-byte aligned");
    int bn = a.block_n;
    if (bn == 0) {
        if (a.M <= 256) bn = 64;              // decode / tiny batches: more CTAs streaming the weights
        else if (a.N >= 2048) bn = 256;
        else if (a.N >= 128) bn = 128;
        else bn = 64;
    }
    if (bn == 256 && gemm_cta_group() == 2 && a.M > 128 && gemm_bk() == 128 && a.K % 128 == 0) return gemm_bn<256, 128, 2, 2>(a, epi, st);
    if (bn == 256 && gemm_cta_group() == 2 && a.M > 128) return gemm_bn<256, 64, 2, 2>(a, epi, st);
    if (bn == 256 && gemm_cluster() == 2 && a.M > 128) return gemm_bn<256, 64, 2>(a, epi, st);
    if (bn == 256 && gemm_bk32()) return gemm_bn<256, 32>(a, epi, st);
    if (bn == 256) return gemm_bn<256>(a, epi, st);
    if (bn == 128) return gemm_bn<128>(a, epi, st);
    if (bn == 64) return gemm_bn<64>(a, epi, st);
    SHOWO_CHECK(false, "gemm: block_n must be 64, 128 or 256");
}

template <int BN, int BK = 64, int CL = 1, int CG = 1>
static int gemm_qkv_bn(const GemmArgs& a, const QkvFuse& f, cudaStream_t st) {
    CUtensorMap ma, mb;
    uint64_t da[2] = {(uint64_t)a.K, (uint64_t)a.M}, sa[1] = {(uint64_t)a.lda * 2};
    constexpr int kBox = BK > 64 ? 64 : BK;       // TMA boxes are one 128B swizzle atom (64 bf16) wide at most
    uint32_t ba[2] = {kBox, 128};
    SHOWO_TRY(make_map(&ma, a.A, 2, da, sa, ba, kBox * 2));
    uint64_t db[2] = {(uint64_t)a.K, (uint64_t)a.N}, sb[1] = {(uint64_t)a.ldb * 2};
    uint32_t bb[2] = {kBox, (uint32_t)(BN / CL)};
    SHOWO_TRY(make_map(&mb, a.B, 2, db, sb, bb, kBox * 2));
    GemmParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K; p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.gelu_from = 3 * f.D;
    p.qkv_D = f.D; p.qkv_H = f.H; p.qkv_rows_per_seq = f.rows_per_seq; p.qkv_pos0 = f.pos0; p.qkv_Lmax = f.Lmax;
    p.q_gamma = f.q_gamma; p.q_beta = f.q_beta; p.k_gamma = f.k_gamma; p.k_beta = f.k_beta; p.qk_eps = f.eps;
    p.cos_tab = f.cos_tab; p.sin_tab = f.sin_tab; p.kcache = f.kcache; p.vtcache = f.vtcache;
    p.ln_part_in = f.ln_part; p.ln_c = f.ln_c; p.ln_eps = f.ln_eps;
    if (f.ln_part != nullptr) SHOWO_CHECK(f.ln_c != nullptr && a.K % 64 == 0, "gemm_qkv: folded LayerNorm needs c_n and K % 64 == 0");
    const int tiles = cdiv(cdiv(a.M, 128), CL) * cdiv(a.N, BN);
    if constexpr (CG == 2) gemm_l2_hints(p, (int64_t)a.M * a.K * 2, (int64_t)a.N * a.K * 2, tiles, gemm_num_sms() / CL);   // M sweeps first: A is re-read by every wave
    return launch<BN, EPI_QKV_BF16, A_PLAIN, BK, CL, CG>(ma, mb, p, tiles, st);
}

int gemm_qkv_bf16(const GemmArgs& a, const QkvFuse& f, cudaStream_t st) {
    if (a.M <= 16 && a.K % 64 == 0) return gemm_skinny(a, 3, &f, st);
    SHOWO_CHECK(a.M > 0 && a.K > 0 && a.N > 3 * f.D, "gemm_qkv: bad problem");
    SHOWO_CHECK((a.lda % 8) == 0 && (a.ldb % 8) == 0 && (a.ldc % 8) == 0, "gemm_qkv: leading dimensions must be multiples of 8");
    SHOWO_CHECK(a.bias != nullptr && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm_qkv: bias must be 16-byte aligned");
    SHOWO_CHECK(f.D % 64 == 0 && a.N % 64 == 0 && f.H * 64 == f.D, "gemm_qkv: D must be H*64 and N a multiple of 64");
    SHOWO_CHECK(f.pos0 + f.rows_per_seq <= f.Lmax && a.M % f.rows_per_seq == 0, "gemm_qkv: rows / positions exceed the KV cache");
    if (f.D % 256 == 0 && a.M > 256 && gemm_cta_group() == 2 && gemm_bk() == 128 && a.K % 128 == 0) return gemm_qkv_bn<256, 128, 2, 2>(a, f, st);
    if (f.D % 256 == 0 && a.M > 256 && gemm_cta_group() == 2) return gemm_qkv_bn<256, 64, 2, 2>(a, f, st);
    if (f.D % 256 == 0 && a.M > 256 && gemm_cluster() == 2) return gemm_qkv_bn<256, 64, 2>(a, f, st);
    if (f.D % 256 == 0 && a.M > 256 && gemm_bk32()) return gemm_qkv_bn<256, 32>(a, f, st);
    if (f.D % 256 == 0 && a.M > 256) return gemm_qkv_bn<256>(a, f, st);
    if (f.D % 128 == 0 && a.M > 256) return gemm_qkv_bn<128>(a, f, st);
    return gemm_qkv_bn<64>(a, f, st);
}

int conv_nhwc_bf16(const ConvArgs& a, cudaStream_t st) {
    SHOWO_CHECK(a.cin % 64 == 0, "conv: cin must be a multiple of 64 (pad the activation)");
    SHOWO_CHECK(a.taps == 9 || a.taps == 1 || a.taps == 4, "conv: taps must be 9 (3x3), 4 (2x2 forward) or 1 (1x1)");
    // spatial patch of 128 output pixels
    int TW = a.W >= 16 ? 16 : a.W;
    int TH = 128 / TW;
    SHOWO_CHECK(TW * TH == 128, "conv: W must be >= 16 or a power of two dividing 128");
    CUtensorMap ma, mb;
    uint64_t da[4] = {(uint64_t)a.cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.NB};
    uint64_t sa[3] = {(uint64_t)a.cin * 2, (uint64_t)a.W * a.cin * 2, (uint64_t)a.H * a.W * a.cin * 2};
    uint32_t ba[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    SHOWO_TRY(make_map(&ma, a.x, 4, da, sa, ba));
    const int cout_pad = cdiv(a.cout, 64) * 64;
    const int BN = (a.cout >= 256) ? 256 : (a.cout >= 128 ? 128 : 64);
    const int K = a.taps * a.cin;
    uint64_t db[2] = {(uint64_t)K, (uint64_t)cout_pad}, sb[1] = {(uint64_t)K * 2};
    uint32_t bb[2] = {64, (uint32_t)BN};
    SHOWO_TRY(make_map(&mb, a.w, 2, db, sb, bb));
    GemmParams p{};
    p.M = a.NB * a.H * a.W; p.N = a.cout; p.K = K;
    p.out = a.out; p.ldc = a.ldc; p.bias = a.bias; p.resid = a.resid; p.ldr = a.ldr; p.gelu_from = a.cout;
    p.conv_H = a.H; p.conv_W = a.W; p.conv_TH = TH; p.conv_TW = TW; p.conv_taps = a.taps; p.conv_cin = a.cin;
    p.conv_pad = a.taps == 9 ? 1 : 0;
    const int tiles = a.NB * cdiv(a.H, TH) * cdiv(a.W, TW) * cdiv(a.cout, BN);
    if (BN == 256) return launch<256, EPI_CONV_BF16, A_CONV3>(ma, mb, p, tiles, st);
    if (BN == 128) return launch<128, EPI_CONV_BF16, A_CONV3>(ma, mb, p, tiles, st);
    return launch<64, EPI_CONV_BF16, A_CONV3>(ma, mb, p, tiles, st);
}

}  // namespace showo
