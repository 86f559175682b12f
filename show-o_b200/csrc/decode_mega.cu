// One decode step of all transformer layers in ONE persistent kernel (mmu_generate, one new token per sequence, M <= 16).
//
// The per-kernel decode path (LayerNorm -> skinny GEMM1 -> attention -> skinny GEMM2, 4 launches x 24 layers) is
// latency-bound: every kernel streams only 40-60 MB and pays its own ramp and drain.  Here one CTA per SM stays resident
// for the whole step, the phases are separated by grid barriers instead of launches,
//
//   per layer:  LN rows (CTAs < M)  |B|  GEMM1 stream-K + q/k-LN/rotary/KV-scatter/gelu epilogue  |B|
//               attention: (sequence, head) units round-robin, 64-key chunks, online softmax  |B|
//               GEMM2 stream-K + residual  |B|
//
// and EVERYTHING that comes from HBM -- weight chunks of both GEMMs and the K / V^T chunks of the attention -- flows through
// one shared-memory ring (10 x 20 KB per SM) fed by a single producer thread that never stops: a slot is refilled as
// soon as it frees, across phase boundaries, so while the CTAs sit in a barrier the first chunks of the next phase are
// already landing.  Only what the previous phase produces is gated: the 4 KB X boxes of a GEMM chunk, and the one K / V
// chunk of each attention unit that holds the current token.  Measured loaded latency is ~3.5 us per ring round trip, so
// bytes in flight are what set the bandwidth -- hence one deep ring instead of per-phase buffers.
// The GEMM phases are the streamed skinny GEMM of gemv.cu (same consumer code, same deterministic stream-K fix-up).
//
// Cross-proxy ordering: activations / cache rows are written with generic stores and read back by TMA (async proxy) in the
// next phase, so every grid barrier is bracketed by fence.proxy.async.
#include <algorithm>
#include <vector>

#include "skinny.cuh"

namespace showo {

struct MegaLayer {
    const float *b1, *b2, *ln_g, *ln_b, *qg, *qb, *kg, *kb;
    bf16 *kc, *vc;
};

struct MegaArgs {
    const MegaLayer* layers;                                // device array [NL]
    int NL, M, D, F, H, W1N;
    float* x; bf16* xh; bf16* buf;
    float ln_eps;
    const float* fln_g; const float* fln_b;                // final LayerNorm -> xh (input of the head GEMM)
    const float* cos_tab; const float* sin_tab;
    int pos0, n_keys, Lmax, cache_seqs;                     // cache_seqs: sequences per layer slab of the KV cache
    const showo_seq_mask_t* masks; float scale;
    float* partials; int* tickets;
    unsigned long long* bar; unsigned long long bar_base;   // monotonic grid-barrier counter, barriers done before this launch
    int stages;
    unsigned long long* prof;                               // optional [grid][NL][8] globaltimer stamps (SHOWO_MEGA_PROF=1)
};

constexpr float kMegaNegBig = -1.0e30f;

__device__ __forceinline__ bool mega_allowed(const showo_seq_mask_t& m, int q, int k) {      // == omni_allowed (attention.cu)
    const bool ok = (k <= q) | ((q >= m.full_begin) & (q < m.full_end)) | ((k >= m.win_begin) & (k < m.win_end));
    return ok & !((k < m.pad_end) & (q >= m.pad_end));
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MEGA_STAMP(k) do { if (a.prof && tid == 0) a.prof[((size_t)cta * a.NL + l) * 8 + (k)] = global_ns(); } while (0)

// all 128 consumer threads; `target` = value the monotonic counter reaches when every CTA has arrived
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __threadfence();
    fence_proxy_async();
    consumer_bar();
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1ULL);
        while (ld_acquire_u64(ctr) < target) { }
        __threadfence();
    }
    consumer_bar();
    fence_proxy_async();
}

// LayerNorm of one fp32 row -> bf16 (phi.py:776 / :1065), 128 threads
__device__ __forceinline__ void mega_ln_row(const float* xr, const float* g, const float* b, float eps, bf16* out, int D,
                                            float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nv = D >> 2;
    float s = 0.f;
    for (int i = tid; i < nv; i += 128) {
        const float4 v = reinterpret_cast<const float4*>(xr)[i];
        s += (v.x + v.y) + (v.z + v.w);
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    consumer_bar();
    const float mean = ((red[0] + red[1]) + (red[2] + red[3])) / (float)D;
    float q = 0.f;
    for (int i = tid; i < nv; i += 128) {
        const float4 v = reinterpret_cast<const float4*>(xr)[i];
        const float a0 = v.x - mean, a1 = v.y - mean, a2 = v.z - mean, a3 = v.w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    q = warp_sum(q);
    if (lane == 0) red[4 + warp] = q;
    consumer_bar();
    const float rstd = rsqrtf(((red[4] + red[5]) + (red[6] + red[7])) / (float)D + eps);
    for (int i = tid; i < nv; i += 128) {
        const float4 v = reinterpret_cast<const float4*>(xr)[i];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i), b4 = __ldg(reinterpret_cast<const float4*>(b) + i);
        uint2 o;
        o.x = pack_bf16((v.x - mean) * rstd * g4.x + b4.x, (v.y - mean) * rstd * g4.y + b4.y);
        o.y = pack_bf16((v.z - mean) * rstd * g4.z + b4.z, (v.w - mean) * rstd * g4.w + b4.w);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
    consumer_bar();                               // `red` is reused by the next call
}

// ---------------------------------------------------------------------------------------------- attention phase
// One ring item = one 64-key chunk of one (sequence, head) unit: K tile [64 keys][64 dims] at stage + 0 and V^T tile
// [64 dims][64 keys] at stage + 8192, both 128 B rows in the TMA SWIZZLE_128B layout (16 B chunk c of row r sits at
// c ^ (r & 7)).  128 threads: QK with 8 lanes per key (one 128 B row per quarter-warp), online softmax across chunks,
// PV with two threads per output dim.  Scores / reductions are double-buffered by chunk parity (two barriers per chunk).
constexpr int kAttnChunk = 64;
constexpr int kAttnBytes = 16384;

struct MegaAttnSmem { float* sc_s; float* red; };     // sc_s [2][64], red [2][8]

__device__ __forceinline__ void mega_attention(const MegaArgs& a, const Sk2Smem& sm, const MegaAttnSmem& am, int cta, int grid,
                                               RingPos& rp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_units = a.M * a.H;
    const int n_chunks = (a.n_keys + kAttnChunk - 1) / kAttnChunk;
    const int sub = lane & 7, kq = lane >> 3;
    const int d = tid >> 1, half = tid & 1;
    const float sc = a.scale * 1.4426950408889634f;
    const uint32_t ring_u32 = smem_u32(sm.ring);
    int par = 0;
    for (int u = cta; u < n_units; u += grid) {
        const int seq = u / a.H, h = u % a.H;
        const showo_seq_mask_t msk = a.masks[seq];
        bf16* qrow = a.buf + (int64_t)seq * a.W1N + 2 * a.D + h * 64;
        float q[8];
        {
            const uint4 v = *reinterpret_cast<const uint4*>(qrow + sub * 8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h2[j]); q[2 * j] = f.x; q[2 * j + 1] = f.y; }
        }
        float m_run = kMegaNegBig, l_run = 0.f, acc = 0.f;
        for (int j = 0; j < n_chunks; ++j, par ^= 1) {
            const int s = rp.stage;
            mbar_wait(&sm.full[s], rp.phase);
            const uint32_t kt = ring_u32 + (uint32_t)s * kSk2StageBytes, vt = kt + 8192;
            float* scs = am.sc_s + par * 64;
            float* red = am.red + par * 8;
            // ---- scores of the 64 keys: warp w takes keys w*4 + 16*i + kq
            float mx = kMegaNegBig;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kl = warp * 4 + 16 * i + kq, k = j * kAttnChunk + kl;
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(kt + kl * 128 + ((sub ^ (kl & 7)) << 4)));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
                float dot = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    dot += q[2 * e] * f.x + q[2 * e + 1] * f.y;
                }
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                const float sv = (k < a.n_keys && mega_allowed(msk, a.pos0, k)) ? dot * sc : kMegaNegBig;
                if (sub == 0) scs[kl] = sv;
                mx = fmaxf(mx, sv);
            }
            mx = warp_max(mx);
            if (lane == 0) red[warp] = mx;
            consumer_bar();
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])));
            const float scale = exp2f(m_run - m_new);         // both -1e30 before the first visible key: exp2(0) = 1 on zeros
            m_run = m_new;
            if (tid < 64) {
                const float sv = scs[tid];
                const float pv = sv <= 0.5f * kMegaNegBig ? 0.f : exp2f(sv - m_new);
                scs[tid] = pv;
                const float ps = warp_sum(pv);
                if (lane == 0) red[4 + warp] = ps;
            }
            consumer_bar();
            l_run = l_run * scale + (red[4] + red[5]);
            // ---- acc[d] = acc[d] * scale + sum_k p[k] V^T[d][k]; this thread: 16 B chunks half, half+2, ...
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ch = half + 2 * c;
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(vt + d * 128 + ((ch ^ (d & 7)) << 4)));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    part += scs[ch * 8 + 2 * e] * f.x + scs[ch * 8 + 2 * e + 1] * f.y;
                }
            }
            acc = acc * scale + part;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
            ring_advance(rp, sm.stages);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        // q of this unit was read by every thread before the first chunk's barriers: safe to overwrite in place
        if (half == 0) qrow[d] = __float2bfloat16(acc / l_run);
    }
}

// mw1 / mw2: all layers' W1 [NL * W1N, D] / W2 [NL * D, D + F] as ONE tensor each (layer l starts at row l * W1N / l * D);
// mk: the K cache as rows of 64 dims [NL * cache_seqs * H * Lmax, 64]; mv: the V^T cache [NL * cache_seqs * H * 64, Lmax]
__global__ void __launch_bounds__(kSk2Threads, 1) decode_mega_kernel(const __grid_constant__ CUtensorMap mw1,
                                                                      const __grid_constant__ CUtensorMap mw2,
                                                                      const __grid_constant__ CUtensorMap mx1,
                                                                      const __grid_constant__ CUtensorMap mx2,
                                                                      const __grid_constant__ CUtensorMap mk,
                                                                      const __grid_constant__ CUtensorMap mv, MegaArgs a) {
    extern __shared__ uint8_t mega_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mega_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = base;
    float* part = reinterpret_cast<float*>(base + (size_t)a.stages * kSk2StageBytes);     // [4][16][72]
    uint64_t* full = reinterpret_cast<uint64_t*>(part + 4 * 16 * kSk2PartStride);
    uint64_t* empty = full + kSk2MaxStages;
    float* sc_s = reinterpret_cast<float*>(empty + kSk2MaxStages);     // [2][64]
    float* red = sc_s + 128;                                           // [2][8]
    int* s_flag = reinterpret_cast<int*>(red + 16);
    volatile int* s_ready = s_flag + 1;                                // phases (3 per layer) whose gated inputs are complete

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, grid = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        mbar_fence_init();
        *s_ready = 0;
        tma_prefetch_desc(&mw1); tma_prefetch_desc(&mw2); tma_prefetch_desc(&mx1); tma_prefetch_desc(&mx2);
        tma_prefetch_desc(&mk); tma_prefetch_desc(&mv);
    }
    __syncthreads();
    pdl_trigger();

    // per-phase schedules: grid <= total so that every CTA below `grid` owns work (the stream-K fix-up counts on it)
    Skinny2Sched sc1{}, sc2{};
    sc1.tiles = a.W1N / 64; sc1.cpt = a.D / kSk2ChunkK; sc1.stages = a.stages; sc1.total = sc1.tiles * sc1.cpt;
    sc1.grid = grid < sc1.total ? grid : sc1.total;
    sc2.tiles = a.D / 64; sc2.cpt = (a.D + a.F) / kSk2ChunkK; sc2.stages = a.stages; sc2.total = sc2.tiles * sc2.cpt;
    sc2.grid = grid < sc2.total ? grid : sc2.total;
    const int b1 = sk2_begin(sc1, cta), e1 = sk2_begin(sc1, cta + 1);
    const int b2 = sk2_begin(sc2, cta), e2 = sk2_begin(sc2, cta + 1);
    const int n1 = e1 - b1, n2 = e2 - b2;
    const int n_units = a.M * a.H;
    const int my_units = cta < n_units ? (n_units - cta + grid - 1) / grid : 0;
    const int n_chunks = (a.n_keys + kAttnChunk - 1) / kAttnChunk;
    const int na = my_units * n_chunks;

    if (warp == 4) {
        // ------------------------------------------------------------------------------------------------ producer
        // The CTA's items in consumption order: per layer n1 GEMM1 chunks, na attention chunks, n2 GEMM2 chunks.  Two
        // cursors walk that sequence: `w` arms the slot and issues the FREE part of an item as soon as the slot is empty
        // (weight boxes; K / V^T chunks of keys written by earlier steps), `x` follows with the GATED part once the phase
        // that produces it is complete (X boxes; the K / V^T chunk that holds the current token).
        struct Cursor { int layer, ph, idx, tile, kc; RingPos rp; };     // ph: 0 GEMM1, 1 attention, 2 GEMM2
        const int cnt[3] = {n1, na, n2};
        auto enter = [&](Cursor& c) {             // skip empty phases, set (tile, kc) at the first item of the phase
            while (c.layer < a.NL && cnt[c.ph] == 0) { if (++c.ph == 3) { c.ph = 0; ++c.layer; } }
            c.idx = 0;
            if (c.ph == 0) { c.tile = b1 / sc1.cpt; c.kc = b1 % sc1.cpt; }
            else if (c.ph == 2) { c.tile = b2 / sc2.cpt; c.kc = b2 % sc2.cpt; }
            else { c.tile = 0; c.kc = 0; }        // attention: tile = unit ordinal of this CTA, kc = key chunk
        };
        auto advance = [&](Cursor& c) {
            const int lim = c.ph == 0 ? sc1.cpt : (c.ph == 2 ? sc2.cpt : n_chunks);
            if (++c.kc == lim) { c.kc = 0; ++c.tile; }
            ring_advance(c.rp, a.stages);
            if (++c.idx == cnt[c.ph]) { if (++c.ph == 3) { c.ph = 0; ++c.layer; } enter(c); }
        };
        if (lane == 0 && n1 + na + n2 > 0) {
            Cursor w{0, 0, 0, 0, 0, RingPos{0, 0u}}, x{0, 0, 0, 0, 0, RingPos{0, 0u}};
            enter(w); enter(x);
            int pending = 0;                      // items armed by `w` whose gated part `x` has not handled yet
            bool waited = false;
            int seen_ready = 0;
            const int last_chunk = n_chunks - 1;  // the key chunk that holds the current token
            while (x.layer < a.NL) {
                bool progressed = false;
                if (w.layer < a.NL && pending < a.stages) {
                    const int s = w.rp.stage;
                    if (mbar_test(&empty[s], w.rp.phase ^ 1u)) {
                        uint8_t* st = ring + (size_t)s * kSk2StageBytes;
                        if (w.ph == 1) {
                            mbar_arrive_expect_tx(&full[s], kAttnBytes);
                            if (w.kc != last_chunk) {
                                const int unit = cta + w.tile * grid;
                                const int slab = w.layer * a.cache_seqs * a.H + unit;
                                tma_load_2d(st, &mk, &full[s], 0, slab * a.Lmax + w.kc * kAttnChunk);
                                tma_load_2d(st + 8192, &mv, &full[s], w.kc * kAttnChunk, slab * 64);
                            }
                        } else {
                            const bool g2 = w.ph == 2;
                            const CUtensorMap* mw = g2 ? &mw2 : &mw1;
                            const int row = w.layer * (g2 ? a.D : a.W1N) + w.tile * 64;
                            mbar_arrive_expect_tx(&full[s], kSk2StageBytes);
                            tma_load_2d(st, mw, &full[s], w.kc * kSk2ChunkK, row);
                            tma_load_2d(st + 8192, mw, &full[s], w.kc * kSk2ChunkK + 64, row);
                        }
                        advance(w);
                        ++pending;
                        progressed = true;
                    }
                }
                if (pending > 0) {
                    const bool gated = x.ph != 1 || x.kc == last_chunk;
                    const int phase = 3 * x.layer + x.ph;
                    if (!gated || phase < seen_ready || *s_ready > phase) {
                        if (gated) {
                            if (phase >= seen_ready) {        // first gated box of a new phase: order the other CTAs' generic
                                if (!waited) { pdl_wait(); waited = true; }   // stores before this thread's async-proxy reads
                                __threadfence();
                                fence_proxy_async();
                                seen_ready = phase + 1;
                            }
                            const int s = x.rp.stage;
                            uint8_t* st = ring + (size_t)s * kSk2StageBytes;
                            if (x.ph == 1) {
                                const int unit = cta + x.tile * grid;
                                const int slab = x.layer * a.cache_seqs * a.H + unit;
                                tma_load_2d(st, &mk, &full[s], 0, slab * a.Lmax + x.kc * kAttnChunk);
                                tma_load_2d(st + 8192, &mv, &full[s], x.kc * kAttnChunk, slab * 64);
                            } else {
                                const CUtensorMap* mx = x.ph == 2 ? &mx2 : &mx1;
                                tma_load_2d(st + kSk2WBytes, mx, &full[s], x.kc * kSk2ChunkK, 0);
                                tma_load_2d(st + kSk2WBytes + 2048, mx, &full[s], x.kc * kSk2ChunkK + 64, 0);
                            }
                        }
                        advance(x);
                        --pending;
                        progressed = true;
                    }
                }
                // Nothing to issue: if only a ring slot is missing, block on its barrier (hardware-assisted wake-up: a polled
                // sleep here costs more than a whole chunk); if a phase flag is missing, spin on it.
                if (!progressed && pending == 0 && w.layer < a.NL) mbar_wait(&empty[w.rp.stage], w.rp.phase ^ 1u);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------------------------------------- consumers
    pdl_wait();
    Sk2Smem sm{ring, part, full, empty, s_flag, s_flag + 2, a.stages};
    MegaAttnSmem am{sc_s, red};
    RingPos rp{0, 0u};
    unsigned long long nbar = a.bar_base;
    const unsigned long long g64 = (unsigned long long)grid;
    for (int l = 0; l < a.NL; ++l) {
        const MegaLayer lay = a.layers[l];
        MEGA_STAMP(0);
        if (cta < a.M) mega_ln_row(a.x + (int64_t)cta * a.D, lay.ln_g, lay.ln_b, a.ln_eps, a.xh + (int64_t)cta * a.D, a.D, red);
        MEGA_STAMP(1);
        grid_barrier(a.bar, ++nbar * g64);
        if (tid == 0) { __threadfence(); *s_ready = 3 * l + 1; }
        MEGA_STAMP(2);
        {
            SkinnyParams p{};
            p.M = a.M; p.N = a.W1N; p.K = a.D; p.splits = 1; p.kc = a.D;
            p.out = a.buf; p.ldc = a.W1N; p.bias = lay.b1;
            p.partials = a.partials; p.tickets = a.tickets;
            p.qf.D = a.D; p.qf.H = a.H; p.qf.rows_per_seq = 1; p.qf.pos0 = a.pos0; p.qf.Lmax = a.Lmax;
            p.qf.q_gamma = lay.qg; p.qf.q_beta = lay.qb; p.qf.k_gamma = lay.kg; p.qf.k_beta = lay.kb; p.qf.eps = a.ln_eps;
            p.qf.cos_tab = a.cos_tab; p.qf.sin_tab = a.sin_tab; p.qf.kcache = lay.kc; p.qf.vtcache = lay.vc;
            sk2_consume<SK_QKV>(p, sc1, sm, cta, b1, e1, rp);
        }
        MEGA_STAMP(3);
        grid_barrier(a.bar, ++nbar * g64);
        if (tid == 0) { __threadfence(); *s_ready = 3 * l + 2; }
        MEGA_STAMP(4);
        mega_attention(a, sm, am, cta, grid, rp);
        MEGA_STAMP(5);
        grid_barrier(a.bar, ++nbar * g64);
        if (tid == 0) { __threadfence(); *s_ready = 3 * l + 3; }
        MEGA_STAMP(6);
        {
            SkinnyParams p{};
            p.M = a.M; p.N = a.D; p.K = a.D + a.F; p.splits = 1; p.kc = p.K;
            p.out = a.x; p.ldc = a.D; p.bias = lay.b2; p.resid = a.x; p.ldr = a.D;
            p.partials = a.partials; p.tickets = a.tickets;
            sk2_consume<SK_RESID_F32>(p, sc2, sm, cta, b2, e2, rp);
        }
        MEGA_STAMP(7);
        grid_barrier(a.bar, ++nbar * g64);
    }
    if (cta < a.M) mega_ln_row(a.x + (int64_t)cta * a.D, a.fln_g, a.fln_b, a.ln_eps, a.xh + (int64_t)cta * a.D, a.D, red);
}

// ------------------------------------------------------------------------------------------------------------ host
struct DecodeMega {
    MegaLayer* d_layers = nullptr; unsigned long long* d_bar = nullptr;
    unsigned long long bars_done = 0;
    int NL = 0;
    const void* key = nullptr;                    // first W1 pointer the maps were built for
    const void* cache_k = nullptr;                // first / last layer K-cache pointers the layer block was filled with
    const void* cache_k_last = nullptr;
};
static DecodeMega g_mega_dev[64];                 // one per device
static DecodeMega& cur_mega() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_mega_dev[dev & 63];
}
#define g_mega (cur_mega())

// Opt-in (SHOWO_DECODE_MEGA=1): parity-green, but on B200 it measures 1.85 ms per decode step against 1.47 ms for the
// per-kernel path -- every phase still pays ~5-7 us of unhidden latency (gated X boxes at its head, the stream-K hand-over
// at its tail) plus a ~2 us grid barrier with 3-5 us of imbalance on top, which is more than the launch gaps it removes.
// Read on every call so that tests can switch it.
static int mega_enabled() {
    const char* e = getenv("SHOWO_DECODE_MEGA");
    return e ? atoi(e) : 0;
}

bool decode_mega_supported(const DecodeMegaDesc& d) {
    if (!mega_enabled()) return false;
    if (d.M < 1 || d.M > 16 || d.D % kSk2ChunkK != 0 || (d.D + d.F) % kSk2ChunkK != 0 || d.W1N != 3 * d.D + d.F) return false;
    if (d.D != d.H * 64 || d.Lmax % 8 != 0 || d.NL < 1 || d.D % 64 != 0 || d.W1N % 64 != 0) return false;
    for (int l = 0; l < d.NL; ++l)                 // the weights must sit back to back in the slabs
        if (d.layers[l].w1 != d.w1_slab + (size_t)l * d.W1N * d.D || d.layers[l].w2 != d.w2_slab + (size_t)l * d.D * (d.D + d.F)) return false;
    if (d.Lmax % 64 != 0 || d.max_keys > d.Lmax || d.n_keys > d.max_keys || d.cache_seqs < d.M) return false;
    const size_t cstride = (size_t)d.cache_seqs * d.H * d.Lmax * 64;             // KV cache: one slab per layer, back to back
    for (int l = 0; l < d.NL; ++l)
        if (d.layers[l].kc != d.layers[0].kc + l * cstride || d.layers[l].vc != d.layers[0].vc + l * cstride) return false;
    return true;
}

static size_t mega_smem_bytes(int stages) { return 1024 + (size_t)stages * kSk2StageBytes + 4 * 16 * kSk2PartStride * 4 + 1024; }

int decode_mega_step(const DecodeMegaDesc& d, cudaStream_t st) {
    SHOWO_CHECK(decode_mega_supported(d), "decode_mega_step: unsupported geometry");
    const int grid = gemm_num_sms();
    if (g_mega.key != d.layers[0].w1 || g_mega.NL != d.NL) {
        // per-layer parameter block, built once per engine
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (g_mega.d_layers) cudaFree(g_mega.d_layers);
        SHOWO_CUDA_OK(cudaMalloc(&g_mega.d_layers, (size_t)d.NL * sizeof(MegaLayer)));
        if (!g_mega.d_bar) {
            SHOWO_CUDA_OK(cudaMalloc(&g_mega.d_bar, 8));
            SHOWO_CUDA_OK(cudaMemset(g_mega.d_bar, 0, 8));
            g_mega.bars_done = 0;
        }
        g_mega.key = d.layers[0].w1; g_mega.NL = d.NL;
        g_mega.cache_k = nullptr;
    }
    if (g_mega.cache_k != d.layers[0].kc || g_mega.cache_k_last != d.layers[d.NL - 1].kc) {
        // the KV cache was (re)allocated: refresh the per-layer cache pointers
        std::vector<MegaLayer> hl(d.NL);
        for (int l = 0; l < d.NL; ++l) {
            const DecodeMegaLayer& s = d.layers[l];
            hl[l] = MegaLayer{s.b1, s.b2, s.ln_g, s.ln_b, s.qg, s.qb, s.kg, s.kb, s.kc, s.vc};
        }
        SHOWO_CUDA_OK(cudaMemcpyAsync(g_mega.d_layers, hl.data(), hl.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice, st));
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));   // hl is a stack temporary
        g_mega.cache_k = d.layers[0].kc; g_mega.cache_k_last = d.layers[d.NL - 1].kc;
    }
    MegaArgs a{};
    a.layers = g_mega.d_layers;
    a.NL = d.NL; a.M = d.M; a.D = d.D; a.F = d.F; a.H = d.H; a.W1N = d.W1N;
    a.x = d.x; a.xh = d.xh; a.buf = d.buf; a.ln_eps = d.ln_eps; a.fln_g = d.fln_g; a.fln_b = d.fln_b;
    a.cos_tab = d.cos_tab; a.sin_tab = d.sin_tab;
    a.pos0 = d.pos0; a.n_keys = d.n_keys; a.Lmax = d.Lmax;
    a.cache_seqs = d.cache_seqs;
    a.masks = d.masks; a.scale = d.scale;
    const int tiles = std::max(d.W1N / 64, d.D / 64);
    SHOWO_TRY(skinny_workspace(grid, tiles, &a.partials, &a.tickets, st));
    a.bar = g_mega.d_bar; a.bar_base = g_mega.bars_done;
    static int stages_env = -1;
    if (stages_env < 0) { const char* se = getenv("SHOWO_MEGA_STAGES"); stages_env = se ? atoi(se) : kSk2MaxStages; }
    const int stages = std::min(std::max(stages_env, 2), kSk2MaxStages);
    a.stages = stages;
    const size_t smem = mega_smem_bytes(stages);
    static PerDeviceOnce once;
    if (once.need()) {
        SHOWO_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int per_sm = 0;
        SHOWO_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, kSk2Threads, smem));
        SHOWO_CHECK(per_sm >= 1, "decode_mega_step: kernel does not fit on an SM");
    }
    static int prof_mode = -1, prof_calls = 0;
    static unsigned long long* d_prof = nullptr;
    if (prof_mode < 0) { const char* pe = getenv("SHOWO_MEGA_PROF"); prof_mode = pe ? atoi(pe) : 0; }
    const bool prof_now = prof_mode && ++prof_calls == 40;
    if (prof_now) {
        SHOWO_CUDA_OK(cudaMalloc(&d_prof, (size_t)grid * d.NL * 8 * 8));
        SHOWO_CUDA_OK(cudaMemset(d_prof, 0, (size_t)grid * d.NL * 8 * 8));
        a.prof = d_prof;
    }
    CUtensorMap mw1, mw2, mx1, mx2, mk, mv;
    SHOWO_TRY(make_tmap_2d(&mw1, d.w1_slab, (uint64_t)d.D, (uint64_t)g_mega.NL * d.W1N, (uint64_t)d.D * 2, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mw2, d.w2_slab, (uint64_t)(d.D + d.F), (uint64_t)g_mega.NL * d.D, (uint64_t)(d.D + d.F) * 2, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mx1, d.xh, (uint64_t)d.D, (uint64_t)d.M, (uint64_t)d.D * 2, 64, 16));
    SHOWO_TRY(make_tmap_2d(&mx2, d.buf + 2 * (size_t)d.D, (uint64_t)(d.D + d.F), (uint64_t)d.M, (uint64_t)d.W1N * 2, 64, 16));
    const uint64_t units = (uint64_t)g_mega.NL * d.cache_seqs * d.H;
    SHOWO_TRY(make_tmap_2d(&mk, d.layers[0].kc, 64, units * d.Lmax, 128, 64, 64));
    SHOWO_TRY(make_tmap_2d(&mv, d.layers[0].vc, (uint64_t)d.Lmax, units * 64, (uint64_t)d.Lmax * 2, 64, 64));
    SHOWO_CUDA_OK(launch_kernel(decode_mega_kernel, dim3(grid), dim3(kSk2Threads), smem, st, 1, mw1, mw2, mx1, mx2, mk, mv, a));
    g_mega.bars_done += 4ULL * (unsigned long long)d.NL;
    note_launch();
    if (prof_now) {
        // phase breakdown from the in-kernel stamps: per phase, mean over layers of (mean | min | max over CTAs) in us
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        std::vector<unsigned long long> h((size_t)grid * d.NL * 8);
        SHOWO_CUDA_OK(cudaMemcpy(h.data(), d_prof, h.size() * 8, cudaMemcpyDeviceToHost));
        const char* names[8] = {"LN", "barrier A", "GEMM1", "barrier B", "attention", "barrier C", "GEMM2", "barrier D"};
        double tot = 0;
        for (int k = 0; k < 8; ++k) {
            double mean = 0, mn = 0, mx = 0; int cnt = 0;
            for (int l = 0; l + (k == 7 ? 1 : 0) < d.NL; ++l) {
                double sm = 0, lo = 1e30, hi = 0;
                for (int c = 0; c < grid; ++c) {
                    const unsigned long long t0 = h[((size_t)c * d.NL + l) * 8 + k];
                    const unsigned long long t1 = k == 7 ? h[((size_t)c * d.NL + l + 1) * 8] : h[((size_t)c * d.NL + l) * 8 + k + 1];
                    const double dt = (double)(t1 - t0) * 1e-3;
                    sm += dt; lo = std::min(lo, dt); hi = std::max(hi, dt);
                }
                mean += sm / grid; mn += lo; mx += hi; ++cnt;
            }
            fprintf(stderr, "[mega prof] %-10s mean %7.2f us   min-CTA %7.2f   max-CTA %7.2f\n", names[k], mean / cnt, mn / cnt, mx / cnt);
            tot += mean / cnt;
        }
        fprintf(stderr, "[mega prof] per layer %.2f us (stages %d, n_keys %d)\n", tot, stages, d.n_keys);
    }
    return 0;
}

}  // namespace showo
