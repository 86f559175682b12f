// Shared device code of the M <= 16 weight-streaming GEMM (gemv.cu) and the decode megakernel (decode_mega.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace showo {

enum { SK_BIAS_BF16 = 0, SK_RESID_F32 = 1, SK_BIAS_F32 = 2, SK_QKV = 3, SK_ARGMAX = 4 };

struct SkinnyParams {
    const bf16* X; int64_t lda; const bf16* W; int64_t ldb;
    int M, N, K, splits, kc;             // kc = K per split (multiple of 64)
    void* out; int64_t ldc; const float* bias; const float* resid; int64_t ldr; int gelu_from;
    float* partials; int* tickets;
    QkvFuse qf;
    unsigned long long* argmax_keys;     // SK_ARGMAX: per-row packed (orderable logit, ~index) maxima, atomicMax'ed
    const void* l2_prefetch; size_t l2_prefetch_bytes;      // the next kernel's weights -> L2 (idle producer lanes)
    // SK_RESID_F32 (the layer's second GEMM, x += ...): the CTA that finishes the LAST feature tile also runs the next LayerNorm over
    // the M rows (phi.py:776 / :1065) and writes its bf16 output -- the stand-alone 16-row LayerNorm launch cost 6 us per layer
    bf16* ln_out; const float* ln_gamma; const float* ln_beta; float ln_eps; int* ln_ctr;
    // SK_RESID_F32 feeding a LayerNorm-folded projection (kernels.h GemmArgs::ln_part): raw bf16 copy of the new rows + slot statistics
    bf16* ln_xb; int64_t ln_xb_ld; float* ln_part_out;
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Epilogue on a complete [16 x 64] fp32 tile: thread -> (row = tid/8, 8 columns at (tid%8)*8), f = its 8 pre-bias sums.
// Called by all 128 threads of 4 full warps (the q/k LayerNorm and the argmax reduce across 8-lane groups).
template <int EPI>
__device__ __forceinline__ void skinny_epilogue(const SkinnyParams& p, float (&f)[8], int n0, int tid) {
    const int r = tid >> 3, cq = (tid & 7) * 8;
    const int n = n0 + cq;
    const bool row_ok = r < p.M;
    if constexpr (EPI == SK_QKV) {
        if (p.qf.ln_part != nullptr) {
            // LayerNorm folded into this GEMM (kernels.h QkvFuse): the row's statistics from the 64-column slots the previous
            // residual GEMM wrote, the row's 8 lanes taking every 8th slot;   y = rstd acc - (mu rstd) c_n   (+ d_n as the bias)
            const int slots = p.K >> 6, l8 = tid & 7;
            const float2* ps = reinterpret_cast<const float2*>(p.qf.ln_part) + r;          // [slot][M]
            const float k0 = row_ok ? __ldcg(ps).x : 0.f;
            float s1 = 0.f, s2 = 0.f, sm = 0.f;
            if (row_ok)
                for (int t = l8; t < slots; t += 8) {
                    const float2 v = __ldcg(ps + (int64_t)t * p.M);
                    const float dm = v.x - k0;
                    s1 += dm; s2 = fmaf(dm, dm, s2); sm += v.y;
                }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); sm += __shfl_xor_sync(0xffffffffu, sm, o);
            }
            const float inv = 1.f / (float)slots;
            const float mu = k0 + s1 * inv;
            const float rstd = rsqrtf(fmaxf((sm + 64.f * (s2 - s1 * s1 * inv)) / (float)p.K, 0.f) + p.qf.ln_eps), mrs = mu * rstd;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (n + j < p.N) f[j] = f[j] * rstd - mrs * __ldg(p.qf.ln_c + n + j);
        }
    }
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (n + j < p.N) f[j] += __ldg(p.bias + n + j);
    }
    if constexpr (EPI == SK_QKV) {
        const QkvFuse& q = p.qf;
        const int region = n0 / q.D;                       // 64-feature tile == one head of k / v / q, or 64 fc1 columns
        if (region >= 3) {
            if (row_ok) {
                bf16* o = reinterpret_cast<bf16*>(p.out) + (int64_t)r * p.ldc + n;
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16(gelu_new_f(f[j]));
            }
            return;
        }
        const int seq = r / q.rows_per_seq, pos = q.pos0 + r % q.rows_per_seq;
        if (region == 1) {
            if (row_ok) {
                const int h = (n0 - q.D) >> 6;
                bf16* vt = q.vtcache + ((int64_t)seq * q.H + h) * 64 * (int64_t)q.Lmax + pos;
#pragma unroll
                for (int j = 0; j < 8; ++j) vt[(int64_t)(cq + j) * q.Lmax] = __float2bfloat16(f[j]);
            }
            return;
        }
        // k or q: LayerNorm over the 64 columns of the row (8 threads, lanes differing in bits 0..2), then rotary
        const float* gam = region == 0 ? q.k_gamma : q.q_gamma;
        const float* bet = region == 0 ? q.k_beta : q.q_beta;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s * (1.f / 64.f);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { f[j] -= mean; v += f[j] * f[j]; }
        v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
        const float rstd = rsqrtf(v * (1.f / 64.f) + q.eps);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * rstd * __ldg(gam + cq + j) + __ldg(bet + cq + j);
        // rotate_half pairing (i, i+16) on dims [0,32): column octets 0,1 pair with octets 2,3 (lane xor 2)
        float pr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pr[j] = __shfl_xor_sync(0xffffffffu, f[j], 2);
        if (cq < 32 && row_ok) {
            const int i0 = cq & 15;                       // frequency index of column cq (emb = cat(freqs, freqs))
            const float sgn = cq < 16 ? -1.f : 1.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float c = __ldg(q.cos_tab + (int64_t)pos * 32 + i0 + j), sn = __ldg(q.sin_tab + (int64_t)pos * 32 + i0 + j);
                f[j] = f[j] * c + sgn * pr[j] * sn;
            }
        }
        if (!row_ok) return;
        if (region == 0) {
            const int h = n0 >> 6;
            bf16* kd = q.kcache + (((int64_t)seq * q.H + h) * q.Lmax + pos) * 64 + cq;
#pragma unroll
            for (int j = 0; j < 8; ++j) kd[j] = __float2bfloat16(f[j]);
        } else {
            bf16* o = reinterpret_cast<bf16*>(p.out) + (int64_t)r * p.ldc + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16(f[j]);
        }
    } else {
        if constexpr (EPI == SK_ARGMAX) {
            // greedy next token without materialising the logits: per-row max of this 64-column tile, then one atomicMax of
            // a packed key (order-preserving float bits << 32 | ~index: ties resolve to the smallest index, like argmax)
            float bv = -3.0e38f; int bi = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (n + j < p.N && f[j] > bv) { bv = f[j]; bi = n + j; }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (row_ok && (tid & 7) == 0 && bi != 0x7fffffff) {
                const uint32_t fb = __float_as_uint(bv);
                const uint32_t ord = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
                atomicMax(p.argmax_keys + r, ((unsigned long long)ord << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)bi));
            }
            return;
        }
        if constexpr (EPI == SK_RESID_F32) {
            if (p.ln_part_out != nullptr) {
                // the next layer's LayerNorm is folded into its projection: leave bf16(new residual row) and the (mean, M2) of this
                // 64-column slot (the row's 8 lanes hold it; N % 64 == 0 is checked by the host)
                float xs[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xs[j] = row_ok ? f[j] + p.resid[(int64_t)r * p.ldr + n + j] : 0.f;
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += xs[j];
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
                const float mean = s * (1.f / 64.f);
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float dd = xs[j] - mean; v = fmaf(dd, dd, v); }
                v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (row_ok) {
                    uint4 pk;
                    pk.x = pack_bf16(xs[0], xs[1]); pk.y = pack_bf16(xs[2], xs[3]); pk.z = pack_bf16(xs[4], xs[5]); pk.w = pack_bf16(xs[6], xs[7]);
                    *reinterpret_cast<uint4*>(p.ln_xb + (int64_t)r * p.ln_xb_ld + n) = pk;
                    if ((tid & 7) == 0) reinterpret_cast<float2*>(p.ln_part_out)[(int64_t)(n0 >> 6) * p.M + r] = make_float2(mean, v);
                }
            }
        }
        if (!row_ok) return;
        if constexpr (EPI == SK_BIAS_BF16) {
            bf16* o = reinterpret_cast<bf16*>(p.out) + (int64_t)r * p.ldc + n;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (n + j < p.N) o[j] = __float2bfloat16(n + j >= p.gelu_from ? gelu_new_f(f[j]) : f[j]);
        } else if constexpr (EPI == SK_RESID_F32) {
            float* o = reinterpret_cast<float*>(p.out) + (int64_t)r * p.ldc + n;
            const float* rs = p.resid + (int64_t)r * p.ldr + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (n + j < p.N) { f[j] += rs[j]; o[j] = f[j]; }
        } else {
            float* o = reinterpret_cast<float*>(p.out) + (int64_t)r * p.ldc + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (n + j < p.N) o[j] = f[j];
        }
    }
}


constexpr int kSk2MaxStages = 10;                                     // 10 x 20 KB = 200 KB in flight per SM
constexpr int kSk2ChunkK = 128;
constexpr int kSk2WBytes = 64 * kSk2ChunkK * 2;                       // 16 KB: two [64 x 64] boxes
constexpr int kSk2XBytes = 16 * kSk2ChunkK * 2;                       // 4 KB: two [16 x 64] boxes
constexpr int kSk2StageBytes = kSk2WBytes + kSk2XBytes;               // 20 KB (multiple of 1024: swizzle atoms stay aligned)
constexpr int kSk2PartStride = 72;                                    // floats per row of a warp's partial tile
constexpr int kSk2Threads = 160;                                      // warps 0-3 consume, warp 4 produces
static inline size_t sk2_smem_bytes(int stages) { return 1024 + (size_t)stages * kSk2StageBytes + 4 * 16 * kSk2PartStride * 4 + 256; }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// cpt = chunks per tile = K / 128; total = tiles * cpt chunks are cut into `grid` equal contiguous ranges (grid <= total:
// every CTA below `grid` owns at least one chunk -- the fix-up counts on it; CTAs >= grid own nothing)
struct Skinny2Sched { int tiles, cpt, grid, stages, total; };
__device__ __forceinline__ int sk2_begin(const Skinny2Sched& sc, int cta) {
    return cta >= sc.grid ? sc.total : (int)((long long)cta * sc.total / sc.grid);
}
__device__ __forceinline__ int sk2_cta_of(const Skinny2Sched& sc, int chunk) {
    return (int)((((long long)chunk + 1) * sc.grid - 1) / sc.total);
}
// position in the smem ring, advanced once per chunk by producer and consumers alike
struct RingPos { int stage; uint32_t phase; };
__device__ __forceinline__ void ring_advance(RingPos& rp, int stages) {
    if (++rp.stage == stages) { rp.stage = 0; rp.phase ^= 1u; }
}


// the last arriver of a split tile: sum the contributors' partial tiles in CTA order (deterministic)
__device__ __forceinline__ void sk2_sum_partials(const SkinnyParams& p, const Skinny2Sched& sc, int tile_c0, int first, int last,
                                                 int r, int cq, float (&f)[8]) {
    __threadfence();
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    for (int c = first; c <= last; ++c) {
        const int b0 = sk2_begin(sc, c);
        const float* src = p.partials + ((long long)c * 2 + (b0 <= tile_c0 ? 1 : 0)) * 1024 + r * 64 + cq;
        const float4 u0 = __ldcg(reinterpret_cast<const float4*>(src));
        const float4 u1 = __ldcg(reinterpret_cast<const float4*>(src) + 1);
        f[0] += u0.x; f[1] += u0.y; f[2] += u0.z; f[3] += u0.w; f[4] += u1.x; f[5] += u1.y; f[6] += u1.z; f[7] += u1.w;
    }
}

// Consumer side of the streamed skinny GEMM (warps 0-3, 128 threads): contracts the chunks [c_begin, c_end) of the
// linearised (tile, k chunk) space out of the smem ring and finishes / hands over the tiles it touches.  `rp` is the
// ring position, in step with the producer's issue order (it carries over between phases in the decode megakernel).
struct Sk2Smem { uint8_t* ring; float* part; uint64_t* full; uint64_t* empty; int* s_flag; int* s_defer; int stages; };


// After a finished tile of the residual GEMM: count it; the CTA that completes the last tile normalises the M rows of the residual
// stream (every tile's stores are ordered before its count by the fence; the reads below bypass L1).  128 consumer threads.
template <int EPI>
__device__ __forceinline__ void sk2_tile_done(const SkinnyParams& p, const Skinny2Sched& sc, int* s_flag, int tid) {
    if constexpr (EPI == SK_RESID_F32) {
        if (p.ln_out == nullptr) return;
        consumer_bar();                           // the tile's stores of all 128 threads precede thread 0's fence
        if (tid == 0) {
            __threadfence();
            const int t = atomicAdd(p.ln_ctr, 1);
            const int last = (t == sc.tiles - 1);
            if (last) *p.ln_ctr = 0;              // self-resetting for the next launch
            *s_flag = last;
        }
        consumer_bar();
        const bool last = *s_flag != 0;
        consumer_bar();                           // s_flag may be rewritten
        if (!last) return;
        __threadfence();
        const int warp = tid >> 5, lane = tid & 31;
        const int D = p.N;                        // the GEMM's N is the hidden size
        for (int r = warp; r < p.M; r += 4) {
            const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + (int64_t)r * p.ldc);
            float4 v[16];
            float s = 0.f;
            const int nvec = D >> 7;              // float4 per lane (D <= 2048)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvec) { v[i] = __ldcg(xr + i * 32 + lane); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
            const float mean = warp_sum(s) / (float)D;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvec) {
                    const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
                    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
                }
            const float rstd = rsqrtf(warp_sum(q) / (float)D + p.ln_eps);
            uint2* orow = reinterpret_cast<uint2*>(p.ln_out + (int64_t)r * D);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvec) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma) + i * 32 + lane);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.ln_beta) + i * 32 + lane);
                    uint2 pk;
                    pk.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
                    pk.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
                    orow[i * 32 + lane] = pk;
                }
        }
    }
}

template <int EPI>
__device__ __forceinline__ void sk2_consume(const SkinnyParams& p, const Skinny2Sched& sc, const Sk2Smem& sm, int cta,
                                            int c_begin, int c_end, RingPos& rp) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* ring = sm.ring; float* part = sm.part; uint64_t* full = sm.full; uint64_t* empty = sm.empty; int* s_flag = sm.s_flag;
    int deferred_tile = -1;                       // partial head segment whose ticket result is picked up at the end of the range
    const int kSk2Stages = sm.stages;
    // warp w contracts k16 steps {2w, 2w+1} of every chunk against all 64 features: k16 step ks lives in box ks>>2 at
    // 16 B chunks (ks&3)*2 + {0,1} of the 128 B rows (SWIZZLE_128B: chunk ^= row & 7)
    const int mat = lane >> 3, r8 = lane & 7;
    uint32_t a_off[2], b_off[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int ks = 2 * warp + kk, box = ks >> 2, kq = ks & 3;
        a_off[kk] = kSk2WBytes + box * 2048 + ((mat & 1) * 8 + r8) * 128 + (((kq * 2 + (mat >> 1)) ^ r8) << 4);
        b_off[kk] = box * 8192 + ((mat >> 1) * 8 + r8) * 128 + (((kq * 2 + (mat & 1)) ^ r8) << 4);
    }
    const uint32_t ring_u32 = smem_u32(ring);
    const int g = lane >> 2, t4 = lane & 3;
    int c = c_begin;
    while (c < c_end) {
        const int tile = c / sc.cpt;
        const int tile_c0 = tile * sc.cpt;
        const int seg_begin = c;
        const int seg_end = (tile_c0 + sc.cpt < c_end) ? tile_c0 + sc.cpt : c_end;
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
        for (; c < seg_end; ++c) {
            const int s = rp.stage;
            mbar_wait(&full[s], rp.phase);
            const uint32_t st = ring_u32 + (uint32_t)s * kSk2StageBytes;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                uint32_t af[4];
                ldsm_x4(af, st + a_off[kk]);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    uint32_t bf[4];
                    ldsm_x4(bf, st + b_off[kk] + jp * 2048);          // n-blocks 2jp, 2jp+1 (16 rows x 128 B)
                    mma16816(acc[2 * jp], af, bf[0], bf[1]);
                    mma16816(acc[2 * jp + 1], af, bf[2], bf[3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            ring_advance(rp, kSk2Stages);
        }
        // ---- cross-warp (k split) reduction through smem, then the [16 x 64] tile in the epilogue's thread mapping
        float* mine = part + warp * 16 * kSk2PartStride;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            *reinterpret_cast<float2*>(mine + g * kSk2PartStride + nb * 8 + t4 * 2) = make_float2(acc[nb][0], acc[nb][1]);
            *reinterpret_cast<float2*>(mine + (g + 8) * kSk2PartStride + nb * 8 + t4 * 2) = make_float2(acc[nb][2], acc[nb][3]);
        }
        consumer_bar();
        const int r = tid >> 3, cq = (tid & 7) * 8;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float4 u0 = *reinterpret_cast<const float4*>(part + (w * 16 + r) * kSk2PartStride + cq);
            const float4 u1 = *reinterpret_cast<const float4*>(part + (w * 16 + r) * kSk2PartStride + cq + 4);
            f[0] += u0.x; f[1] += u0.y; f[2] += u0.z; f[3] += u0.w; f[4] += u1.x; f[5] += u1.y; f[6] += u1.z; f[7] += u1.w;
        }
        consumer_bar();                           // `part` may be overwritten by the next segment from here on
        const bool whole = seg_begin == tile_c0 && seg_end == tile_c0 + sc.cpt;
        bool finish = true;
        if (!whole) {
            // partial tile -> workspace slot (slot 1: the segment starts the tile, slot 0: it starts inside it)
            float* ws = p.partials + ((long long)cta * 2 + (seg_begin == tile_c0 ? 1 : 0)) * 1024 + r * 64 + cq;
            reinterpret_cast<float4*>(ws)[0] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(ws)[1] = make_float4(f[4], f[5], f[6], f[7]);
            consumer_bar();                       // the CTA's partial stores are ordered before thread 0's fence + ticket
            const int first = sk2_cta_of(sc, tile_c0), last = sk2_cta_of(sc, tile_c0 + sc.cpt - 1);
            if (seg_end < c_end) {
                // More chunks follow: do not stall the stream on the fence + atomic round trip (~2 us).  Thread 0 takes the
                // ticket in the background; whether this CTA has to finish the tile is looked at after its last chunk.
                if (tid == 0) {
                    __threadfence();
                    const int t = atomicAdd(&p.tickets[tile], 1);
                    const int fin = (t == last - first);
                    if (fin) p.tickets[tile] = 0;
                    *sm.s_defer = fin;
                }
                deferred_tile = tile;
                finish = false;
            } else {
                if (tid == 0) {
                    __threadfence();
                    const int t = atomicAdd(&p.tickets[tile], 1);
                    const int fin = (t == last - first);
                    if (fin) p.tickets[tile] = 0;      // self-resetting for the next launch
                    *s_flag = fin;
                }
                consumer_bar();
                finish = *s_flag != 0;
                consumer_bar();                       // s_flag is rewritten by the next partial segment
                if (finish) sk2_sum_partials(p, sc, tile_c0, first, last, r, cq, f);
            }
        }
        if (finish) { skinny_epilogue<EPI>(p, f, tile * 64, tid); sk2_tile_done<EPI>(p, sc, s_flag, tid); }
    }
    if (deferred_tile >= 0) {
        consumer_bar();                           // thread 0's ticket result is in smem
        if (*sm.s_defer != 0) {
            const int tile_c0 = deferred_tile * sc.cpt;
            float f[8];
            sk2_sum_partials(p, sc, tile_c0, sk2_cta_of(sc, tile_c0), sk2_cta_of(sc, tile_c0 + sc.cpt - 1), tid >> 3, (tid & 7) * 8, f);
            skinny_epilogue<EPI>(p, f, deferred_tile * 64, tid);
            sk2_tile_done<EPI>(p, sc, s_flag, tid);
        }
        consumer_bar();                           // s_defer may be rewritten by the next phase
    }
}

}  // namespace showo
