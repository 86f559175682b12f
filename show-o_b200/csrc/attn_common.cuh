// Shared device helpers of the attention kernels (attention.cu forward, attention_bwd.cu backward): the omni-mask
// predicate of showo_seq_mask_t and its tile classifications, ldmatrix / mma.sync wrappers.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace showo {

__device__ __forceinline__ bool omni_allowed(const showo_seq_mask_t& m, int q, int k) {
    const bool ok = (k <= q) | ((q >= m.full_begin) & (q < m.full_end)) | ((k >= m.win_begin) & (k < m.win_end));
    return ok & !((k < m.pad_end) & (q >= m.pad_end));
}
// conservative: can ANY (q in [q_lo,q_hi], k in [k_lo,k_hi)) pair be allowed?
__device__ __forceinline__ bool omni_tile_possible(const showo_seq_mask_t& m, int q_lo, int q_hi, int k_lo, int k_hi) {
    if (k_hi <= m.pad_end && q_lo >= m.pad_end) return false;
    const bool causal = k_lo <= q_hi;
    const bool full = (q_hi >= m.full_begin) && (q_lo < m.full_end);
    const bool win = (k_lo < m.win_end) && (k_hi > m.win_begin);
    return causal || full || win;
}

// is EVERY (q in [q_lo,q_hi], k in [k_lo,k_hi)) pair allowed (so the per-element predicate can be skipped)?
__device__ __forceinline__ bool omni_tile_all_allowed(const showo_seq_mask_t& m, int q_lo, int q_hi, int k_lo, int k_hi,
                                                      int n_keys) {
    if (k_hi > n_keys) return false;
    if (k_lo < m.pad_end && q_hi >= m.pad_end) return false;       // some pad column x some row past the pads
    const bool causal = (k_hi - 1) <= q_lo;
    const bool full = (q_lo >= m.full_begin) && (q_hi < m.full_end);
    const bool win = (k_lo >= m.win_begin) && (k_hi <= m.win_end);
    return causal || full || win;
}

// The keys one query row may attend to, as (at most) two half-open intervals [a0,a1) U [b0,b1) of [0, n_keys):
//   row past the left pads : [pad_end, ...) only;  bidirectional row : everything up to n_keys;
//   otherwise              : the causal prefix [.., q] plus the always-visible window.
struct RowKeys { int a0, a1, b0, b1; };
__device__ __forceinline__ RowKeys omni_row_keys(const showo_seq_mask_t& m, int q, int n_keys) {
    const int lo = q >= m.pad_end ? m.pad_end : 0;
    RowKeys r;
    r.a0 = lo;
    if (q >= m.full_begin && q < m.full_end) { r.a1 = n_keys; r.b0 = 0; r.b1 = 0; return r; }
    r.a1 = min(q + 1, n_keys);
    r.b0 = max(m.win_begin, lo);
    r.b1 = max(min(m.win_end, n_keys), r.b0);
    return r;
}
// bit c of the result: is key k0 + c inside [lo, hi) ?   (64 keys per block)
__device__ __forceinline__ uint64_t omni_range_bits(int lo, int hi, int k0) {
    lo = max(lo - k0, 0); hi = min(hi - k0, 64);
    return hi > lo ? ((~0ull >> (64 - (hi - lo))) << lo) : 0ull;
}
// 1-D bulk copy global -> shared, completing on an mbarrier (size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kTileK = 64;     // keys per smem tile
constexpr int kPad = 72;       // padded smem row (elements) -> conflict-free 32-bit fragment loads
constexpr float kNegBig = -1.0e30f;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(smem_row)));
}

}  // namespace showo
