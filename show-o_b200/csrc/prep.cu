// Kernels of the step immediately BEFORE Showo.forward / t2i_generate: recovering the closed-form omni-mask descriptor
// (showo_seq_mask_t) from the caller's dense [B,1,L,L] attention mask (inference_t2i.py:300,321 hands the dense tensor
// of training/prompting_utils.py:466-511,591-624 to the model) and verifying that it reproduces the tensor bit for bit.
#include "engine_state.h"

namespace showo {

template <class T>
__device__ __forceinline__ bool mask_attend(const T* p) {
    if constexpr (sizeof(T) == 1) return *p != 0;        // bool / uint8: non-zero = attend
    else return *p == (T)0;                              // additive float: 0 = attend
}

// One CTA per sequence.  out[b] = {pad_end, full_begin, full_end, win_begin, win_end, mismatches}
template <class T>
__global__ void __launch_bounds__(256) mask_descriptor_kernel(const T* __restrict__ mask, int64_t batch_stride, int L, int* __restrict__ out) {
    __shared__ int s_min[3], s_max;
    const int b = blockIdx.x, tid = threadIdx.x;
    const T* m = mask + (int64_t)b * batch_stride;
    if (tid == 0) { s_min[0] = s_min[1] = s_min[2] = L; s_max = -1; }
    __syncthreads();
    // pad_end: first column the last row attends to; full_begin: first row (< L-1) that attends to the last column
    int pe = L, fb = L;
    for (int i = tid; i < L; i += 256) {
        if (mask_attend(m + (int64_t)(L - 1) * L + i)) pe = min(pe, i);
        if (i < L - 1 && mask_attend(m + (int64_t)i * L + (L - 1))) fb = min(fb, i);
    }
    atomicMin(&s_min[0], pe);
    atomicMin(&s_min[1], fb);
    __syncthreads();
    const int pad_end = s_min[0];
    const bool any_full = s_min[1] < L;
    const int full_begin = any_full ? s_min[1] : 0, full_end = any_full ? L : 0;
    // always-visible window: what the first non-pad row sees beyond causality (unless that row is inside the full span)
    const int q0 = min(pad_end, L - 1);
    const bool q0_full = q0 >= full_begin && q0 < full_end;
    int wb = L, we = -1;
    if (!q0_full)
        for (int k = q0 + 1 + tid; k < L; k += 256)
            if (mask_attend(m + (int64_t)q0 * L + k)) { wb = min(wb, k); we = max(we, k); }
    atomicMin(&s_min[2], wb);
    atomicMax(&s_max, we);
    __syncthreads();
    const bool any_win = s_max >= 0;
    const int win_begin = any_win ? s_min[2] : 0, win_end = any_win ? s_max + 1 : 0;
    // verification on every non-pad query row
    int bad = 0;
    for (int q = pad_end; q < L; ++q) {
        const bool qfull = q >= full_begin && q < full_end;
        for (int k = tid; k < L; k += 256) {
            bool ok = (k <= q) | qfull | ((k >= win_begin) & (k < win_end));
            ok = ok & !((k < pad_end) & (q >= pad_end));
            bad += (ok != mask_attend(m + (int64_t)q * L + k)) ? 1 : 0;
        }
    }
    bad = __syncthreads_count(bad != 0);
    if (tid == 0) {
        int* o = out + b * 6;
        o[0] = pad_end; o[1] = full_begin; o[2] = full_end; o[3] = win_begin; o[4] = win_end; o[5] = bad;
    }
}

}  // namespace showo

using namespace showo;

extern "C" int showo_mask_descriptors(const void* mask_dev, int elem_bytes, int B, int L, int64_t batch_stride_elems,
                                      showo_seq_mask_t* out_host, int32_t* mismatches_host, void* stream) {
    SHOWO_CHECK(mask_dev && out_host && mismatches_host && B > 0 && L > 0, "mask_descriptors: bad arguments");
    SHOWO_CHECK(elem_bytes == 4 || elem_bytes == 1, "mask_descriptors: fp32 (additive) or 1-byte (bool) masks only");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    int* d = nullptr;
    SHOWO_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d), (size_t)B * 6 * 4, st));
    if (elem_bytes == 4) mask_descriptor_kernel<float><<<B, 256, 0, st>>>((const float*)mask_dev, batch_stride_elems, L, d);
    else mask_descriptor_kernel<uint8_t><<<B, 256, 0, st>>>((const uint8_t*)mask_dev, batch_stride_elems, L, d);
    note_launch();
    std::vector<int> h((size_t)B * 6);
    SHOWO_CUDA_OK(cudaMemcpyAsync(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost, st));
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    cudaFreeAsync(d, st);
    for (int b = 0; b < B; ++b) {
        out_host[b] = showo_seq_mask_t{h[b * 6], h[b * 6 + 1], h[b * 6 + 2], h[b * 6 + 3], h[b * 6 + 4]};
        mismatches_host[b] = h[b * 6 + 5];
    }
    return 0;
}
