// Kernels of the step immediately BEFORE Showo.forward / t2i_generate: recovering the closed-form omni-mask descriptor
// (showo_seq_mask_t) from the caller's dense [B,1,L,L] attention mask (inference_t2i.py:300,321 hands the dense tensor
// of training/prompting_utils.py:466-511,591-624 to the model) and verifying that it reproduces the tensor bit for bit.
#include "engine_state.h"
#include "philox.cuh"

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


// ------------------------------------------------------------------------------------------------ t2i training inputs
// training/utils.py:77-154 mask_or_random_replace_tokens (noise_type "mask", no contiguous-region masking, predict_all_tokens
// off: the configuration of every shipped yaml) and training/prompting_utils.py:39-90 UniversalPrompting.t2i_prompt in one
// launch, one CTA per batch row:
//   mask_prob = max(schedule(t), min_masking_rate);  n = max(1, rint(N * mask_prob))
//   position j is masked  <=>  argsort(rand[b, :])[j] < n   (the reference compares the PERMUTATION, not the rank:
//                                                            for i < n the position rank(i) is masked), stable ties
//   row = [pad ...] [t2i] [bos] text [eos] | soi | image ids (mask_id where masked) | eoi        (left-padded to P = max_text_len + 1)
//   labels = same row with pad -> ignore, image part = (masked ? code : ignore)
struct T2iPrepArgs {
    const int64_t* image_tokens;     // [B, N]  codes already offset by the text vocabulary (train.py:476-477)
    const int64_t* text_ids; const int32_t* text_len; int64_t text_stride;       // tokenised captions, no specials
    int B, N, P;                     // P = max_text_len + 1 (task token included)
    int64_t pad_id, bos_id, eos_id, task_id, soi_id, eoi_id, mask_id, ignore_id;
    float min_masking_rate, cond_dropout_prob;
    int schedule;                    // 0 cosine, 1 linear, 2 pow (schedule_param = exponent), 3: mask_prob given
    float schedule_param;
    const float* timesteps;          // [B] uniform draws (torch.rand(batch_size)) or, schedule == 3, the mask_prob values; NULL -> Philox
    const float* rand;               // [B, N] torch.rand(batch_size, seq_len) or NULL -> Philox
    const float* drop_probs;         // [B] torch.rand(len(text_ids)) of t2i_prompt or NULL -> Philox
    uint64_t seed;
    int do_mask, do_prompt;
    const int64_t* masked_in; const int64_t* labels_in;     // prompt-only mode: the [B, N] outputs of an earlier mask-only call
    int64_t* masked_out; int64_t* labels_img_out;           // mask-only mode outputs [B, N]
    int64_t* input_ids; int64_t* labels; int64_t* attn_ones;   // [B, L], [B, L], [B, L + 1]
    showo_seq_mask_t* descs; float* mask_prob_out;
};

__global__ void __launch_bounds__(256) t2i_train_prep_kernel(T2iPrepArgs a) {
    extern __shared__ float prep_smem[];
    float* rnd = prep_smem;                                  // [N]
    uint8_t* msk = reinterpret_cast<uint8_t*>(rnd + a.N);    // [N]
    const int b = blockIdx.x, tid = threadIdx.x, N = a.N;
    const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    if (a.do_mask) {
        float t = a.timesteps ? a.timesteps[b] : u01(philox4x32_10(make_uint4((uint32_t)b, 0u, 0u, 0x21u), key).x);
        float mp;
        if (a.schedule == 0) mp = cosf(t * (float)(3.14159265358979323846 * 0.5));
        else if (a.schedule == 1) mp = fminf(fmaxf(1.f - t, 1e-6f), 1.f);
        else if (a.schedule == 2) mp = fminf(fmaxf(1.f - powf(t, a.schedule_param), 1e-6f), 1.f);
        else mp = t;
        mp = fmaxf(mp, a.min_masking_rate);
        const int n = (int)fmaxf(rintf((float)N * mp), 1.f);
        if (tid == 0 && a.mask_prob_out) a.mask_prob_out[b] = mp;
        for (int i = tid; i < N; i += 256) {
            rnd[i] = a.rand ? a.rand[(int64_t)b * N + i]
                            : u01(philox4x32_10(make_uint4((uint32_t)i, (uint32_t)b, 0u, 0x22u), key).x);
            msk[i] = 0;
        }
        __syncthreads();
        const int nm = n < N ? n : N;
        for (int i = tid; i < nm; i += 256) {
            const float v = rnd[i];
            int rank = 0;
            for (int k = 0; k < N; ++k) { const float u = rnd[k]; rank += (u < v || (u == v && k < i)) ? 1 : 0; }
            msk[rank] = 1;
        }
        __syncthreads();
    }
    const int64_t* codes = a.image_tokens + (int64_t)b * N;
    if (!a.do_prompt) {
        for (int i = tid; i < N; i += 256) {
            a.masked_out[(int64_t)b * N + i] = msk[i] ? a.mask_id : codes[i];
            a.labels_img_out[(int64_t)b * N + i] = msk[i] ? codes[i] : a.ignore_id;
        }
        return;
    }
    // ---- text part: [t2i] (+bos unless the caption already starts with it) text [eos]; dropped caption -> [t2i][bos][eos]
    const int P = a.P, L = P + N + 2;
    const int64_t* txt = a.text_ids + (int64_t)b * a.text_stride;
    int tl = a.text_len[b];
    const float dp = a.drop_probs ? a.drop_probs[b] : u01(philox4x32_10(make_uint4((uint32_t)b, 0u, 0u, 0x23u), key).x);
    const bool drop = dp < a.cond_dropout_prob;
    const bool add_bos = drop || tl == 0 || txt[0] != a.bos_id;
    if (drop) tl = 0;
    int len = 1 + (add_bos ? 1 : 0) + tl + 1;
    const bool trunc = len > P;                              // keep the first P - 1 ids, then eos
    const int pad = trunc ? 0 : P - len;
    int64_t* ids = a.input_ids + (int64_t)b * L;
    int64_t* lab = a.labels + (int64_t)b * L;
    for (int j = tid; j < P; j += 256) {
        int64_t v;
        if (j < pad) v = a.pad_id;
        else {
            const int k = j - pad;
            if (trunc && k == P - 1) v = a.eos_id;
            else if (k == 0) v = a.task_id;
            else if (add_bos && k == 1) v = a.bos_id;
            else {
                const int ti = k - 1 - (add_bos ? 1 : 0);
                v = ti < tl ? txt[ti] : a.eos_id;
            }
        }
        ids[j] = v;
        lab[j] = v == a.pad_id ? a.ignore_id : v;
    }
    for (int i = tid; i < N; i += 256) {
        int64_t vi, vl;
        if (a.do_mask) { vi = msk[i] ? a.mask_id : codes[i]; vl = msk[i] ? codes[i] : a.ignore_id; }
        else { vi = a.masked_in[(int64_t)b * N + i]; vl = a.labels_in[(int64_t)b * N + i]; }
        ids[P + 1 + i] = vi;
        lab[P + 1 + i] = vl == a.pad_id ? a.ignore_id : vl;
    }
    if (tid == 0) {
        ids[P] = a.soi_id; lab[P] = a.soi_id;
        ids[L - 1] = a.eoi_id; lab[L - 1] = a.eoi_id;
        if (a.descs) a.descs[b] = showo_seq_mask_t{pad, P, L, 0, 0};
    }
    if (a.attn_ones)
        for (int j = tid; j <= L; j += 256) a.attn_ones[(int64_t)b * (L + 1) + j] = 1;
}

}  // namespace showo

using namespace showo;

extern "C" int showo_mask_descriptors(const void* mask_dev, int elem_bytes, int B, int L, int64_t batch_stride_elems,
                                      showo_seq_mask_t* out_host, int32_t* mismatches_host, void* stream) {
    SHOWO_CHECK(mask_dev && out_host && mismatches_host && B > 0 && L > 0, "mask_descriptors: bad arguments");
    SHOWO_CHECK(elem_bytes == 4 || elem_bytes == 1, "mask_descriptors: fp32 (additive) or 1-byte (bool) masks only");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    // result buffer: one per device, grown on demand (a cudaMallocAsync per call made the first calls after every sync slow:
    // the pool hands its memory back at each synchronisation)
    static int* bufs[64] = {};
    static int caps[64] = {};
    int dev = 0;
    SHOWO_CUDA_OK(cudaGetDevice(&dev));
    int*& d = bufs[dev & 63];
    if (caps[dev & 63] < B) {
        SHOWO_CUDA_OK(cudaStreamSynchronize(st));
        if (d) cudaFree(d);
        d = nullptr;
        SHOWO_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&d), (size_t)B * 6 * 4));
        caps[dev & 63] = B;
    }
    if (elem_bytes == 4) mask_descriptor_kernel<float><<<B, 256, 0, st>>>((const float*)mask_dev, batch_stride_elems, L, d);
    else mask_descriptor_kernel<uint8_t><<<B, 256, 0, st>>>((const uint8_t*)mask_dev, batch_stride_elems, L, d);
    note_launch();
    std::vector<int> h((size_t)B * 6);
    SHOWO_CUDA_OK(cudaMemcpyAsync(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost, st));
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) {
        out_host[b] = showo_seq_mask_t{h[b * 6], h[b * 6 + 1], h[b * 6 + 2], h[b * 6 + 3], h[b * 6 + 4]};
        mismatches_host[b] = h[b * 6 + 5];
    }
    return 0;
}


extern "C" int showo_t2i_train_prep(const int64_t* image_tokens_dev, int B, int N, const int64_t* text_ids_dev,
                                    const int32_t* text_len_dev, int64_t text_stride, int max_text_len, const int64_t* special_ids_host,
                                    float min_masking_rate, float cond_dropout_prob, int schedule, float schedule_param,
                                    const float* timesteps_dev, const float* rand_dev, const float* drop_probs_dev, uint64_t seed,
                                    int mode, const int64_t* masked_in_dev, const int64_t* labels_in_dev, int64_t* input_ids_out_dev,
                                    int64_t* labels_out_dev, int64_t* attn_ones_out_dev, showo_seq_mask_t* descs_out_dev,
                                    float* mask_prob_out_dev, void* stream) {
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(B > 0 && N > 0 && N <= 16384 && special_ids_host && input_ids_out_dev && labels_out_dev, "t2i_train_prep: bad arguments");
    SHOWO_CHECK(mode >= 1 && mode <= 3, "t2i_train_prep: mode is a bit set {1: mask, 2: prompt}");
    SHOWO_CHECK(schedule >= 0 && schedule <= 3, "t2i_train_prep: unknown schedule");
    T2iPrepArgs a{};
    a.image_tokens = image_tokens_dev; a.text_ids = text_ids_dev; a.text_len = text_len_dev; a.text_stride = text_stride;
    a.B = B; a.N = N; a.P = max_text_len + 1;
    a.pad_id = special_ids_host[0]; a.bos_id = special_ids_host[1]; a.eos_id = special_ids_host[2]; a.task_id = special_ids_host[3];
    a.soi_id = special_ids_host[4]; a.eoi_id = special_ids_host[5]; a.mask_id = special_ids_host[6]; a.ignore_id = special_ids_host[7];
    a.min_masking_rate = min_masking_rate; a.cond_dropout_prob = cond_dropout_prob; a.schedule = schedule; a.schedule_param = schedule_param;
    a.timesteps = timesteps_dev; a.rand = rand_dev; a.drop_probs = drop_probs_dev; a.seed = seed;
    a.do_mask = mode & 1; a.do_prompt = (mode >> 1) & 1;
    a.masked_in = masked_in_dev; a.labels_in = labels_in_dev;
    a.masked_out = input_ids_out_dev; a.labels_img_out = labels_out_dev;
    a.input_ids = input_ids_out_dev; a.labels = labels_out_dev; a.attn_ones = attn_ones_out_dev; a.descs = descs_out_dev;
    a.mask_prob_out = mask_prob_out_dev;
    if (a.do_mask) SHOWO_CHECK(image_tokens_dev != nullptr, "t2i_train_prep: the mask step needs image_tokens");
    if (a.do_prompt) {
        SHOWO_CHECK(text_ids_dev && text_len_dev && max_text_len >= 3, "t2i_train_prep: the prompt step needs the tokenised captions");
        if (!a.do_mask) SHOWO_CHECK(masked_in_dev && labels_in_dev, "t2i_train_prep: prompt-only mode needs the masked ids and labels");
    }
    const size_t smem = (size_t)N * 5 + 16;
    static PerDeviceOnce once;
    if (once.need()) SHOWO_CUDA_OK(cudaFuncSetAttribute(t2i_train_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 5 + 16));
    t2i_train_prep_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(a);
    SHOWO_CUDA_OK(cudaGetLastError());
    note_launch();
    return 0;
}
