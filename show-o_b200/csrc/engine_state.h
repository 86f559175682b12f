// Engine state shared by the inference orchestration (engine.cu) and the training step (train.cu).
#pragma once
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace showo {
template <class T>
inline int dev_alloc(T** p, size_t n) {
    *p = nullptr;
    if (n == 0) return 0;
    SHOWO_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
    return 0;
}
template <class T>
inline void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}


int64_t launches_total();
struct TrainState;
void train_state_destroy(TrainState* t);
struct OptState;
void opt_state_destroy(OptState* o);
}  // namespace showo

using showo::bf16;

struct LayerW {
    bf16* w1 = nullptr;   // [3D+F, D]
    float* b1 = nullptr;  // [3D+F]
    bf16* w2 = nullptr;   // [D, D+F]
    float* b2 = nullptr;  // [D]  (= dense.bias + fc2.bias)
    float* b_dense = nullptr; float* b_fc2 = nullptr;
    float* ln_g = nullptr; float* ln_b = nullptr;
    float* qg = nullptr; float* qb = nullptr; float* kg = nullptr; float* kb = nullptr;
    // input LayerNorm folded into the fused projection (engine.cu ln_fold_enabled, layers >= 1): w1f = bf16(W1 * gamma),
    // ln_c[n] = sum_j w1f[n][j], ln_d[n] = b1[n] + sum_j beta[j] W1[n][j]
    bf16* w1f = nullptr; float* ln_c = nullptr; float* ln_d = nullptr;
};

struct showo_engine {
    showo_config_t cfg{};
    int device = 0;
    int D = 0, H = 0, F = 0, NL = 0, V = 0, W1N = 0, W2K = 0;
    bf16* embed = nullptr;
    bf16* head_w = nullptr; float* head_b = nullptr;
    float* head_b_img = nullptr;                         // 16B-aligned copy of head_b[image_offset : image_offset + C]
    float* fln_g = nullptr; float* fln_b = nullptr;
    std::vector<LayerW> layers;
    float* cos_tab = nullptr; float* sin_tab = nullptr;
    std::set<std::string> loaded;
    bool finalized = false;
    int64_t weights_version = 0;                         // bumped by every showo_load_weight (train.cu re-derives its transposed copies)
    float* stage = nullptr; size_t stage_cap = 0;       // fp32 staging for weight uploads
    // workspaces
    int cap_rows = 0, cap_seq = 0, cap_L = 0; int64_t cap_logit_elems = 0;
    float* x = nullptr; bf16* xh = nullptr; bf16* buf = nullptr;
    bf16* w1_slab = nullptr; bf16* w2_slab = nullptr;    // all layers' W1 / W2 back to back (one tensor map each)
    bf16* w1f_slab = nullptr; float* ln_cd = nullptr;    // LayerNorm-folded copies of W1 and their c / d vectors (LayerW::w1f)
    float* ln_part = nullptr;                             // [cap_rows][D / 64][mean, M2]: slot statistics of the residual stream
    bf16* kcache = nullptr; bf16* vtcache = nullptr;     // [NL][cap_seq][H][cap_L][64] each
    showo_seq_mask_t* d_masks = nullptr;
    float* logits_ws = nullptr; float* conf_ws = nullptr; int* sampled_ws = nullptr;
    int64_t* tok_ws = nullptr; int64_t tok_ws_cap = 0;
    unsigned long long* argmax_keys = nullptr;            // [16] packed (logit, ~index) maxima of the fused greedy head
    // Showo.mm_projector (w_clip_vit): Linear(1024, 2048) -> GELU -> Linear(2048, 2048), modeling_showo.py:49-54
    bf16* mmp_w0 = nullptr; float* mmp_b0 = nullptr; bf16* mmp_w2 = nullptr; float* mmp_b2 = nullptr;
    bf16* mmp_in = nullptr; bf16* mmp_mid = nullptr; int64_t mmp_cap = 0;
    // its training side (train.cu: showo_mm_projector_backward, training/train_w_clip_vit.py:599-601): the pre-GELU activations of the
    // last forward call, fp32 gradients [w0 | b0 | w2 | b2] (6,295,552 elements), a transposed copy of w2 for the dgrad GEMM
    bf16* mmp_pre = nullptr; int64_t mmp_n = 0; int64_t mmp_version = 0;
    float* mmp_grads = nullptr; bool mmp_grads_valid = false;
    bf16* mmp_w2t = nullptr; int64_t mmp_w2t_version = -1;
    bf16* mmp_dy = nullptr; bf16* mmp_dmid = nullptr; int64_t mmp_bwd_cap = 0;
    int* attn_ctr = nullptr;                              // work counter of the attention kernel's tail phase (self-resetting, zero between launches)
    int* finished_ws = nullptr;                           // [64] rows of the running mmu_generate that have produced eot_token
    int rng_row_base = 0;                                 // showo_set_rng_row_base: global index of the first batch row (Philox noise keyed by it)
    int64_t launches_last = 0;
    showo::OptState* opt = nullptr;                      // fp32 master weights + Adam moments (train.cu), showo_optimizer_enable
    showo::TrainState* train = nullptr;                  // training-step buffers (train.cu), allocated on first use
};


// helpers defined in engine.cu
int engine_check_ready(showo_engine* e);
int engine_upload_masks(showo_engine* e, const showo_seq_mask_t* masks_host, int n, cudaStream_t st);
int engine_ensure_ws(showo_engine* e, int rows, int n_seq, int L, int64_t logit_elems, cudaStream_t st);

// train.cu: keep the fp32 value of a reference parameter as the optimizer's master copy (called by showo_load_weight when the
// optimizer is enabled); engine.cu: re-derive b2 = b_dense + b_fc2 and the image slice of the head bias after an optimizer step
namespace showo {
int opt_store_master(showo_engine* e, const std::string& name, const float* src_dev, int64_t numel, cudaStream_t st);
// the fp32 master of one reference parameter inside the optimizer state: pointer + [rows, cols] with row stride ld (verify.cu)
int opt_master_slot(showo_engine* e, const std::string& name, float** ptr, int64_t* rows, int64_t* cols, int64_t* ld);
}
int engine_refresh_derived(showo_engine* e, cudaStream_t st);
