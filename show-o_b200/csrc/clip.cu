// CLIP ViT vision tower (the frozen encoder in front of the w_clip_vit MMU path) on sm_100a.
//
// Reference call sites: models/clip_encoder.py:29-51 (`CLIPVisionTower.forward`: `vision_tower(images, output_hidden_states=True)`,
// `hidden_states[select_layer = -2]`, CLS token dropped), used by inference_mmu.py:100-131 and training/train_w_clip_vit.py:532-537.
// The network itself lives in the third-party dependency `transformers` (requirements.txt pins 4.41.1; `CLIPVisionModel`,
// `openai/clip-vit-large-patch14-336`: 24 layers, hidden 1024, 16 heads of 64, MLP 4096, quick_gelu, 336 / 14 -> 576 patches + CLS);
// its published algorithm is restated in oracle/clip_oracle.py, which is pinned to the live `transformers` model in the tests.
//
//   pixels [B,3,S,S] fp32 -> im2col (bf16, K = 3*P*P padded to a multiple of 8) -> patch GEMM (tcgen05, no bias)
//   -> [CLS | patches] + position embedding -> pre_layrnorm = the fp32 residual stream x
//   per layer: LN1 -> fused q|k|v GEMM (+bias, bf16) -> K / V^T tiles -> omni attention kernel with the all-visible descriptor
//              (tcgen05 for the full 128-row tiles, mma.sync for the ragged tail) -> out_proj GEMM (+bias +residual, fp32, in place)
//              -> LN2 -> fc1 GEMM (+bias, bf16) -> quick_gelu -> fc2 GEMM (+bias +residual)
//   hidden_states[select_layer] without the CLS row -> out [B, T-1, hidden] fp32.
// bf16 operands with fp32 accumulation, fp32 residual stream / LayerNorm / softmax: the backbone's numerics.
#include <set>
#include <string>
#include <vector>

#include "engine_state.h"

extern "C" {
struct clip_engine;
}

namespace showo {

// ------------------------------------------------------------------------------------------------ kernels
// patches[(b*G + gy)*G + gx][c*P*P + ky*P + kx] = bf16(pixels[b][c][gy*P + ky][gx*P + kx]); columns >= 3*P*P are zero
__global__ void __launch_bounds__(256) clip_im2col_kernel(const float* __restrict__ px, bf16* __restrict__ out, int B, int S, int P, int G, int KP) {
    const int K = 3 * P * P;
    const int64_t n = (int64_t)B * G * G * KP;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % KP);
        const int64_t r = i / KP;
        float v = 0.f;
        if (k < K) {
            const int gx = (int)(r % G), gy = (int)((r / G) % G), b = (int)(r / ((int64_t)G * G));
            const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
            v = px[(((int64_t)b * 3 + c) * S + gy * P + ky) * S + gx * P + kx];
        }
        out[i] = __float2bfloat16(v);
    }
}

// x[b*T + t] = LayerNorm((t == 0 ? cls : pe[b*(T-1) + t-1]) + pos[t])      (CLIPVisionEmbeddings + pre_layrnorm), one CTA per row
__global__ void __launch_bounds__(256) clip_embed_ln_kernel(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ pos,
                                                            const float* __restrict__ g, const float* __restrict__ bta, float eps,
                                                            float* __restrict__ x, int T, int D) {
    __shared__ float red[16];
    const int row = blockIdx.x, b = row / T, t = row % T;
    const float* src = t == 0 ? cls : pe + ((int64_t)b * (T - 1) + t - 1) * D;
    const float* pr = pos + (int64_t)t * D;
    float v[8];                                              // D <= 2048
    float s = 0.f;
    int n = 0;
    for (int d = threadIdx.x; d < D; d += 256) { v[n] = src[d] + pr[d]; s += v[n]; ++n; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float mean = tot / (float)D;
    float q = 0.f;
    for (int i = 0; i < n; ++i) { const float c = v[i] - mean; q += c * c; }
    q = warp_sum(q);
    if (lane == 0) red[8 + warp] = q;
    __syncthreads();
    float qt = 0.f;
    for (int w = 0; w < 8; ++w) qt += red[8 + w];
    const float rstd = rsqrtf(qt / (float)D + eps);
    float* xr = x + (int64_t)row * D;
    n = 0;
    for (int d = threadIdx.x; d < D; d += 256) { xr[d] = (v[n] - mean) * rstd * g[d] + bta[d]; ++n; }
}

// K tiles [seq][H][Lmax][64] and V^T tiles [seq][H][64][Lmax] of the attention kernels from the fused q|k|v buffer [B*T, 3D];
// one CTA per (64-position chunk, head, sequence)
__global__ void __launch_bounds__(256) clip_kv_scatter_kernel(const bf16* __restrict__ qkv, int64_t ld, int T, int H, int D, int Lmax,
                                                              bf16* __restrict__ kc, bf16* __restrict__ vt) {
    __shared__ bf16 tile[64][66];
    const int p0 = blockIdx.x * 64, h = blockIdx.y, s = blockIdx.z;
    const bf16* base = qkv + ((int64_t)s * T) * ld + h * 64;
    bf16* kdst = kc + (((int64_t)s * H + h) * Lmax) * 64;
    bf16* vdst = vt + (((int64_t)s * H + h) * 64) * Lmax;
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {       // bf16 pairs
        const int p = i >> 5, d2 = (i & 31) * 2;
        const int pos = p0 + p;
        __nv_bfloat162 kv = __floats2bfloat162_rn(0.f, 0.f), vv = kv;
        if (pos < T) {
            kv = *reinterpret_cast<const __nv_bfloat162*>(base + (int64_t)pos * ld + D + d2);
            vv = *reinterpret_cast<const __nv_bfloat162*>(base + (int64_t)pos * ld + 2 * D + d2);
        }
        if (pos < Lmax) *reinterpret_cast<__nv_bfloat162*>(kdst + (int64_t)pos * 64 + d2) = kv;
        tile[p][d2] = vv.x; tile[p][d2 + 1] = vv.y;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int d = i >> 6, p = i & 63;
        if (p0 + p < Lmax) vdst[(int64_t)d * Lmax + p0 + p] = tile[p][d];
    }
}

// quick_gelu (transformers activations: x * sigmoid(1.702 x)), in place on bf16
__global__ void quick_gelu_bf16_kernel(bf16* __restrict__ x, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __bfloat162float(x[i]);
        x[i] = __float2bfloat16(v / (1.f + __expf(-1.702f * v)));
    }
}

}  // namespace showo

using namespace showo;

struct ClipLayer {
    bf16* wqkv = nullptr; float* bqkv = nullptr;     // [3D, D] rows q | k | v
    bf16* wo = nullptr; float* bo = nullptr;
    bf16* w1 = nullptr; float* b1 = nullptr;         // [F, D]
    bf16* w2 = nullptr; float* b2 = nullptr;         // [D, F]
    float *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr;
};

struct clip_engine {
    clip_config_t cfg{};
    int device = 0;
    int D = 0, H = 0, F = 0, NL = 0, S = 0, P = 0, G = 0, T = 0, KP = 0, Lmax = 0;
    bf16* wpatch = nullptr; float* cls = nullptr; float* pos = nullptr; float* pre_g = nullptr; float* pre_b = nullptr;
    std::vector<ClipLayer> layers;
    std::set<std::string> loaded;
    bool finalized = false;
    float* stage = nullptr; size_t stage_cap = 0;
    int cap_B = 0;
    bf16* patches = nullptr; float* pe = nullptr; float* x = nullptr; bf16* xh = nullptr; bf16* qkv = nullptr; bf16* mid = nullptr;
    bf16* kcache = nullptr; bf16* vtcache = nullptr; showo_seq_mask_t* d_masks = nullptr; int* attn_ctr = nullptr;
    int64_t launches_last = 0;
};

static int clip_ensure_ws(clip_engine* e, int B, cudaStream_t st) {
    if (B <= e->cap_B) return 0;
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    dev_free(e->patches); dev_free(e->pe); dev_free(e->x); dev_free(e->xh); dev_free(e->qkv); dev_free(e->mid);
    dev_free(e->kcache); dev_free(e->vtcache); dev_free(e->d_masks);
    const size_t M = (size_t)B * e->T, Mp = (size_t)B * (e->T - 1);
    SHOWO_TRY(dev_alloc(&e->patches, Mp * e->KP));
    SHOWO_TRY(dev_alloc(&e->pe, Mp * e->D));
    SHOWO_TRY(dev_alloc(&e->x, M * e->D));
    SHOWO_TRY(dev_alloc(&e->xh, M * e->D));
    SHOWO_TRY(dev_alloc(&e->qkv, M * 3 * e->D));
    SHOWO_TRY(dev_alloc(&e->mid, M * e->F));
    const size_t kv = (size_t)B * e->H * e->Lmax * 64;
    SHOWO_TRY(dev_alloc(&e->kcache, kv));
    SHOWO_TRY(dev_alloc(&e->vtcache, kv));
    SHOWO_CUDA_OK(cudaMemset(e->kcache, 0, kv * sizeof(bf16)));          // positions in [T, Lmax) stay zero: read by the tile loads, masked by n_keys
    SHOWO_CUDA_OK(cudaMemset(e->vtcache, 0, kv * sizeof(bf16)));
    SHOWO_TRY(dev_alloc(&e->d_masks, (size_t)B));
    std::vector<showo_seq_mask_t> hm((size_t)B);
    for (auto& m : hm) { m.pad_end = 0; m.full_begin = 0; m.full_end = e->T; m.win_begin = 0; m.win_end = 0; }   // every row sees every key
    SHOWO_CUDA_OK(cudaMemcpy(e->d_masks, hm.data(), (size_t)B * sizeof(showo_seq_mask_t), cudaMemcpyHostToDevice));
    if (!e->attn_ctr) {
        SHOWO_TRY(dev_alloc(&e->attn_ctr, (size_t)16));
        SHOWO_CUDA_OK(cudaMemset(e->attn_ctr, 0, 64));
    }
    e->cap_B = B;
    return 0;
}

static int clip_gemm(const bf16* A, int64_t lda, const bf16* Bw, int64_t ldb, int M, int N, int K, void* out, int64_t ldc, const float* bias,
                     const float* resid, GemmEpi epi, cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.B = Bw; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.out = out; g.ldc = ldc; g.bias = bias; g.gelu_from = N;
    g.resid = resid; g.ldr = ldc;
    if (M <= 16) g.block_n = 64;                     // keep tiny batches on the tcgen05 kernel (the M <= 16 route is the decode path's)
    return gemm_bf16(g, epi, st);
}

extern "C" {

int clip_engine_create(const clip_config_t* cfg, int device, clip_engine_t** out) {
    SHOWO_CHECK(cfg && out, "null argument");
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(cfg->n_heads > 0 && cfg->hidden == cfg->n_heads * 64, "clip: head_dim must be 64 (hidden = n_heads * 64)");
    SHOWO_CHECK(cfg->hidden % 128 == 0 && cfg->hidden <= 2048 && cfg->ffn % 8 == 0, "clip: hidden must be a multiple of 128 (<= 2048)");
    SHOWO_CHECK(cfg->patch_size > 0 && cfg->image_size % cfg->patch_size == 0 && cfg->n_layers > 0, "clip: bad geometry");
    SHOWO_CUDA_OK(cudaSetDevice(device));
    clip_engine* e = new clip_engine();
    e->cfg = *cfg; e->device = device;
    e->D = cfg->hidden; e->H = cfg->n_heads; e->F = cfg->ffn; e->NL = cfg->n_layers; e->S = cfg->image_size; e->P = cfg->patch_size;
    e->G = e->S / e->P; e->T = e->G * e->G + 1; e->KP = (3 * e->P * e->P + 7) / 8 * 8; e->Lmax = (e->T + 63) / 64 * 64;
    SHOWO_CHECK(e->T <= 2048, "clip: more than 2048 tokens per image");
    const size_t D = (size_t)e->D, F = (size_t)e->F;
    int rc = dev_alloc(&e->wpatch, D * e->KP);
    if (!rc) rc = cudaMemset(e->wpatch, 0, D * e->KP * sizeof(bf16)) == cudaSuccess ? 0 : -1;
    if (!rc) rc = dev_alloc(&e->cls, D);
    if (!rc) rc = dev_alloc(&e->pos, (size_t)e->T * D);
    if (!rc) rc = dev_alloc(&e->pre_g, D);
    if (!rc) rc = dev_alloc(&e->pre_b, D);
    e->layers.resize((size_t)e->NL);
    for (auto& l : e->layers) {
        if (rc) break;
        rc = dev_alloc(&l.wqkv, 3 * D * D);
        if (!rc) rc = dev_alloc(&l.bqkv, 3 * D);
        if (!rc) rc = dev_alloc(&l.wo, D * D);
        if (!rc) rc = dev_alloc(&l.bo, D);
        if (!rc) rc = dev_alloc(&l.w1, F * D);
        if (!rc) rc = dev_alloc(&l.b1, F);
        if (!rc) rc = dev_alloc(&l.w2, D * F);
        if (!rc) rc = dev_alloc(&l.b2, D);
        if (!rc) rc = dev_alloc(&l.ln1g, D);
        if (!rc) rc = dev_alloc(&l.ln1b, D);
        if (!rc) rc = dev_alloc(&l.ln2g, D);
        if (!rc) rc = dev_alloc(&l.ln2b, D);
    }
    if (rc) { clip_engine_destroy(e); return rc; }
    *out = e;
    return 0;
}

int clip_engine_destroy(clip_engine_t* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    dev_free(e->wpatch); dev_free(e->cls); dev_free(e->pos); dev_free(e->pre_g); dev_free(e->pre_b);
    for (auto& l : e->layers) {
        dev_free(l.wqkv); dev_free(l.bqkv); dev_free(l.wo); dev_free(l.bo); dev_free(l.w1); dev_free(l.b1); dev_free(l.w2); dev_free(l.b2);
        dev_free(l.ln1g); dev_free(l.ln1b); dev_free(l.ln2g); dev_free(l.ln2b);
    }
    dev_free(e->stage); dev_free(e->patches); dev_free(e->pe); dev_free(e->x); dev_free(e->xh); dev_free(e->qkv); dev_free(e->mid);
    dev_free(e->kcache); dev_free(e->vtcache); dev_free(e->d_masks); dev_free(e->attn_ctr);
    delete e;
    return 0;
}

// one fp32 tensor of CLIPVisionModel.state_dict() ("vision_model.encoder.layers.3.self_attn.q_proj.weight", ...)
int clip_load_weight(clip_engine_t* e, const char* name_c, const float* data, int64_t numel, int is_device) {
    SHOWO_CHECK(e && name_c && data, "null argument");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    std::string name(name_c);
    const std::string vp = "vision_model.";
    SHOWO_CHECK(name.compare(0, vp.size(), vp) == 0, "clip: unknown weight name " + name);
    const std::string k = name.substr(vp.size());
    const int64_t D = e->D, F = e->F;
    const float* src = data;
    if (!is_device) {
        if ((size_t)numel > e->stage_cap) {
            dev_free(e->stage);
            SHOWO_TRY(dev_alloc(&e->stage, (size_t)numel));
            e->stage_cap = (size_t)numel;
        }
        SHOWO_CUDA_OK(cudaMemcpy(e->stage, data, (size_t)numel * 4, cudaMemcpyHostToDevice));
        src = e->stage;
    }
    cudaStream_t st = 0;
    auto expect = [&](int64_t n) -> int {
        SHOWO_CHECK(numel == n, "clip weight " + name + ": expected " + std::to_string(n) + " elements, got " + std::to_string(numel));
        return 0;
    };
    auto copy_f32 = [&](float* dst, int64_t n) -> int {
        SHOWO_TRY(expect(n));
        SHOWO_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    };
    auto mat = [&](bf16* dst, int64_t dst_ld, int64_t rows, int64_t cols) -> int {
        SHOWO_TRY(expect(rows * cols));
        return pack_block_bf16(src, cols, dst, dst_ld, (int)rows, (int)cols, st);
    };
    int rc = -2;
    const std::string lp = "encoder.layers.";
    if (k == "embeddings.class_embedding") rc = copy_f32(e->cls, D);
    else if (k == "embeddings.patch_embedding.weight") rc = mat(e->wpatch, e->KP, D, 3 * (int64_t)e->P * e->P);     // [D, 3, P, P] flattened = the im2col column order
    else if (k == "embeddings.position_embedding.weight") rc = copy_f32(e->pos, (int64_t)e->T * D);
    else if (k == "embeddings.position_ids") rc = 0;                                                             // a buffer of older checkpoints
    else if (k == "pre_layrnorm.weight") rc = copy_f32(e->pre_g, D);
    else if (k == "pre_layrnorm.bias") rc = copy_f32(e->pre_b, D);
    else if (k == "post_layernorm.weight" || k == "post_layernorm.bias") rc = expect(D);                         // only feeds pooler_output, which the tower never reads
    else if (k.compare(0, lp.size(), lp) == 0) {
        const size_t dot = k.find('.', lp.size());
        SHOWO_CHECK(dot != std::string::npos, "clip: bad weight name " + name);
        const int l = atoi(k.substr(lp.size(), dot - lp.size()).c_str());
        SHOWO_CHECK(l >= 0 && l < e->NL, "clip: layer index out of range in " + name);
        ClipLayer& w = e->layers[(size_t)l];
        const std::string t = k.substr(dot + 1);
        if (t == "self_attn.q_proj.weight") rc = mat(w.wqkv, D, D, D);
        else if (t == "self_attn.k_proj.weight") rc = mat(w.wqkv + D * D, D, D, D);
        else if (t == "self_attn.v_proj.weight") rc = mat(w.wqkv + 2 * D * D, D, D, D);
        else if (t == "self_attn.q_proj.bias") rc = copy_f32(w.bqkv, D);
        else if (t == "self_attn.k_proj.bias") rc = copy_f32(w.bqkv + D, D);
        else if (t == "self_attn.v_proj.bias") rc = copy_f32(w.bqkv + 2 * D, D);
        else if (t == "self_attn.out_proj.weight") rc = mat(w.wo, D, D, D);
        else if (t == "self_attn.out_proj.bias") rc = copy_f32(w.bo, D);
        else if (t == "mlp.fc1.weight") rc = mat(w.w1, D, F, D);
        else if (t == "mlp.fc1.bias") rc = copy_f32(w.b1, F);
        else if (t == "mlp.fc2.weight") rc = mat(w.w2, F, D, F);
        else if (t == "mlp.fc2.bias") rc = copy_f32(w.b2, D);
        else if (t == "layer_norm1.weight") rc = copy_f32(w.ln1g, D);
        else if (t == "layer_norm1.bias") rc = copy_f32(w.ln1b, D);
        else if (t == "layer_norm2.weight") rc = copy_f32(w.ln2g, D);
        else if (t == "layer_norm2.bias") rc = copy_f32(w.ln2b, D);
        else { set_last_error("clip: unknown weight name " + name); rc = -2; }
    } else {
        set_last_error("clip: unknown weight name " + name);
    }
    if (rc) return rc;
    SHOWO_CUDA_OK(cudaStreamSynchronize(st));
    e->loaded.insert(name);
    e->finalized = false;
    return 0;
}

int clip_weights_complete(clip_engine_t* e) {
    SHOWO_CHECK(e, "null engine");
    std::vector<std::string> need = {"embeddings.class_embedding", "embeddings.patch_embedding.weight", "embeddings.position_embedding.weight",
                                     "pre_layrnorm.weight", "pre_layrnorm.bias"};
    const char* per_layer[] = {"self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight", "self_attn.k_proj.bias",
                               "self_attn.v_proj.weight", "self_attn.v_proj.bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                               "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "layer_norm1.weight", "layer_norm1.bias",
                               "layer_norm2.weight", "layer_norm2.bias"};
    for (int l = 0; l < e->NL; ++l)
        for (const char* p : per_layer) need.push_back("encoder.layers." + std::to_string(l) + "." + p);
    for (const auto& n : need) SHOWO_CHECK(e->loaded.count("vision_model." + n) == 1, "clip: weight not loaded: vision_model." + n);
    e->finalized = true;
    return 0;
}

// CLIPVisionTower.forward (models/clip_encoder.py:39-51): pixels_dev [B, 3, S, S] fp32 (already normalised by the image processor)
// -> hidden_states[select_layer] (negative: from the end, -2 = the penultimate block's output, the reference's choice; 0 = the
// embeddings after pre_layrnorm), CLS row dropped when drop_cls != 0 ('patch'), kept otherwise ('cls_patch').
// out_dev fp32 [B, T - 1 or T, hidden].
int clip_forward(clip_engine_t* e, const float* pixels_dev, int B, int select_layer, int drop_cls, float* out_dev, void* stream) {
    SHOWO_CHECK(e && pixels_dev && out_dev && B > 0, "clip_forward: bad arguments");
    SHOWO_CUDA_OK(cudaSetDevice(e->device));
    SHOWO_CHECK(showo_device_count() > 0, "no sm_100 CUDA device visible: libshowo_b200 has no CPU fallback");
    SHOWO_CHECK(e->finalized, "clip_forward: call clip_weights_complete first");
    const int n_run = select_layer >= 0 ? select_layer : e->NL + 1 + select_layer;
    SHOWO_CHECK(n_run >= 0 && n_run <= e->NL, "clip_forward: select_layer out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t l0 = launches_total();
    SHOWO_TRY(clip_ensure_ws(e, B, st));
    const int D = e->D, F = e->F, T = e->T, H = e->H, M = B * T, Mp = B * (T - 1);
    {
        const int64_t n = (int64_t)Mp * e->KP;
        const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
        clip_im2col_kernel<<<grid, 256, 0, st>>>(pixels_dev, e->patches, B, e->S, e->P, e->G, e->KP);
        note_launch();
    }
    SHOWO_TRY(clip_gemm(e->patches, e->KP, e->wpatch, e->KP, Mp, D, e->KP, e->pe, D, nullptr, nullptr, GEMM_BIAS_F32, st));
    clip_embed_ln_kernel<<<M, 256, 0, st>>>(e->pe, e->cls, e->pos, e->pre_g, e->pre_b, e->cfg.ln_eps, e->x, T, D);
    note_launch();
    SHOWO_CUDA_OK(cudaGetLastError());
    for (int l = 0; l < n_run; ++l) {
        const ClipLayer& w = e->layers[(size_t)l];
        SHOWO_TRY(layernorm_bf16(e->x, w.ln1g, w.ln1b, e->cfg.ln_eps, e->xh, M, D, M, M, 0, st));
        SHOWO_TRY(clip_gemm(e->xh, D, w.wqkv, D, M, 3 * D, D, e->qkv, 3 * D, w.bqkv, nullptr, GEMM_BIAS_BF16, st));
        clip_kv_scatter_kernel<<<dim3(e->Lmax / 64, H, B), 256, 0, st>>>(e->qkv, 3 * D, T, H, D, e->Lmax, e->kcache, e->vtcache);
        note_launch();
        SHOWO_CUDA_OK(cudaGetLastError());
        AttnArgs a{};
        a.q = e->qkv; a.ld = 3 * D; a.n_seq = B; a.H = H; a.rows_per_seq = T; a.pos0 = 0;
        a.kcache = e->kcache; a.vtcache = e->vtcache; a.Lmax = e->Lmax; a.n_keys = T; a.masks = e->d_masks; a.scale = 0.125f;   // head_dim^-0.5
        a.work_ctr = e->attn_ctr;
        SHOWO_TRY(omni_attention(a, st));                    // the output overwrites the q columns
        SHOWO_TRY(clip_gemm(e->qkv, 3 * D, w.wo, D, M, D, D, e->x, D, w.bo, e->x, GEMM_RESID_F32, st));
        SHOWO_TRY(layernorm_bf16(e->x, w.ln2g, w.ln2b, e->cfg.ln_eps, e->xh, M, D, M, M, 0, st));
        SHOWO_TRY(clip_gemm(e->xh, D, w.w1, D, M, F, D, e->mid, F, w.b1, nullptr, GEMM_BIAS_BF16, st));
        {
            const int64_t n = (int64_t)M * F;
            const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
            quick_gelu_bf16_kernel<<<grid, 256, 0, st>>>(e->mid, n);
            note_launch();
        }
        SHOWO_TRY(clip_gemm(e->mid, F, w.w2, F, M, D, F, e->x, D, w.b2, e->x, GEMM_RESID_F32, st));
    }
    const int skip = drop_cls ? 1 : 0;
    SHOWO_CUDA_OK(cudaMemcpy2DAsync(out_dev, (size_t)(T - skip) * D * 4, e->x + (size_t)skip * D, (size_t)T * D * 4, (size_t)(T - skip) * D * 4,
                                    (size_t)B, cudaMemcpyDeviceToDevice, st));
    SHOWO_CUDA_OK(cudaGetLastError());
    e->launches_last = launches_total() - l0;
    return 0;
}

int64_t clip_kernel_launches(clip_engine_t* e) { return e ? e->launches_last : 0; }

}  // extern "C"
