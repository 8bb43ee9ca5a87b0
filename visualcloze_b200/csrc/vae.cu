// vae.cu -- VAE decoder (models/modules/autoencoder.py:183-259) as a sequence of libvcb200 launches, NHWC bf16.
#include <algorithm>
#include <vector>

#include "../../include/vcb200.h"
#include "host_util.cuh"
#include "vae_kernels.cuh"

using namespace vcb;

struct vcb_vae {
    vcb_vae_config cfg;
    vcb_vae_weights w;
    std::vector<vcb_resblock_w> up;
    std::vector<vcb_conv_w> ups;
};

namespace {

constexpr int kZPad = 64;     // latent channels padded to one 64-channel k-block

struct Ws {
    uint16_t *a, *b, *c, *d;          // activation ping-pong buffers (max activation each)
    float* gn_part; float2* gn_stats;
    float* scores; uint16_t* probs;   // mid attention
    uint16_t* vT;
    int64_t total;
};

int64_t max_act_elems(const vcb_vae_config& c, int n, int lh, int lw) {
    // largest NHWC activation: track (H, W, C) through the decoder
    int64_t best = (int64_t)n * lh * lw * kZPad;
    int H = lh, W = lw;
    int C = c.ch * c.ch_mult[c.n_levels - 1];
    best = std::max<int64_t>(best, (int64_t)n * H * W * C);
    for (int lvl = c.n_levels - 1; lvl >= 0; --lvl) {
        const int co = c.ch * c.ch_mult[lvl];
        best = std::max<int64_t>(best, (int64_t)n * H * W * std::max(C, co));
        C = co;
        if (lvl != 0) { H *= 2; W *= 2; best = std::max<int64_t>(best, (int64_t)n * H * W * C); }
    }
    return best;
}

Ws carve(const vcb_vae_config& c, uint8_t* base, int n, int lh, int lw) {
    auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
    int64_t off = 0;
    auto take = [&](int64_t bytes) { uint8_t* p = base ? base + off : nullptr; off += al(bytes); return p; };
    Ws w{};
    // largest activation (the encoder mirrors the decoder's resolutions/widths; its 64-channel padded input image is
    // 16 H W * 64 <= 16 H W * ch elements); mid-block attention needs q,k = 2 P C in one buffer
    int64_t act_e = max_act_elems(c, n, lh, lw);
    int64_t full = (int64_t)n * lh * lw;
    for (int i = 1; i < c.n_levels; ++i) full *= 4;
    act_e = std::max<int64_t>(act_e, full * std::max(64, c.ch));
    act_e = std::max<int64_t>(act_e, 2 * (int64_t)lh * lw * c.ch * c.ch_mult[c.n_levels - 1]);
    const int64_t act = act_e * 2;
    w.a = (uint16_t*)take(act); w.b = (uint16_t*)take(act); w.c = (uint16_t*)take(act); w.d = (uint16_t*)take(act);
    // GroupNorm partials: the largest pixel count is the final resolution
    int64_t maxP = (int64_t)lh * lw;
    for (int i = 1; i < c.n_levels; ++i) maxP *= 4;
    const int64_t chunks = (maxP + kGnPixelsPerBlock - 1) / kGnPixelsPerBlock;
    w.gn_part = (float*)take((int64_t)n * chunks * kGnGroups * 2 * 4);
    w.gn_stats = (float2*)take((int64_t)n * kGnGroups * 8);
    const int64_t P = (int64_t)lh * lw;
    const int64_t Ppad = (P + 7) / 8 * 8;
    w.scores = (float*)take(P * Ppad * 4);
    w.probs = (uint16_t*)take(P * Ppad * 2);
    const int Cmid = c.ch * c.ch_mult[c.n_levels - 1];
    w.vT = (uint16_t*)take((int64_t)Cmid * Ppad * 2);
    w.total = off;
    return w;
}

int group_norm(const Ws& ws, const uint16_t* x, uint16_t* y, const vcb_gn_w& g, int n, int64_t P, int C, bool swish, cudaStream_t st) {
    if (C % 32 || C > 512 || C % 8) return set_error("vae: GroupNorm needs C %% 32 == 0 and C <= 512 (got %d)", C);
    const int chunks = (int)((P + kGnPixelsPerBlock - 1) / kGnPixelsPerBlock);
    {
        ProfScope ps(PROF_VAE_EW, st);
        gn_partial_kernel<<<dim3(chunks, n), kGnThreads, 0, st>>>((const __nv_bfloat16*)x, ws.gn_part, (int)P, C);
        if (int rc = check_launch("gn_partial")) return rc;
        gn_finalize_kernel<<<dim3(kGnGroups, n), 32, 0, st>>>(ws.gn_part, ws.gn_stats, chunks, P * (C / kGnGroups));
        if (int rc = check_launch("gn_finalize")) return rc;
        const int64_t work = P * (C / 8);
        gn_apply_kernel<<<dim3((unsigned)((work + 255) / 256), n), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, ws.gn_stats,
                                                                                g.gamma, g.beta, P, C, swish ? 1 : 0);
        if (int rc = check_launch("gn_apply")) return rc;
    }
    return 0;
}

int conv3(const uint16_t* x, const vcb_conv_w& cw, const uint16_t* res, uint16_t* out, int n, int H, int W, void* st) {
    return vcb_conv3x3_nhwc(x, cw.w, cw.b, res, out, n, H, W, cw.cin, cw.cout, 1, st);
}

// 1x1 conv == GEMM over pixels; optional residual
int conv1(const uint16_t* x, const vcb_conv_w& cw, const uint16_t* res, uint16_t* out, int64_t pixels, void* st) {
    vcb_gemm_args g{};
    g.M = (int32_t)pixels; g.N = cw.cout; g.K = cw.cin;
    g.A = x; g.lda = cw.cin; g.W = cw.w; g.ldw = cw.cin; g.bias = cw.b;
    g.out = out; g.ldo = cw.cout; g.rows_per_batch = g.M; g.out_batch_rows = g.M;
    g.epilogue = res ? VCB_EPI_GATE_RES : VCB_EPI_BIAS;
    g.res = res; g.ld_res = cw.cout;
    return vcb_gemm_bf16(&g, st);
}

// ResnetBlock: out = shortcut(x) + conv2(gn_swish(conv1(gn_swish(x))))   (autoencoder.py:68-82)
int resblock(const Ws& ws, const vcb_resblock_w& rb, uint16_t* x, uint16_t* t1, uint16_t* t2, uint16_t* out, int n, int H, int W,
             cudaStream_t st) {
    const int64_t P = (int64_t)H * W;
    const int cin = rb.conv1.cin, cout = rb.conv1.cout;
    int rc;
    if ((rc = group_norm(ws, x, t1, rb.norm1, n, P, cin, true, st))) return rc;
    if ((rc = conv3(t1, rb.conv1, nullptr, t2, n, H, W, st))) return rc;
    if ((rc = group_norm(ws, t2, t1, rb.norm2, n, P, cout, true, st))) return rc;
    const uint16_t* skip = x;
    if (rb.shortcut.w) {
        if ((rc = conv1(x, rb.shortcut, nullptr, t2, n * P, st))) return rc;      // t2 is free again after norm2
        skip = t2;
    }
    return conv3(t1, rb.conv2, skip, out, n, H, W, st);
}

// AttnBlock (autoencoder.py:25-52): x + proj_out(softmax(q k^T / sqrt(C)) v), single head over the P = H*W pixels.
int mid_attention(const Ws& ws, const vcb_gn_w& norm, const vcb_conv_w& wq, const vcb_conv_w& wk, const vcb_conv_w& wv,
                  const vcb_conv_w& wproj, const uint16_t* x, uint16_t* t1, uint16_t* t2, uint16_t* y, int n, int H, int W,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    const int C = wq.cin;
    const int64_t P = (int64_t)H * W, Ppad = (P + 7) / 8 * 8;
    const float scale = 1.0f / sqrtf((float)C);
    for (int b = 0; b < n; ++b) {               // attention is per image
        const uint16_t* xb = x + (int64_t)b * P * C;
        uint16_t* hb = t1;                       // normed input [P, C]
        if ((rc = group_norm(ws, xb, hb, norm, 1, P, C, false, st))) return rc;
        uint16_t* q = t2;                        // [P, C]
        uint16_t* k = t2 + P * C;                // [P, C]  (activation buffers hold >= 2*P*C elements, see carve)
        if ((rc = conv1(hb, wq, nullptr, q, P, stream))) return rc;
        if ((rc = conv1(hb, wk, nullptr, k, P, stream))) return rc;
        // vT [C, P] = Wv [C, C] x h^T  (bias of v folded into the PV epilogue: softmax rows sum to 1)
        {
            vcb_gemm_args g{};
            g.M = C; g.N = (int32_t)P; g.K = C;
            g.A = wv.w; g.lda = C; g.W = hb; g.ldw = C; g.bias = nullptr;
            g.out = ws.vT; g.ldo = Ppad; g.rows_per_batch = C; g.out_batch_rows = C; g.epilogue = VCB_EPI_BIAS;
            if ((rc = vcb_gemm_bf16(&g, stream))) return rc;
        }
        {   // scores fp32 [P, P] = q k^T
            vcb_gemm_args g{};
            g.M = (int32_t)P; g.N = (int32_t)P; g.K = C;
            g.A = q; g.lda = C; g.W = k; g.ldw = C; g.bias = nullptr;
            g.out = ws.scores; g.ldo = Ppad; g.rows_per_batch = (int32_t)P; g.out_batch_rows = (int32_t)P;
            g.epilogue = VCB_EPI_BIAS_F32;
            if ((rc = vcb_gemm_bf16(&g, stream))) return rc;
        }
        {
            ProfScope ps(PROF_VAE_EW, stream);
            softmax_rows_kernel<<<(unsigned)P, 256, 0, st>>>(ws.scores, (__nv_bfloat16*)ws.probs, (int)P, Ppad, Ppad, scale);
            if ((rc = check_launch("softmax_rows"))) return rc;
        }
        {   // o [P, C] = probs [P, P] x vT^T + b_v
            vcb_gemm_args g{};
            g.M = (int32_t)P; g.N = C; g.K = (int32_t)P;
            g.A = ws.probs; g.lda = Ppad; g.W = ws.vT; g.ldw = Ppad; g.bias = wv.b;
            g.out = q; g.ldo = C; g.rows_per_batch = (int32_t)P; g.out_batch_rows = (int32_t)P; g.epilogue = VCB_EPI_BIAS;
            if ((rc = vcb_gemm_bf16(&g, stream))) return rc;
        }
        if ((rc = conv1(q, wproj, xb, y + (int64_t)b * P * C, P, stream))) return rc;   // x + proj_out(o)
    }
    return 0;
}

}  // namespace

extern "C" int vcb_vae_create(const vcb_vae_config* cfg, const vcb_vae_weights* w, vcb_vae** out) {
    if (!cfg || !w || !out) return set_error("vae_create: null argument");
    if (cfg->n_levels < 1 || cfg->n_levels > 8 || cfg->num_res_blocks < 0) return set_error("vae_create: bad config");
    if (4 * cfg->z_channels % 8 || cfg->z_channels > kZPad) return set_error("vae_create: z_channels must be <= 64 and even");
    for (int i = 0; i < cfg->n_levels; ++i)
        if ((cfg->ch * cfg->ch_mult[i]) % 64) return set_error("vae_create: every level width must be a multiple of 64");
    vcb_vae* v = new vcb_vae();
    v->cfg = *cfg;
    v->w = *w;
    v->up.assign(w->up_blocks, w->up_blocks + cfg->n_levels * (cfg->num_res_blocks + 1));
    if (cfg->n_levels > 1) v->ups.assign(w->upsample, w->upsample + cfg->n_levels - 1);
    v->w.up_blocks = v->up.data();
    v->w.upsample = v->ups.data();
    *out = v;
    return 0;
}
extern "C" void vcb_vae_destroy(vcb_vae* v) { delete v; }

extern "C" int64_t vcb_vae_workspace_bytes(const vcb_vae* v, int32_t n, int32_t h, int32_t w) {
    if (!v || n <= 0 || h <= 0 || w <= 0) return -1;
    return carve(v->cfg, nullptr, n, 2 * h, 2 * w).total;
}

extern "C" int vcb_vae_decode(vcb_vae* v, void* workspace, int64_t workspace_bytes, const void* tokens, int32_t n, int32_t h,
                              int32_t w, float* raw, uint8_t* img, void* stream) {
    if (!v || !workspace || !tokens || (!raw && !img) || n <= 0 || h <= 0 || w <= 0) return set_error("vae_decode: bad arguments");
    ProfContext in_vae(1);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return set_error("vae_decode: workspace must be 256-byte aligned");
    if (int rc = ensure_device()) return rc;
    const vcb_vae_config& c = v->cfg;
    int H = 2 * h, W = 2 * w;                       // latent resolution
    const Ws ws = carve(c, static_cast<uint8_t*>(workspace), n, H, W);
    if (workspace_bytes < ws.total) return set_error("vae_decode: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)ws.total);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    // latent tokens -> NHWC (64-channel padded) with z / scale + shift
    {
        ProfScope ps(PROF_VAE_EW, stream);
        const int64_t per = (int64_t)H * W * kZPad;
        tokens_to_nhwc_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>((const __nv_bfloat16*)tokens, (__nv_bfloat16*)ws.a, h, w,
                                                                                      c.z_channels, kZPad, c.scale_factor, c.shift_factor);
        if ((rc = check_launch("tokens_to_nhwc"))) return rc;
    }
    uint16_t *x = ws.b, *t1 = ws.c, *t2 = ws.d, *y = ws.a;
    if ((rc = conv3(ws.a, v->w.conv_in, nullptr, x, n, H, W, stream))) return rc;
    // ---- middle: ResnetBlock, AttnBlock, ResnetBlock (autoencoder.py:242-244) ----
    if ((rc = resblock(ws, v->w.mid1, x, t1, t2, y, n, H, W, st))) return rc;
    std::swap(x, y);
    if ((rc = mid_attention(ws, v->w.attn_norm, v->w.attn_q, v->w.attn_k, v->w.attn_v, v->w.attn_proj, x, t1, t2, y, n, H, W, stream))) return rc;
    std::swap(x, y);
    if ((rc = resblock(ws, v->w.mid2, x, t1, t2, y, n, H, W, st))) return rc;
    std::swap(x, y);
    // ---- upsampling path (autoencoder.py:247-253) ----
    int bi = 0, ui = 0;
    for (int lvl = c.n_levels - 1; lvl >= 0; --lvl) {
        for (int r = 0; r < c.num_res_blocks + 1; ++r) {
            if ((rc = resblock(ws, v->up[bi++], x, t1, t2, y, n, H, W, st))) return rc;
            std::swap(x, y);
        }
        if (lvl != 0) {
            const vcb_conv_w& uw = v->ups[ui++];
            {
                ProfScope ps(PROF_VAE_EW, stream);
                const int64_t per = (int64_t)(2 * H) * (2 * W) * (uw.cin / 8);
                upsample2x_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)t1, H, W, uw.cin);
                if ((rc = check_launch("upsample2x"))) return rc;
            }
            H *= 2; W *= 2;
            if ((rc = conv3(t1, uw, nullptr, y, n, H, W, stream))) return rc;
            std::swap(x, y);
        }
    }
    // ---- norm_out, swish, conv_out (autoencoder.py:256-258) ----
    const int Cl = v->w.conv_out.cin;
    if ((rc = group_norm(ws, x, t1, v->w.norm_out, n, (int64_t)H * W, Cl, true, st))) return rc;
    if ((rc = conv3(t1, v->w.conv_out, nullptr, y, n, H, W, stream))) return rc;
    {
        ProfScope ps(PROF_VAE_EW, stream);
        const int64_t per = (int64_t)c.out_ch * H * W;
        nhwc_to_image_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>((const __nv_bfloat16*)y, raw, img, H, W,
                                                                                     v->w.conv_out.cout, c.out_ch);
        if ((rc = check_launch("nhwc_to_image"))) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// encoder (autoencoder.py:109-180): conv_in -> [ResnetBlock x R, Downsample]* -> mid -> GN+swish -> conv_out (moments)
// ------------------------------------------------------------------------------------------------
struct vcb_vae_enc {
    vcb_vae_config cfg;
    vcb_vae_enc_weights w;
    std::vector<vcb_resblock_w> down;
    std::vector<vcb_conv_w> dsc;
};

extern "C" int vcb_vae_enc_create(const vcb_vae_config* cfg, const vcb_vae_enc_weights* w, vcb_vae_enc** out) {
    if (!cfg || !w || !out) return set_error("vae_enc_create: null argument");
    if (cfg->n_levels < 1 || cfg->n_levels > 8 || cfg->num_res_blocks < 1) return set_error("vae_enc_create: bad config");
    for (int i = 0; i < cfg->n_levels; ++i)
        if ((cfg->ch * cfg->ch_mult[i]) % 64) return set_error("vae_enc_create: every level width must be a multiple of 64");
    if (cfg->ch % 64 || (2 * cfg->z_channels) % 8) return set_error("vae_enc_create: ch %% 64 and 2 z %% 8 required");
    vcb_vae_enc* e = new vcb_vae_enc();
    e->cfg = *cfg;
    e->w = *w;
    e->down.assign(w->down_blocks, w->down_blocks + cfg->n_levels * cfg->num_res_blocks);
    if (cfg->n_levels > 1) e->dsc.assign(w->downsample, w->downsample + cfg->n_levels - 1);
    e->w.down_blocks = e->down.data();
    e->w.downsample = e->dsc.data();
    *out = e;
    return 0;
}
extern "C" void vcb_vae_enc_destroy(vcb_vae_enc* e) { delete e; }

extern "C" int64_t vcb_vae_enc_workspace_bytes(const vcb_vae_enc* e, int32_t n, int32_t H, int32_t W) {
    if (!e || n <= 0 || H <= 0 || W <= 0) return -1;
    const int f = 1 << (e->cfg.n_levels - 1);
    if (H % (2 * f) || W % (2 * f)) return -1;
    return carve(e->cfg, nullptr, n, H / f, W / f).total;
}

extern "C" int vcb_vae_encode(vcb_vae_enc* e, void* workspace, int64_t workspace_bytes, const float* image, int32_t n, int32_t H,
                              int32_t W, const float* noise, void* tokens, float* moments, void* stream) {
    if (!e || !workspace || !image || !tokens || n <= 0) return set_error("vae_encode: bad arguments");
    ProfContext in_vae(1);
    const vcb_vae_config& c = e->cfg;
    const int f = 1 << (c.n_levels - 1);
    if (H <= 0 || W <= 0 || H % (2 * f) || W % (2 * f)) return set_error("vae_encode: H and W must be multiples of %d", 2 * f);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return set_error("vae_encode: workspace must be 256-byte aligned");
    if (int rc = ensure_device()) return rc;
    const Ws ws = carve(c, static_cast<uint8_t*>(workspace), n, H / f, W / f);
    if (workspace_bytes < ws.total) return set_error("vae_encode: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)ws.total);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    {
        ProfScope ps(PROF_VAE_EW, stream);
        const int64_t per = (int64_t)H * W * 64;
        image_to_nhwc_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>(image, (__nv_bfloat16*)ws.a, H, W, 3, 64);
        if ((rc = check_launch("image_to_nhwc"))) return rc;
    }
    uint16_t *x = ws.b, *t1 = ws.c, *t2 = ws.d, *y = ws.a;
    if ((rc = conv3(ws.a, e->w.conv_in, nullptr, x, n, H, W, stream))) return rc;
    int h = H, w = W, bi = 0;
    for (int lvl = 0; lvl < c.n_levels; ++lvl) {
        for (int r = 0; r < c.num_res_blocks; ++r) {
            if ((rc = resblock(ws, e->down[bi++], x, t1, t2, y, n, h, w, st))) return rc;
            std::swap(x, y);
        }
        if (lvl != c.n_levels - 1) {
            const vcb_conv_w& dw = e->dsc[lvl];
            if ((rc = vcb_conv3x3_nhwc(x, dw.w, dw.b, nullptr, y, n, h, w, dw.cin, dw.cout, 2, stream))) return rc;
            std::swap(x, y);
            h /= 2; w /= 2;
        }
    }
    if ((rc = resblock(ws, e->w.mid1, x, t1, t2, y, n, h, w, st))) return rc;
    std::swap(x, y);
    if ((rc = mid_attention(ws, e->w.attn_norm, e->w.attn_q, e->w.attn_k, e->w.attn_v, e->w.attn_proj, x, t1, t2, y, n, h, w, stream))) return rc;
    std::swap(x, y);
    if ((rc = resblock(ws, e->w.mid2, x, t1, t2, y, n, h, w, st))) return rc;
    std::swap(x, y);
    if ((rc = group_norm(ws, x, t1, e->w.norm_out, n, (int64_t)h * w, e->w.conv_out.cin, true, st))) return rc;
    if ((rc = conv3(t1, e->w.conv_out, nullptr, y, n, h, w, stream))) return rc;
    {
        ProfScope ps(PROF_VAE_EW, stream);
        const int64_t per = (int64_t)h * w * c.z_channels;
        moments_to_tokens_kernel<<<dim3((unsigned)((per + 255) / 256), n), 256, 0, st>>>((const __nv_bfloat16*)y, noise, (__nv_bfloat16*)tokens,
                                                                                         moments, h, w, c.z_channels, c.scale_factor, c.shift_factor);
        if ((rc = check_launch("moments_to_tokens"))) return rc;
    }
    return 0;
}
