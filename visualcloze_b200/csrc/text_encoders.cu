// text_encoders.cu -- C entry points of the text-encoder helpers (text_kernels.cuh): the row norms, activations, embedding lookup
// and the small attention of the T5 / CLIP encoders behind models/modules/conditioner.py:5-37.  Their Linears are vcb_gemm_bf16.
#include "../../include/vcb200.h"
#include "host_util.cuh"
#include "text_kernels.cuh"

using namespace vcb;

namespace {
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace

extern "C" int vcb_embedding_bf16(const void* table, int64_t vocab, int32_t dim, const int64_t* ids, const void* pos_table, int32_t L,
                                  void* out, int64_t ldo, int64_t n_tokens, void* stream) {
    if (!table || !ids || !out || vocab <= 0 || n_tokens <= 0 || n_tokens > 0x7fffffffLL) return set_error("embedding: bad arguments");
    if (dim <= 0 || dim % 8 || ldo % 8 || !aligned16(table) || !aligned16(out) || (pos_table && (!aligned16(pos_table) || L <= 0)))
        return set_error("embedding: dim / ldo must be multiples of 8, pointers 16-byte aligned, L > 0 with a position table");
    if (int rc = ensure_device()) return rc;
    ProfScope prof(PROF_OTHER, stream);
    embedding_kernel<<<(unsigned)n_tokens, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)table, (const long long*)ids,
                                                                          (const __nv_bfloat16*)pos_table, (__nv_bfloat16*)out, (long long)ldo,
                                                                          (long long)n_tokens, (int)dim, (int)(L > 0 ? L : 1), (long long)vocab);
    return check_launch("embedding");
}

static int rownorm(const char* what, bool affine, const void* x, int64_t ldx, const void* weight, const void* bias, void* y, int64_t ldy,
                   int64_t rows, int32_t dim, float eps, void* stream) {
    if (!x || !weight || !y || (affine && !bias) || rows <= 0 || dim <= 0) return set_error("%s: bad arguments", what);
    if (dim % 8 || ldx % 8 || ldy % 8 || !aligned16(x) || !aligned16(y) || !aligned16(weight) || (bias && !aligned16(bias)))
        return set_error("%s: dim / ldx / ldy must be multiples of 8 and the pointers 16-byte aligned", what);
    if (int rc = ensure_device()) return rc;
    ProfScope prof(PROF_LN, stream);
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (affine)
        rownorm_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (long long)ldx, (const __nv_bfloat16*)weight,
                                                                       (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, (long long)ldy, (long long)rows, (int)dim, eps);
    else
        rownorm_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (long long)ldx, (const __nv_bfloat16*)weight, nullptr,
                                                                        (__nv_bfloat16*)y, (long long)ldy, (long long)rows, (int)dim, eps);
    return check_launch(what);
}

extern "C" int vcb_rmsnorm_weight(const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int64_t rows, int32_t dim, float eps,
                                  void* stream) {
    return rownorm("rmsnorm_weight", false, x, ldx, weight, nullptr, y, ldy, rows, dim, eps, stream);
}

extern "C" int vcb_layernorm_affine(const void* x, int64_t ldx, const void* weight, const void* bias, void* y, int64_t ldy, int64_t rows,
                                    int32_t dim, float eps, void* stream) {
    return rownorm("layernorm_affine", true, x, ldx, weight, bias, y, ldy, rows, dim, eps, stream);
}

extern "C" int vcb_gated_gelu(const void* ab, int64_t ld, void* out, int64_t ldo, int64_t rows, int32_t dff, void* stream) {
    if (!ab || !out || rows <= 0 || dff <= 0) return set_error("gated_gelu: bad arguments");
    if (dff % 8 || ld % 8 || ldo % 8 || ld < 2 * (int64_t)dff || !aligned16(ab) || !aligned16(out))
        return set_error("gated_gelu: dff / ld / ldo must be multiples of 8, ld >= 2 * dff, pointers 16-byte aligned");
    if (int rc = ensure_device()) return rc;
    ProfScope prof(PROF_OTHER, stream);
    const long long work = rows * (long long)(dff / 8);
    gated_gelu_kernel<<<(unsigned)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)ab, (long long)ld, (__nv_bfloat16*)out,
                                                                                          (long long)ldo, (long long)rows, (int)dff);
    return check_launch("gated_gelu");
}

extern "C" int vcb_quick_gelu(const void* x, void* y, int64_t n, void* stream) {
    if (!x || !y || n <= 0 || n % 2) return set_error("quick_gelu: bad arguments");
    if (int rc = ensure_device()) return rc;
    ProfScope prof(PROF_OTHER, stream);
    quick_gelu_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, (long long)n);
    return check_launch("quick_gelu");
}

extern "C" int vcb_attention_small(const vcb_attn_small_args* a, void* stream) {
    if (!a || !a->q || !a->k || !a->v || !a->out) return set_error("attention_small: null argument");
    if (a->B <= 0 || a->heads <= 0 || a->L <= 0 || a->L > kSaMaxL) return set_error("attention_small: need 1 <= L <= %d, B, heads > 0", kSaMaxL);
    if (a->head_dim != 64) return set_error("attention_small: head_dim must be 64 (T5 v1.1 and CLIP-L text heads)");
    if (a->ld % 8 || a->ldo % 8 || !aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->out))
        return set_error("attention_small: ld / ldo must be multiples of 8 and the pointers 16-byte aligned");
    if (int rc = ensure_device()) return rc;
    static bool attr_done = false;
    static cudaError_t attr_err = cudaSuccess;
    if (!attr_done) {
        attr_err = cudaFuncSetAttribute(small_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_attn_smem(kSaMaxL));
        attr_done = true;
    }
    if (attr_err != cudaSuccess) return set_error("attention_small: cannot raise the shared-memory limit: %s", cudaGetErrorString(attr_err));
    SmallAttnParams p{(const __nv_bfloat16*)a->q, (const __nv_bfloat16*)a->k, (const __nv_bfloat16*)a->v, (long long)a->ld,
                      (const __nv_bfloat16*)a->bias, (__nv_bfloat16*)a->out, (long long)a->ldo, a->L, a->heads,
                      a->scale == 0.f ? 1.0f : a->scale, a->causal};
    ProfScope prof(PROF_ATTN, stream, a->B, a->L, a->heads, 0);
    const dim3 grid((a->L + kSaRows - 1) / kSaRows, a->heads, a->B);
    small_attention_kernel<<<grid, kSaThreads, small_attn_smem(a->L), (cudaStream_t)stream>>>(p);
    return check_launch("attention_small");
}
