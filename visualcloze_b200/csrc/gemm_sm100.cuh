// gemm_sm100.cuh -- persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] = epilogue( A[M,K] (bf16, K-major) * W[N,K]^T (bf16, K-major), fp32 accumulate in TMEM )
//
// One CTA (or CTA pair, cta_group::2) per SM, 320 threads:
//   warp 0      TMA producer      global -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer        one thread issues tcgen05.mma; accumulators double-buffered in TMEM
//   warps 2..9  epilogue          tcgen05.ld (thread == accumulator row, two warps per row block, half the columns
//                                 each) -> fused math -> global
//
// Epilogues replace the reference's un-fused elementwise passes and keep the
// reference's bf16 rounding points under CUDA autocast (models/modules/layers.py:158-245).
#pragma once
#include "vcb_common.cuh"

namespace vcb {

enum GemmEpilogue : int {
    EPI_BIAS = 0,       // out = bf16(acc + bias)
    EPI_BIAS_GELU = 1,  // out = bf16(gelu_tanh(bf16(acc + bias)))                         layers.py:143,154
    EPI_GATE_RES = 2,   // out = bf16(res + bf16(gate * bf16(acc + bias)))                  layers.py:190-195,245
    EPI_QKV = 3,        // q,k: bias -> QK-RMSNorm -> RoPE ; v: bias                        layers.py:165-174, math.py:112
    EPI_LINEAR1 = 4,    // cols < 3H as EPI_QKV into out ; cols >= 3H as EPI_BIAS_GELU into out2   layers.py:235-244
    EPI_BIAS_F32 = 5,   // out(fp32) = acc + bias   (VAE mid-block attention scores, autoencoder.py:47)
};
// EPI_GATE_RES with gate == nullptr is the plain residual epilogue out = bf16(res + bf16(acc + bias))
// (ResnetBlock / AttnBlock skip connections, autoencoder.py:52,82).

// A-operand addressing: a [batch, rows, K] matrix, or the 3x3 neighbourhood of an NHWC image (implicit GEMM, no im2col):
// k-block kb covers filter tap kb / (Cin/64) and channels 64*(kb % (Cin/64)); the tile's 128 rows are a 16 x 8 pixel
// patch loaded by ONE 4-D TMA box at the tap's (dx, dy) shift -- out-of-image pixels are zero-filled by the TMA unit,
// which is exactly the conv's zero padding.
enum GemmAMode : int { A_MATRIX = 0, A_CONV3X3 = 1 };
constexpr int kSpMaxRanks = 8;
constexpr int kConvTileW = 16, kConvTileH = 8;

struct GemmParams {
    int N, K;
    // A is a [batch, rows_per_batch, K] tensor (3-D TMA map; batch == 1 for a plain matrix).  M tiles never
    // straddle samples.  Row i of sample b is written to output row  b * out_batch_rows + out_row_offset + i,
    // which lets the txt / img streams read from and write to one joint [B, L, *] buffer without copies.
    int batch;
    int rows_per_batch;
    int out_batch_rows;
    int out_row_offset;
    // Tile raster.  0 (default): M-fastest -- the CTAs of a wave share few B (weight) tiles and cover all A rows.  1: N-fastest -- a
    // wave covers a band of A rows x ALL B tiles, A evict-first, B evict-last.  Measured on linear2 (A 122 MB, B 94 MB, three waves):
    // 522 MB of DRAM traffic per launch M-fastest, 642 MB N-fastest (algorithmic 265 MB), same time -- kept as an experiment switch.
    int n_fastest;
    const float* bias;           // [N] fp32 (may be null)
    // FP8 (e4m3) operands (kFp8 instantiations): acc is rescaled by a_scale[mapped output row] * w_scale[column] before the bias --
    // per-row activation scales written by the producing LayerNorm kernel, per-output-channel weight scales from packing time
    const float* a_scale;        // [rows of the joint buffer] fp32 (null = 1)
    const float* w_scale;        // [N] fp32 (null = 1)
    __nv_bfloat16* out;          // primary output
    long long ldo;               // elements
    int out_col_offset;
    // EPI_GATE_RES
    const __nv_bfloat16* gate;   // [B, gate_stride] bf16, column n
    long long gate_stride;
    const __nv_bfloat16* res;    // [rows, ld_res]; may alias out
    long long ld_res;
    // EPI_GATE_RES, optional: per-row partial sums of the bf16 residual stream this GEMM writes, one float2 (sum, sum of squares)
    // per 64 output columns at row_stats[orow * (N / 64) + column / 64] -- the statistics pass of the AdaLN LayerNorm that
    // follows (layers.py:191,195,234).  Every slot is written by exactly one thread: deterministic, no atomics.  Needs N % 64 == 0
    // and a tile whose column halves are multiples of 64 (BLOCK_N 128 or 256).
    float2* row_stats;
    // EPI_QKV / EPI_LINEAR1
    int hidden;                  // H (3H = end of the qkv columns); head_dim is 128
    const __nv_bfloat16* q_scale;  // [128] RMSNorm scale for q
    const __nv_bfloat16* k_scale;  // [128]
    const float2* rope;          // pair-major [64][rope_rows] (cos, sin); row index = the mapped output row
    long long rope_rows;
    __nv_bfloat16* out2;         // EPI_LINEAR1: gelu(mlp) destination
    long long ldo2;
    int out2_col_offset;
    // A_CONV3X3: OUTPUT height / width, input channels (K == 9 * conv_C); batch = images; rows_per_batch = H * W.
    // conv_stride 1 (pad 1 all round) or 2 (pad right/bottom only, autoencoder.py:91-95): for stride 2 the tensor map carries
    // elementStrides = 2, so one box still lands as a dense 16 x 8 pixel patch; coordinates are in input pixels.
    int conv_H, conv_W, conv_C, conv_stride;
    // Sequence-parallel head routing (EPI_QKV / EPI_LINEAR1, batch == 1; sp_world <= 1 = off).  Token rows are sharded over
    // sp_world ranks, heads over the same ranks for attention: the q/k/v columns of head h are stored straight into rank
    // h / (heads/W)'s peer-mapped qkv buffer [W * rows, 3 * hidden / W] over NVLink, at row sp_row_offset + output row.
    // The kernel's kSp instantiation stages each 32-row x 64-column piece in shared memory and ships it with one TMA tile
    // store through SpMaps (one map per destination rank, covering exactly this problem's rows of that rank's buffer).
    int sp_world, sp_row_offset;
    __nv_bfloat16* sp_out[kSpMaxRanks];
};

// Tensor maps of the sequence-parallel destinations: m[g][r] = rows [sp_row_offset + out_row_offset, + rows_per_batch) of rank
// r's qkv buffer for problem g of the launch; box 64 columns x 32 rows, 128-byte swizzle.  Rows past the problem's extent
// are clipped by the TMA unit, which is what masks the partial last M tile.
template <bool kSp>
struct SpMapsT { CUtensorMap m[2][kSpMaxRanks]; };
template <>
struct SpMapsT<false> { int unused; };
constexpr int kSpStageBytes = 8 * 32 * 128;    // one 32-row x 128-byte staging tile per epilogue warp

// Stream-K tail (optional): the tiles of the partial last wave are cut along K into one equal contiguous range per CTA
// (pair), so no SM idles while a few CTAs finish whole tiles.  A split tile is produced by up to kSkMaxParts CTAs: the one
// holding the HEAD k range (kb0 == 0) finalises; the others dump fp32 partials to their workspace slot and raise a flag.
// A waiter only waits on CTAs with a larger index, whose contributing item is their FIRST item of the phase and waits on
// nothing: deadlock-free.  (A fully contiguous stream-K over ALL tiles was measured 30 % slower: CTAs of one wave no
// longer share B tiles, and the 100-200 MB operands stop fitting in L2.)
constexpr int kSkMaxParts = 8;
struct StreamKParams {
    float* ws;              // [gridDim.x][128][BLOCK_N] fp32 partial accumulators (one slot per CTA)
    int* flags;             // [gridDim.x]
    int epoch;              // value that marks "partial of this launch is ready"
    int enabled;
};

struct WorkIter {
    // phase 1: the full waves with the plain strided schedule (CTA pairs of one wave share B tiles in L2);
    // phase 2 (stream-K only): the R = tiles % pairs tiles of the partial last wave are cut into equal contiguous
    // (tile, k-block) ranges, one per CTA pair.
    bool sk;
    int t, stride, full_tiles, num_tiles, num_kb;
    long long u, u1;
    __device__ __forceinline__ WorkIter(const StreamKParams& skp, int cluster_id, int num_clusters, int num_tiles_, int num_kb_)
        : sk(skp.enabled != 0), t(cluster_id), stride(num_clusters), num_tiles(num_tiles_), num_kb(num_kb_) {
        full_tiles = sk ? (num_tiles_ / num_clusters) * num_clusters : num_tiles_;
        const long long U = (long long)(num_tiles_ - full_tiles) * num_kb_;
        u = U * cluster_id / num_clusters;
        u1 = U * (cluster_id + 1) / num_clusters;
    }
    __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
        if (t < full_tiles) {
            tile = t; kb0 = 0; kb1 = num_kb; t += stride;
            return true;
        }
        if (!sk || u >= u1) return false;
        const int lt = (int)(u / num_kb);
        tile = full_tiles + lt;
        kb0 = (int)(u - (long long)lt * num_kb);
        const long long rem = u1 - u;
        kb1 = rem < (long long)(num_kb - kb0) ? kb0 + (int)rem : num_kb;
        u += kb1 - kb0;
        return true;
    }
};

VCB_DEVICE int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
VCB_DEVICE void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;       // TMA warp + MMA warp + 8 epilogue warps

template <int BLOCK_N, int kCtaGroup>
struct GemmCfg {
    static constexpr int kBRows = BLOCK_N / kCtaGroup;           // B rows held by one CTA
    static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
    static constexpr int kBBytes = kBRows * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kAccStride = BLOCK_N;                   // TMEM columns per accumulator stage
    static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                     : (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
};

// ----------------------------------------------------------------------------------------------
// epilogue helpers (one thread == one accumulator row; v[] holds 32 consecutive columns in fp32)
// ----------------------------------------------------------------------------------------------
VCB_DEVICE void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], int n0, int N) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (n0 + g * 8 < N) {
            uint4 u;
            u.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
            u.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
            u.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
            u.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + g * 8) = u;
        }
    }
}

// One 32-column chunk of this thread's accumulator row, in two halves so that callers can put their own global loads (gate,
// residual, RoPE, norm scales) between them: acc_issue starts the asynchronous TMEM load; acc_finish first issues the bias /
// weight-scale loads, THEN waits for the accumulator -- the three latencies overlap instead of adding up (the epilogue warps are
// latency-bound: two per scheduler, ncu long_scoreboard was the top stall).  The registers of `r` must not be touched between the
// two calls.  sa: this thread's row scale (1 for bf16 operands); ws: per-column weight scales or null.
// bf16: v = bf16(acc + bias) (the reference's Linear output); fp8: v = bf16(fma(acc, sa * ws[n], bias)).
VCB_DEVICE void acc_issue(uint32_t taddr, uint32_t (&r)[32]) {
    __syncwarp();                                   // tcgen05.ld is .sync.aligned: reconverge after predicated stores
    tmem_ld_x32(taddr, r);
}
// kRound: round the Linear output to the bf16 grid (the reference's Linear returns a bf16 tensor).  The fp8 instantiations skip
// the intermediate roundings -- their values are off the reference's grid already and every rounding is two more instructions
// in an epilogue that bounds the short-K fp8 tiles; the final store still rounds once.
// kHoist: issue the bias / scale loads before the accumulator wait (default); false = after it, for callers that keep other
// prefetched data live across the wait and cannot spare 64 registers (the loads hit L1: every row reads the same addresses).
template <bool kRound = true, bool kHoist = true>
VCB_DEVICE void acc_finish(uint32_t (&r)[32], const float* __restrict__ bias, int n0, int N, float (&v)[32], float sa = 1.0f,
                           const float* __restrict__ ws = nullptr) {
    auto rnd = [](float x) { return kRound ? bf16_round(x) : x; };
    if (n0 + 32 <= N) {
        if constexpr (kHoist) {
            float4 b4[8];
            if (bias != nullptr) {
#pragma unroll
                for (int q = 0; q < 8; ++q) b4[q] = __ldg(reinterpret_cast<const float4*>(bias + n0) + q);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) b4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (ws != nullptr) {
                float4 w4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w4[q] = __ldg(reinterpret_cast<const float4*>(ws + n0) + q);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    v[q * 4 + 0] = rnd(fmaf(__uint_as_float(r[q * 4 + 0]), sa * w4[q].x, b4[q].x));
                    v[q * 4 + 1] = rnd(fmaf(__uint_as_float(r[q * 4 + 1]), sa * w4[q].y, b4[q].y));
                    v[q * 4 + 2] = rnd(fmaf(__uint_as_float(r[q * 4 + 2]), sa * w4[q].z, b4[q].z));
                    v[q * 4 + 3] = rnd(fmaf(__uint_as_float(r[q * 4 + 3]), sa * w4[q].w, b4[q].w));
                }
            } else {
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    v[q * 4 + 0] = rnd(__uint_as_float(r[q * 4 + 0]) + b4[q].x);
                    v[q * 4 + 1] = rnd(__uint_as_float(r[q * 4 + 1]) + b4[q].y);
                    v[q * 4 + 2] = rnd(__uint_as_float(r[q * 4 + 2]) + b4[q].z);
                    v[q * 4 + 3] = rnd(__uint_as_float(r[q * 4 + 3]) + b4[q].w);
                }
            }
        } else {
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + n0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (ws != nullptr) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ws + n0) + q);
                    v[q * 4 + 0] = rnd(fmaf(__uint_as_float(r[q * 4 + 0]), sa * w4.x, b4.x));
                    v[q * 4 + 1] = rnd(fmaf(__uint_as_float(r[q * 4 + 1]), sa * w4.y, b4.y));
                    v[q * 4 + 2] = rnd(fmaf(__uint_as_float(r[q * 4 + 2]), sa * w4.z, b4.z));
                    v[q * 4 + 3] = rnd(fmaf(__uint_as_float(r[q * 4 + 3]), sa * w4.w, b4.w));
                } else {
                    v[q * 4 + 0] = rnd(__uint_as_float(r[q * 4 + 0]) + b4.x);
                    v[q * 4 + 1] = rnd(__uint_as_float(r[q * 4 + 1]) + b4.y);
                    v[q * 4 + 2] = rnd(__uint_as_float(r[q * 4 + 2]) + b4.z);
                    v[q * 4 + 3] = rnd(__uint_as_float(r[q * 4 + 3]) + b4.w);
                }
            }
        }
    } else {                                        // ragged last chunk of N: column by column
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float sw = (ws != nullptr && n0 + j < N) ? sa * __ldg(ws + n0 + j) : 1.0f;
            const float bj = (bias != nullptr && n0 + j < N) ? __ldg(bias + n0 + j) : 0.0f;
            v[j] = ws != nullptr ? rnd(fmaf(__uint_as_float(r[j]), sw, bj)) : rnd(__uint_as_float(r[j]) + bj);
        }
    }
}
template <bool kRound = true, bool kHoist = true>
VCB_DEVICE void load_acc_bias(uint32_t taddr, const float* __restrict__ bias, int n0, int N, float (&v)[32], float sa = 1.0f,
                              const float* __restrict__ ws = nullptr) {
    uint32_t r[32];
    acc_issue(taddr, r);
    acc_finish<kRound, kHoist>(r, bias, n0, N, v, sa, ws);
}

// ----------------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------------
// kFp8: both operands are e4m3 bytes (K-major, 128 elements per 128-byte swizzle row); tcgen05.mma kind::f8f6f4, UMMA K = 32.
// A k-block is 128 bytes of K in either mode, so the pipeline (stage bytes, 4 MMAs of 32 bytes of K per k-block, descriptor
// advance) is unchanged; only the element count per k-block, the MMA kind and the epilogue's rescaling differ.
template <int BLOCK_N, int kCtaGroup, int kEpi, int kAMode = A_MATRIX, bool kSp = false, bool kFp8 = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p, const __grid_constant__ CUtensorMap tmap_a1,
                         const __grid_constant__ CUtensorMap tmap_b1, const GemmParams p1, const StreamKParams skp,
                         const __grid_constant__ SpMapsT<kSp> spm) {
    static_assert(!kSp || kEpi == EPI_QKV || kEpi == EPI_LINEAR1, "sequence-parallel routing lives in the head-structured epilogues");
    static_assert(!kFp8 || (kAMode == A_MATRIX && !kSp), "fp8 operands: plain matrices, single-GPU path");
    constexpr int kBlockKEl = kFp8 ? 2 * kBlockK : kBlockK;          // elements of K per k-block (128 bytes either way)
    // Grouped launch: an optional second problem (p1.batch > 0) with the same N, K and epilogue but its own operands --
    // the txt stream of a DoubleStreamBlock rides in the img stream's launch and fills its partial last wave.
    using Cfg = GemmCfg<BLOCK_N, kCtaGroup>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_stage = smem + kStages * Cfg::kStageBytes;                 // kSp: 8 x 4 KB epilogue staging (1 KB aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + (kSp ? kSpStageBytes : 0));
    uint64_t* full_bar = bars;                    // [kStages]
    uint64_t* empty_bar = bars + kStages;         // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;     // [2]
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const uint32_t warp = warp_id_uniform();
    const uint32_t lane = lane_id();
    const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
    const bool is_leader = (cta_rank == 0);

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (p1.batch > 0) {
            tma_prefetch_desc(&tmap_a1);
            tma_prefetch_desc(&tmap_b1);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 8 * kCtaGroup);     // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kCtaGroup>(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                     // everything above overlapped the previous kernel's tail

    static_assert(kAMode == A_MATRIX || kCtaGroup == 1, "conv mode is single-CTA");
    const int tile_m = kBlockM * kCtaGroup;
    const int conv_tx = (kAMode == A_CONV3X3) ? (p.conv_W + kConvTileW - 1) / kConvTileW : 1;
    const int m_per_sample = (kAMode == A_CONV3X3) ? conv_tx * ((p.conv_H + kConvTileH - 1) / kConvTileH)
                                                   : (p.rows_per_batch + tile_m - 1) / tile_m;
    const int num_m = m_per_sample * p.batch;
    const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles0 = num_m * num_n;
    const int m_per_sample1 = (p1.rows_per_batch + tile_m - 1) / tile_m;
    const int num_m1 = p1.batch > 0 ? m_per_sample1 * p1.batch : 0;
    const int num_tiles = tiles0 + num_m1 * num_n;
    const int num_kb = (p.K + kBlockKEl - 1) / kBlockKEl;
    const int cluster_id = blockIdx.x / kCtaGroup;
    const int num_clusters = gridDim.x / kCtaGroup;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            WorkIter it(skp, cluster_id, num_clusters, num_tiles, num_kb);
            int t, kb0, kb1;
            while (it.next(t, kb0, kb1)) {
                const bool g1 = t >= tiles0;
                const int tt = g1 ? t - tiles0 : t, nm = g1 ? num_m1 : num_m, mps = g1 ? m_per_sample1 : m_per_sample;
                const CUtensorMap* ta = g1 ? &tmap_a1 : &tmap_a;
                const CUtensorMap* tb = g1 ? &tmap_b1 : &tmap_b;
                const int mt = p.n_fastest ? tt / num_n : tt % nm;
                const int bi = mt / mps;
                const int m0 = (mt % mps) * tile_m + (int)cta_rank * kBlockM;
                const int n0 = (p.n_fastest ? tt % num_n : tt / nm) * BLOCK_N + (int)cta_rank * Cfg::kBRows;
                const uint64_t hint_a = p.n_fastest ? kEvictFirst : kEvictNormal, hint_b = p.n_fastest ? kEvictLast : kEvictNormal;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (is_leader) mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes * kCtaGroup);
                    if constexpr (kAMode == A_CONV3X3) {
                        const int cpb = p.conv_C / kBlockK;
                        const int tap = kb / cpb, cb = kb - tap * cpb;
                        const int mi = mt % mps;
                        const int pad = p.conv_stride == 1 ? 1 : 0;
                        const int x0 = (mi % conv_tx) * kConvTileW * p.conv_stride + (tap % 3) - pad;
                        const int y0 = (mi / conv_tx) * kConvTileH * p.conv_stride + (tap / 3) - pad;
                        tma_load_4d<false>(ta, &full_bar[stage], smem_a + stage * Cfg::kABytes, cb * kBlockK, x0, y0, bi,
                                           kEvictNormal);
                    } else {
                        tma_load_3d<kCtaGroup == 2>(ta, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBlockKEl,
                                                    m0, bi, hint_a);
                    }
                    tma_load_2d<kCtaGroup == 2>(tb, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockKEl,
                                                n0, hint_b);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        // The warp runs the loop convergently (descriptors in uniform registers); one elected lane issues tcgen05 ops.
        if (is_leader) {
            constexpr uint32_t idesc = kFp8 ? make_idesc_e4m3(kBlockM * kCtaGroup, BLOCK_N) : make_idesc_bf16(kBlockM * kCtaGroup, BLOCK_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            WorkIter it(skp, cluster_id, num_clusters, num_tiles, num_kb);
            int t, kb0, kb1;
            while (it.next(t, kb0, kb1)) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes), 16, 1024, kSwizzle128B);
                    const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes), 16, 1024, kSwizzle128B);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // advance 16 bf16 = 32 B inside the 128B swizzle span: +2 in the (addr >> 4) field
                            umma_ss<kCtaGroup, kFp8>(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                                                     (kb > kb0 || k != 0) ? 1u : 0u);
                        }
                        umma_commit<kCtaGroup>(&empty_bar[stage]);
                        if (kb == kb1 - 1) umma_commit<kCtaGroup>(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps (8: two per TMEM lane quarter, each owning half the tile's columns) ============
        const uint32_t quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const uint32_t half = (warp - 2) >> 2;              // 0: columns [0, BLOCK_N/2), 1: [BLOCK_N/2, BLOCK_N)
        const int row_in_tile = quarter * 32 + lane;
        constexpr int kHalfN = BLOCK_N / 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        WorkIter it(skp, cluster_id, num_clusters, num_tiles, num_kb);
        int t, kb0, kb1;
        const int my_cta = cluster_id * kCtaGroup + (int)cta_rank;
        while (it.next(t, kb0, kb1)) {
            const bool g1 = t >= tiles0;
            const GemmParams& P = g1 ? p1 : p;
            const bool sk_contrib = kb0 > 0;                       // tail part of a split tile: dump partials, no epilogue
            const bool sk_final = !sk_contrib && kb1 < num_kb;     // head part: add the partner's partials, then epilogue
            const int tt = g1 ? t - tiles0 : t, nm = g1 ? num_m1 : num_m, mps = g1 ? m_per_sample1 : m_per_sample;
            const int mt = p.n_fastest ? tt / num_n : tt % nm;
            const int b = mt / mps;
            const int n_tile0 = (p.n_fastest ? tt % num_n : tt / nm) * BLOCK_N;
            int i;
            bool row_ok;
            if constexpr (kAMode == A_CONV3X3) {
                const int mi = mt % mps;
                const int px = (mi % conv_tx) * kConvTileW + (row_in_tile % kConvTileW);
                const int py = (mi / conv_tx) * kConvTileH + (row_in_tile / kConvTileW);
                row_ok = px < P.conv_W && py < P.conv_H;
                i = py * P.conv_W + px;
            } else {
                i = (mt % mps) * tile_m + (int)cta_rank * kBlockM + row_in_tile;
                row_ok = i < P.rows_per_batch;
            }
            const long long orow = (long long)b * P.out_batch_rows + P.out_row_offset + (row_ok ? i : 0);
            [[maybe_unused]] const float sa = (kFp8 && P.a_scale) ? __ldg(P.a_scale + orow) : 1.0f;
            [[maybe_unused]] const float* ws = kFp8 ? P.w_scale : nullptr;
            // intermediate roundings to the bf16 grid: the reference's autocast semantics; skipped for e4m3 operands (see acc_finish)
            [[maybe_unused]] auto rnd = [](float x) { return kFp8 ? x : bf16_round(x); };

            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * Cfg::kAccStride;

            if (sk_contrib) {
                // dump this CTA's raw fp32 accumulator rows into its workspace slot, then publish
                // slot layout [BLOCK_N / 4 float4 columns][128 rows]: the lanes of a warp (consecutive rows) write consecutive
                // 16-byte words -- fully coalesced 512-byte stores (a row-major slot costs one sector per lane and store)
                uint4* slot = reinterpret_cast<uint4*>(skp.ws + (long long)my_cta * kBlockM * BLOCK_N) + row_in_tile;
#pragma unroll 1
                for (int c = 0; c < kHalfN / 32; ++c) {
                    const int cc = half * (kHalfN / 32) + c;
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld_x32(taddr + cc * 32, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4)
                        slot[(cc * 8 + q4) * kBlockM] = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
                }
                __threadfence();
                named_bar_sync(1, 256);                             // all 8 epilogue warps have written and fenced
                if (warp == 2 && lane == 0) st_release_gpu(skp.flags + my_cta, skp.epoch);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty[acc], 0);
                    else mbar_arrive(&tmem_empty[acc]);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            if (sk_final) {
                // partners: the following CTA pairs (same rank) whose range starts inside this tile; each produced its part as
                // the FIRST item of its phase 2.  Fold their fp32 partials into my TMEM accumulator, then run the normal epilogue.
                const long long U = (long long)(num_tiles - it.full_tiles) * num_kb;
                const long long tile_end = (long long)(t - it.full_tiles + 1) * num_kb;
#pragma unroll 1
                for (int pc = cluster_id + 1; pc < num_clusters && U * pc / num_clusters < tile_end; ++pc) {
                    const int partner = pc * kCtaGroup + (int)cta_rank;
                    if (lane == 0) {
                        while (ld_acquire_gpu(skp.flags + partner) != skp.epoch) {
                        }
                    }
                    __syncwarp();
                    const float4* part = reinterpret_cast<const float4*>(skp.ws + (long long)partner * kBlockM * BLOCK_N) + row_in_tile;
#pragma unroll 1
                    for (int c = 0; c < kHalfN / 32; ++c) {
                        const int cc = half * (kHalfN / 32) + c;
                        uint32_t r[32];
                        __syncwarp();
                        tmem_ld_x32(taddr + cc * 32, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            const float4 f = __ldcg(part + (cc * 8 + q4) * kBlockM);
                            r[q4 * 4 + 0] = __float_as_uint(__uint_as_float(r[q4 * 4 + 0]) + f.x);
                            r[q4 * 4 + 1] = __float_as_uint(__uint_as_float(r[q4 * 4 + 1]) + f.y);
                            r[q4 * 4 + 2] = __float_as_uint(__uint_as_float(r[q4 * 4 + 2]) + f.z);
                            r[q4 * 4 + 3] = __float_as_uint(__uint_as_float(r[q4 * 4 + 3]) + f.w);
                        }
                        tmem_st_x32(taddr + cc * 32, r);
                    }
                    tmem_wait_st();
                }
            }

            if constexpr (kEpi == EPI_BIAS_F32) {
#pragma unroll 1
                for (int c = 0; c < kHalfN / 32; ++c) {
                    const int cc = half * (kHalfN / 32) + c;
                    const int n0 = n_tile0 + cc * 32;
                    if (n0 >= P.N) break;
                    uint32_t r[32];
                    __syncwarp();
                    tmem_ld_x32(taddr + cc * 32, r);
                    tmem_wait_ld();
                    if (row_ok) {
                        float* dst = reinterpret_cast<float*>(P.out) + orow * P.ldo + P.out_col_offset + n0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (n0 + q * 4 < P.N) {
                                float4 f;
                                f.x = __uint_as_float(r[q * 4 + 0]) + (P.bias ? __ldg(P.bias + n0 + q * 4 + 0) : 0.f);
                                f.y = __uint_as_float(r[q * 4 + 1]) + (P.bias ? __ldg(P.bias + n0 + q * 4 + 1) : 0.f);
                                f.z = __uint_as_float(r[q * 4 + 2]) + (P.bias ? __ldg(P.bias + n0 + q * 4 + 2) : 0.f);
                                f.w = __uint_as_float(r[q * 4 + 3]) + (P.bias ? __ldg(P.bias + n0 + q * 4 + 3) : 0.f);
                                *reinterpret_cast<float4*>(dst + q * 4) = f;
                            }
                        }
                    }
                }
            } else if constexpr (kEpi == EPI_BIAS || kEpi == EPI_BIAS_GELU || kEpi == EPI_GATE_RES) {
                [[maybe_unused]] float st_sum = 0.f, st_sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < kHalfN / 32; ++c) {
                    const int cc = half * (kHalfN / 32) + c;
                    const int n0 = n_tile0 + cc * 32;
                    if (n0 >= P.N) break;
                    float v[32];
                    uint32_t racc[32];
                    acc_issue(taddr + cc * 32, racc);
                    // gate / residual of this chunk: issued before the accumulator wait (neither depends on it)
                    [[maybe_unused]] uint4 gu[4], ru[4];
                    if constexpr (kEpi == EPI_GATE_RES) {
                        const uint32_t one2 = 0x3F803F80u;          // bf16 (1.0, 1.0)
                        const __nv_bfloat16* g = P.gate ? P.gate + (long long)b * P.gate_stride + n0 : nullptr;
                        const __nv_bfloat16* rs = P.res + orow * P.ld_res + n0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool in = row_ok && n0 + q * 8 < P.N;
                            gu[q] = (in && g) ? __ldg(reinterpret_cast<const uint4*>(g + q * 8)) : make_uint4(one2, one2, one2, one2);
                            ru[q] = in ? *reinterpret_cast<const uint4*>(rs + q * 8) : make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    acc_finish<!kFp8>(racc, P.bias, n0, P.N, v, sa, ws);
                    if (row_ok) {
                        if constexpr (kEpi == EPI_BIAS_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = kFp8 ? gelu_tanh_fast(v[j]) : gelu_tanh(v[j]);
                        }
                        if constexpr (kEpi == EPI_GATE_RES) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (n0 + q * 8 < P.N) {
                                    const uint32_t gw[4] = {gu[q].x, gu[q].y, gu[q].z, gu[q].w};
                                    const uint32_t rw[4] = {ru[q].x, ru[q].y, ru[q].z, ru[q].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        float2 gf = unpack_bf16x2(gw[e]);
                                        float2 rf = unpack_bf16x2(rw[e]);
                                        float a0 = rnd(gf.x * v[q * 8 + 2 * e]);
                                        float a1 = rnd(gf.y * v[q * 8 + 2 * e + 1]);
                                        v[q * 8 + 2 * e] = rf.x + a0;
                                        v[q * 8 + 2 * e + 1] = rf.y + a1;
                                    }
                                }
                            }
                            if constexpr (kAMode == A_MATRIX && (kHalfN % 64 == 0)) {
                                if (P.row_stats != nullptr) {
                                    // statistics of the values as STORED (bf16): the LayerNorm reads the stream back in bf16
                                    if ((c & 1) == 0) { st_sum = 0.f; st_sq = 0.f; }
#pragma unroll
                                    for (int j = 0; j < 32; ++j) {
                                        const float xr = bf16_round(v[j]);
                                        st_sum += xr;
                                        st_sq = fmaf(xr, xr, st_sq);
                                    }
                                    if (c & 1) P.row_stats[orow * (long long)(P.N >> 6) + (n0 >> 6)] = make_float2(st_sum, st_sq);
                                }
                            }
                        }
                        store_bf16x32(P.out + orow * P.ldo + P.out_col_offset + n0, v, n0, P.N);
                    }
                }
            } else {
                // EPI_QKV / EPI_LINEAR1: 128-column groups (one head each); group hg is handled by column-half hg % 2
                static_assert(kEpi != EPI_QKV && kEpi != EPI_LINEAR1 || BLOCK_N % 128 == 0, "head-structured epilogue");
#pragma unroll 1
                for (int hg = (int)half; hg < BLOCK_N / 128; hg += 2) {
                    const int ng = n_tile0 + hg * 128;           // first column of this 128-group
                    if (ng >= P.N) break;
                    const uint32_t tg = taddr + hg * 128;
                    const int region = ng / P.hidden;            // 0 q, 1 k, 2 v, >= 3 mlp (LINEAR1)
                    // destination of this head's 128 q / k / v columns: the local qkv buffer, or (sequence-parallel) the
                    // qkv buffer of the rank that owns the head
                    __nv_bfloat16* qkv_dst = P.out + orow * P.ldo + P.out_col_offset + ng;
                    [[maybe_unused]] const CUtensorMap* sp_map = nullptr;
                    [[maybe_unused]] int sp_col = 0;
                    [[maybe_unused]] uint8_t* stg = smem_stage + (warp - 2) * 4096;
                    [[maybe_unused]] const int sp_row0 = (mt % mps) * tile_m + (int)cta_rank * kBlockM + (int)quarter * 32;
                    if constexpr (kSp) {
                        if (region < 3) {
                            const int hw = P.hidden / P.sp_world;    // q (or k, v) columns per rank
                            const int cin = ng - region * P.hidden;
                            const int owner = cin / hw;
                            sp_map = &spm.m[g1 ? 1 : 0][owner];
                            sp_col = region * hw + (cin - owner * hw);
                        }
                    }
                    // kSp: 32 columns of this thread's row -> staging tile (128-byte swizzle); every second chunk the warp's
                    // 32 x 64 piece leaves as ONE asynchronous TMA tile store (NVLink for a remote owner)
                    auto sp_emit = [&](int c, const float (&v)[32]) {
                        if ((c & 1) == 0) {
                            if (lane == 0) tma_store_wait_read();           // previous piece has left the staging tile
                            __syncwarp();
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 u;
                            u.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
                            u.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                            u.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
                            u.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                            const int chunk = (c & 1) * 4 + g;
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((chunk ^ ((int)lane & 7)) << 4)) = u;
                        }
                        if (c & 1) {
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0 && sp_row0 < P.rows_per_batch) {
                                tma_store_2d(sp_map, stg, sp_col + (c >> 1) * 64, sp_row0);
                                tma_store_commit();
                            }
                        }
                    };
                    if (region >= 2) {
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            const int n0 = ng + c * 32;
                            float v[32];
                            load_acc_bias<!kFp8>(tg + c * 32, P.bias, n0, P.N, v, sa, ws);
                            if (kEpi == EPI_LINEAR1 && region >= 3) {
                                if (row_ok) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = kFp8 ? gelu_tanh_fast(v[j]) : gelu_tanh(v[j]);
                                    store_bf16x32(P.out2 + orow * P.ldo2 + P.out2_col_offset + (n0 - 3 * P.hidden), v, n0, P.N);
                                }
                            } else if constexpr (kSp) {
                                sp_emit(c, v);          // whole warp; rows past the problem's extent are clipped by the TMA unit
                            } else {
                                if (row_ok) store_bf16x32(qkv_dst + c * 32, v, n0, P.N);
                            }
                        }
                    } else {
                        const __nv_bfloat16* sc = region == 0 ? P.q_scale : P.k_scale;
                        // rope table is stored pair-major [64][rope_rows]: consecutive lanes (rows) read consecutive float2.
                        // Its entries are fetched ONE CHUNK AHEAD of their use (plain global loads are scoreboarded, so they may
                        // stay in flight across the accumulator waits and the math): the table is the only per-row global
                        // traffic of this epilogue and its L2 latency was what the two warps per scheduler could not hide.
                        const float2* rp = P.rope + orow;
                        auto load_rope = [&](int c, float2 (&dst)[16]) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                dst[j] = row_ok ? __ldg(rp + (long long)(c * 16 + j) * P.rope_rows) : make_float2(1.f, 0.f);
                        };
                        float2 cs_a[16], cs_b[16];
                        load_rope(0, cs_a);
                        // pass 1: sum of squares of the (bf16-rounded) projection over the head (RMSNorm, layers.py:68-72)
                        float ss = 0.f;
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            float v[32];
                            load_acc_bias<!kFp8, false>(tg + c * 32, P.bias, ng + c * 32, P.N, v, sa, ws);
#pragma unroll
                            for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
                        }
                        const float rrms = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
                        // pass 2: normalise, scale, rotate, store
                        auto chunk = [&](int c, float2 (&cs)[16], float2 (&cs_next)[16]) {
                            const int n0 = ng + c * 32;
                            float v[32];
                            uint32_t racc[32];
                            acc_issue(tg + c * 32, racc);
                            if (c < 3) load_rope(c + 1, cs_next);
                            acc_finish<!kFp8, false>(racc, P.bias, n0, P.N, v, sa, ws);
                            if (row_ok) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const uint4 su = __ldg(reinterpret_cast<const uint4*>(sc + c * 32 + q * 8));
                                    const uint32_t sw[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const int j = q * 8 + 2 * e;
                                        const float2 s2 = unpack_bf16x2(sw[e]);
                                        const float x0 = rnd(rnd(v[j] * rrms) * s2.x);
                                        const float x1 = rnd(rnd(v[j + 1] * rrms) * s2.y);
                                        const float2 t = cs[j >> 1];
                                        // math.py:112-117: two fp32 products, one fp32 add (no contraction)
                                        v[j] = __fadd_rn(__fmul_rn(t.x, x0), __fmul_rn(-t.y, x1));
                                        v[j + 1] = __fadd_rn(__fmul_rn(t.y, x0), __fmul_rn(t.x, x1));
                                    }
                                }
                                if constexpr (!kSp) store_bf16x32(qkv_dst + c * 32, v, n0, P.N);
                            }
                            if constexpr (kSp) sp_emit(c, v);
                        };
                        chunk(0, cs_a, cs_b);
                        chunk(1, cs_b, cs_a);
                        chunk(2, cs_a, cs_b);
                        chunk(3, cs_b, cs_a);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty[acc], 0);
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        __syncwarp();
    }

    // ===================== teardown =====================
    if constexpr (kSp) {
        if (warp >= 2 && lane == 0) tma_store_wait_all();       // this warp's tile stores have landed (staging smem is released)
    }
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kCtaGroup>(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace vcb
