// attn_sm100.cuh -- joint txt||img flash attention forward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces models/math.py:63-99 (attention -> _upad_input -> flash_attn_varlen_func -> pad_input): non-causal
// softmax(Q K^T / sqrt(128)) V over right-padded samples, head_dim 128, bf16 in / fp32 accumulate / bf16 out.
// Padding is handled with per-sample `seqlens` instead of the reference's gather/scatter unpad: keys >= seqlen are
// masked to -inf, query rows >= seqlen produce zeros (pad_input semantics).  RoPE and QK-RMSNorm are already applied
// by the QKV GEMM epilogue (gemm_sm100.cuh, EPI_QKV).
//
// One CTA per (128-query tile, head, sample), 192 threads:
//   warp 0      TMA producer  Q once; K/V tiles through a ring of 32 KB slots (two 128x64 SW128 boxes per tile)
//   warp 1      MMA issuer    S[j%2] = Q K_j^T (SS, K-major x K-major) ; O += P_j V_j (TS: P from TMEM, V MN-major)
//   warps 2..5  softmax       thread == query row: S -> registers, online softmax (exp2, lazy rescale), P -> TMEM
//                             (bf16, aliasing the S buffer), O correction in TMEM, final O / l -> global
// TMEM columns: S0 [0,128)  S1 [128,256)  O [256,384)   (fp32) ; P_j occupies the first 64 columns of S[j%2].
#pragma once
#include "vcb_common.cuh"

namespace vcb {

struct AttnParams {
    int B, L, H;                 // samples, padded tokens per sample, heads
    const int* seqlens;          // [B] valid tokens per sample (<= L); null = all L
    __nv_bfloat16* out;          // [B*L, ldo], head h at columns out_col_offset + h*128
    long long ldo;
    int out_col_offset;
    int q_col, k_col, v_col;     // column of head 0 of q / k / v inside the qkv matrix
    float scale_log2;            // head_dim^-0.5 * log2(e)
    float fixed_max;             // > 0: upper bound of |score * scale_log2| guaranteed by the caller (attn_fwd3 kFixed)
    // Sequence-parallel output routing (attn_fwd3 only, B == 1; sp_world <= 1 = off): query row r belongs to rank
    // r / sp_rows and is stored over NVLink into that rank's peer-mapped buffer sp_out[rank] at row r % sp_rows.
    int sp_world, sp_rows;
    __nv_bfloat16* sp_out[8];
};

// first element of output row `row` of sample b (columns are added by the caller)
VCB_DEVICE __nv_bfloat16* attn_out_row(const AttnParams& p, int b, int row) {
    if (p.sp_world > 1) {
        const int owner = min(row / p.sp_rows, p.sp_world - 1);
        return p.sp_out[owner] + (long long)(row - owner * p.sp_rows) * p.ldo;
    }
    return p.out + ((long long)b * p.L + row) * p.ldo;
}

constexpr int kAttnThreads = 192;
constexpr int kAttnTile = 128;           // query rows per CTA == kv rows per tile == head_dim
constexpr int kKvSlots = 5;              // 32 KB each
constexpr int kSlotBytes = 128 * 128 * 2;
constexpr int kAttnSmemBytes = (1 + kKvSlots) * kSlotBytes + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;   // log2 units: only rescale O when the row max grows by > 2^8

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
    const int q_tile = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int seqlen = p.seqlens ? min(p.seqlens[b], p.L) : p.L;
    const int q0 = q_tile * kAttnTile;
    const uint32_t warp = warp_id_uniform();
    const uint32_t lane = lane_id();

    if (q0 >= seqlen) {
        // whole tile is padding: zero rows (pad_input), nothing else to do
        if (warp >= 2) {
            const int row = q0 + (warp & 3) * 32 + lane;
            if (row < p.L) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)b * p.L + row) * p.ldo + p.out_col_offset + head * 128);
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[i] = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;
    uint8_t* smem_kv = smem + kSlotBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + kKvSlots) * kSlotBytes);
    uint64_t* q_full = bars;                 // [1]
    uint64_t* kv_full = bars + 1;            // [kKvSlots]
    uint64_t* kv_empty = kv_full + kKvSlots; // [kKvSlots]
    uint64_t* s_full = kv_empty + kKvSlots;  // [2]
    uint64_t* p_full = s_full + 2;           // [2]
    uint64_t* o_done = p_full + 2;           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < kKvSlots; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); }
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s0 = tmem_base, tmem_o = tmem_base + 256;

    const int n_kv = (seqlen + kAttnTile - 1) / kAttnTile;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, kSlotBytes);
            tma_load_3d<false>(&tmap_qkv, q_full, smem_q, p.q_col + head * 128, q0, b, kEvictFirst);
            tma_load_3d<false>(&tmap_qkv, q_full, smem_q + kSlotBytes / 2, p.q_col + head * 128 + 64, q0, b, kEvictFirst);
            int slot = 0;
            uint32_t phase = 0;
            auto load_tile = [&](int col, int j) {
                mbar_wait(&kv_empty[slot], phase ^ 1);
                mbar_expect_tx(&kv_full[slot], kSlotBytes);
                uint8_t* dst = smem_kv + slot * kSlotBytes;
                tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst, col + head * 128, j * kAttnTile, b, kEvictLast);
                tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst + kSlotBytes / 2, col + head * 128 + 64, j * kAttnTile, b, kEvictLast);
                if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
            };
            // consumption order of the MMA warp: K0, [K(j+1), V(j)] for j = 0..n-1
            load_tile(p.k_col, 0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) load_tile(p.k_col, j + 1);
                load_tile(p.v_col, j);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);   // S = Q K^T : both K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);   // O += P V   : V is MN-major (d contiguous)
            int slot = 0;
            uint32_t phase = 0;
            auto issue_qk = [&](int j) {
                mbar_wait(&kv_full[slot], phase);
                tc_fence_after();
                const uint32_t qa = smem_u32(smem_q), ka = smem_u32(smem_kv + slot * kSlotBytes);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t off = (ks >> 2) * (kSlotBytes / 2) + (ks & 3) * 32;
                    umma_ss<1>(tmem_s0 + (j & 1) * 128, make_smem_desc(qa + off, 16, 1024, kSwizzle128B),
                               make_smem_desc(ka + off, 16, 1024, kSwizzle128B), idesc_qk, ks != 0);
                }
                umma_commit<1>(&kv_empty[slot]);
                umma_commit<1>(&s_full[j & 1]);
                if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) issue_qk(j + 1);
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);       // P_j in TMEM, O rescaled if needed
                mbar_wait(&kv_full[slot], phase);              // V_j landed
                tc_fence_after();
                const uint32_t va = smem_u32(smem_kv + slot * kSlotBytes);
                const uint32_t p_tmem = tmem_s0 + (j & 1) * 128;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    // 16 kv rows per step: V rows are 128 B apart, 8-row groups 1024 B apart, second 64-wide d block +16 KB
                    umma_ts(tmem_o, p_tmem + ks * 8, make_smem_desc(va + ks * 2048, kSlotBytes / 2, 1024, kSwizzle128B),
                            idesc_pv, (j | ks) != 0);
                }
                umma_commit<1>(&kv_empty[slot]);
                umma_commit<1>(o_done);
                if (++slot == kKvSlots) { slot = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax / correction / epilogue =====================
        const uint32_t quarter = warp & 3;
        const int row = q0 + quarter * 32 + lane;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        float m_run = -INFINITY;      // running max in scaled log2 units
        float l_run = 0.f;            // running sum of exp2
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t s_addr = tmem_s0 + lane_addr + (j & 1) * 128;
            uint32_t sr[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_x32(s_addr + c * 32, sr[c]);
            tmem_wait_ld();
            const int kv_left = seqlen - j * kAttnTile;     // columns >= kv_left are padding
            float m_tile = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s = __uint_as_float(sr[c][i]) * p.scale_log2;
                    if (c * 32 + i >= kv_left) s = -INFINITY;
                    sr[c][i] = __float_as_uint(s);
                    m_tile = fmaxf(m_tile, s);
                }
            // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
            const bool grow = (m_tile - m_run) > kRescaleThreshold;     // true on the first tile (m_run = -inf)
            const float m_new = grow ? m_tile : m_run;
            const float alpha = grow ? exp2f(m_run - m_new) : 1.0f;     // exp2(-inf) = 0 on the first tile
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                mbar_wait(o_done, (j - 1) & 1);                         // P_{j-1} V_{j-1} has landed in O
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_addr + c * 32, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_x32(tmem_o + lane_addr + c * 32, o);
                }
            }
            l_run *= alpha;
            m_run = m_new;
            float l_tile = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = exp2f(__uint_as_float(sr[c][i]) - m_new);
                    float p1 = exp2f(__uint_as_float(sr[c][i + 1]) - m_new);
                    l_tile += p0 + p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
                tmem_st_x16(s_addr + c * 16, pk);               // P_j: 2 bf16 per 32-bit TMEM column
            }
            l_run += l_tile;
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);
        }
        // ---- epilogue: O / l -> bf16 -> global ----
        mbar_wait(o_done, (n_kv - 1) & 1);
        tc_fence_after();
        const bool valid = row < seqlen;
        const float inv_l = 1.0f / l_run;
        __nv_bfloat16* dst = p.out + ((long long)b * p.L + row) * p.ldo + p.out_col_offset + head * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            __syncwarp();
            tmem_ld_x32(tmem_o + lane_addr + c * 32, o);
            tmem_wait_ld();
            if (!valid) {                                              // padded query rows -> 0 (pad_input)
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
            if (row < p.L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = u;
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

}  // namespace vcb
