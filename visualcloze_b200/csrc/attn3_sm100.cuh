// attn3_sm100.cuh -- joint attention forward, two 128-query tiles per CTA, ROW-SPLIT softmax (16 softmax warps).
//
// Two 128-query tiles per CTA share one K/V stream, and every query row is handled by TWO threads (64 key columns each; the two warps of a
// pair share a TMEM lane quarter and exchange the row max through shared memory with a 64-thread named barrier).  The
// per-tile softmax latency -- the serial link S_t(j) -> P_t(j) -> PV_t(j) -> QK_t(j+1) that bounded the thread-per-row version
// (ncu: 34 % of samples waiting on s_full, tensor pipe 51 %) -- is halved, and each SM sub-partition hosts 4 softmax warps.
//
// Contract: attn_common.cuh (models/math.py:63-99; per-sample seqlens instead of unpad/pad).  The tensor pipe is fed from two
// independent softmax pipelines that share one K/V stream:
//   warp 0       TMA producer   Q0, Q1 once; K/V tiles through a ring of 32 KB slots in the order K0 V0 K1 V1 ...
//   warp 1       MMA issuer     S_t = Q_t K_j^T (SS) ; O_t += P_t V_j (TS, P from TMEM, V MN-major), issue order
//                               QK0(0) QK1(0) | PV0(j) QK0(j+1) PV1(j) QK1(j+1) | ...  so while group t does its softmax the
//                               pipe runs the other tile's PV and QK
//   warps 2..9   softmax group 0 (tile 0; warp pair (w, w+4) = column halves of the same 32 rows)    warps 10..17  group 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512), fp32; P_t overwrites the first 64 columns of S_t
// as packed bf16 (here: per column half, at the start of that half's own score columns).  K and V are fetched once per 256 queries.
//
// Two refinements of the S -> P -> PV -> QK chain:
//   * P leaves in kPC instalments with their own mbarriers, so the first k steps of P V overlap the remaining exponentials;
//   * kFixed: a caller-promised bound of the scaled scores (QK-RMSNorm) replaces the running row max -- no max pass over S, no
//     exchange between the two threads of a row, no O rescaling.
// kSp (sequence parallelism): the epilogue ships O to the row owners' buffers with TMA tile stores (NVLink for peers).
#pragma once
#include "attn_common.cuh"

namespace vcb {

constexpr int kAttn3Threads = 576;          // TMA warp + MMA warp + 2 tiles x 8 softmax warps
constexpr int kAttn3Slots = 4;              // K/V ring of 32 KB slots
constexpr int kAttn3SmemBytes = (2 + kAttn3Slots) * kSlotBytes + 1024 + 256 + 4096;

// Sequence-parallel output maps (kSp): m[r] = rank r's [rows_per_rank, ldo] output buffer (all columns), box 64 columns x
// 32 rows, 128-byte swizzle.  Each softmax warp stages its 32 x 64 piece of O in the (by then idle) Q tile and ships it with
// one TMA tile store per owner; rows outside an owner's extent are clipped by the TMA unit.
template <bool kSp>
struct AttnSpMapsT { CUtensorMap m[8]; };
template <>
struct AttnSpMapsT<false> { int unused; };

// kPC: number of instalments (2 or 4) in which a softmax thread publishes the P of its 64 key columns per K/V tile
// kFixed: the caller guarantees |scaled score| <= p.fixed_max (log2 units) -- true after QK-RMSNorm, where |q| and |k| are
// bounded by the norm scales (layers.py:63-84).  softmax is shift invariant, so exp2(s - fixed_max) needs no running row
// max: the max pass over S, the per-tile exchange between the two threads of a row and the O rescaling all disappear.
template <bool kSp, int kPC = 2, bool kFixed = false>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attn_fwd3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p, const __grid_constant__ AttnSpMapsT<kSp> spm) {
    const int q_pair = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int seqlen = p.seqlens ? min(p.seqlens[b], p.L) : p.L;
    const int q0 = q_pair * 2 * kAttnTile;
    const uint32_t warp = warp_id_uniform();
    const uint32_t lane = lane_id();
    const bool tile1 = (q0 + kAttnTile) < seqlen;              // second tile has at least one valid query
    pdl_launch_dependents();

    if (q0 >= seqlen) {
        // both tiles are padding: zero rows (pad_input)
        pdl_wait();
        if (warp >= 2) {
            const int g = (int)(warp - 2) & 7;
            const int row = q0 + ((int)(warp - 2) >> 3) * kAttnTile + (int)(warp & 3) * 32 + (int)lane;
            if (row < p.L) {
                uint4* dst = reinterpret_cast<uint4*>(attn_out_row(p, b, row) + p.out_col_offset + head * 128 + (g >> 2) * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;                                  // 2 x 32 KB
    uint8_t* smem_kv = smem + 2 * kSlotBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kAttn3Slots) * kSlotBytes);
    uint64_t* q_full = bars;                       // [1]
    uint64_t* kv_full = bars + 1;                  // [slots]
    uint64_t* kv_empty = kv_full + kAttn3Slots;    // [slots]
    uint64_t* s_full = kv_empty + kAttn3Slots;     // [2] per tile
    uint64_t* p_full = s_full + 2;                 // [2 tiles][kPC column chunks]
    uint64_t* o_done = p_full + 2 * kPC;           // [2]
    static_assert(kPC == 2 || kPC == 4, "P instalments");
    constexpr int kCW = 64 / kPC;                  // key columns per instalment and thread
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);
    // [parity][tile][column half][row]: row-max / row-sum exchange between the two threads of a row
    float (*xch)[2][2][128] = reinterpret_cast<float (*)[2][2][128]>(smem + (2 + kAttn3Slots) * kSlotBytes + 256);

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < kAttn3Slots; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            for (int c = 0; c < kPC; ++c) mbar_init(&p_full[kPC * t + c], 8 /* one arrive per softmax warp */);
            mbar_init(&o_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_kv = (seqlen + kAttnTile - 1) / kAttnTile;
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int ntile = tile1 ? 2 : 1;
            mbar_expect_tx(q_full, kSlotBytes * ntile);
            for (int t = 0; t < ntile; ++t) {
                tma_load_3d<false>(&tmap_qkv, q_full, smem_q + t * kSlotBytes, p.q_col + head * 128, q0 + t * kAttnTile, b, kEvictFirst);
                tma_load_3d<false>(&tmap_qkv, q_full, smem_q + t * kSlotBytes + kSlotBytes / 2, p.q_col + head * 128 + 64,
                                   q0 + t * kAttnTile, b, kEvictFirst);
            }
            for (int seq = 0; seq < 2 * n_kv; ++seq) {           // K0 V0 K1 V1 ...
                const int slot = seq % kAttn3Slots;
                const uint32_t ph = (uint32_t)(seq / kAttn3Slots) & 1u;
                const int j = seq >> 1;
                const int col = ((seq & 1) ? p.v_col : p.k_col) + head * 128;
                mbar_wait(&kv_empty[slot], ph ^ 1);
                mbar_expect_tx(&kv_full[slot], kSlotBytes);
                uint8_t* dst = smem_kv + slot * kSlotBytes;
                tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst, col, j * kAttnTile, b, kEvictLast);
                tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst + kSlotBytes / 2, col + 64, j * kAttnTile, b, kEvictLast);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs this loop convergently (descriptors stay in uniform registers); one elected lane issues.
        {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);
            auto slot_of = [](int seq) { return seq % kAttn3Slots; };
            auto wait_kv = [&](int seq) {
                mbar_wait(&kv_full[slot_of(seq)], (uint32_t)(seq / kAttn3Slots) & 1u);
                tc_fence_after();
            };
            auto issue_qk = [&](int t, int j) {                  // S_t = Q_t K_j^T
                const uint32_t qa = smem_u32(smem_q + t * kSlotBytes), ka = smem_u32(smem_kv + slot_of(2 * j) * kSlotBytes);
                const uint64_t qd = make_smem_desc(qa, 16, 1024, kSwizzle128B), kd = make_smem_desc(ka, 16, 1024, kSwizzle128B);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        // +32 B per 16-wide k step inside a 64-wide block, +16 KB for the second block (descriptor units of 16 B)
                        const uint64_t off = (uint64_t)(((ks >> 2) * (kSlotBytes / 2) + (ks & 3) * 32) >> 4);
                        umma_ss<1>(tmem_base + t * 128, qd + off, kd + off, idesc_qk, ks != 0);
                    }
                    umma_commit<1>(&s_full[t]);
                }
                __syncwarp();
            };
            // O_t += P_t V_j in kPC instalments: every softmax thread publishes its P in kPC chunks of 64/kPC keys, so the k steps
            // of chunk 0 (kPC = 2: keys [0,32) and [64,96)) run on the tensor pipe while the later chunks' exponentials are
            // still being computed
            auto issue_pv = [&](int t, int j) {
                const uint32_t va = smem_u32(smem_kv + slot_of(2 * j + 1) * kSlotBytes);
                const uint64_t vd = make_smem_desc(va, kSlotBytes / 2, 1024, kSwizzle128B);
#pragma unroll
                for (int c = 0; c < kPC; ++c) {
                    mbar_wait(&p_full[kPC * t + c], (uint32_t)j & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        constexpr int kPer = kCW / 16;           // 16-key k steps per chunk and column half
#pragma unroll
                        for (int i = 0; i < 2 * kPer; ++i) {
                            const int ks = (i / kPer) * 4 + c * kPer + (i % kPer);
                            umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + (ks >> 2) * 64 + (ks & 3) * 8,   // split P layout
                                    vd + (uint64_t)(ks * (2048 >> 4)), idesc_pv, (j | c | i) != 0);
                        }
                        if (c == kPC - 1) umma_commit<1>(&o_done[t]);
                    }
                    __syncwarp();
                }
            };
            mbar_wait(q_full, 0);
            wait_kv(0);
            issue_qk(0, 0);
            if (tile1) issue_qk(1, 0);
            if (elect_one()) umma_commit<1>(&kv_empty[slot_of(0)]);
            __syncwarp();
            for (int j = 0; j < n_kv; ++j) {
                const bool more = (j + 1) < n_kv;
                wait_kv(2 * j + 1);                               // V_j
                issue_pv(0, j);
                if (more) { wait_kv(2 * j + 2); issue_qk(0, j + 1); }
                if (tile1) issue_pv(1, j);
                if (elect_one()) umma_commit<1>(&kv_empty[slot_of(2 * j + 1)]);    // V_j free once both PVs have run
                __syncwarp();
                if (more) {
                    if (tile1) issue_qk(1, j + 1);
                    if (elect_one()) umma_commit<1>(&kv_empty[slot_of(2 * j + 2)]);   // K_{j+1} free once both QKs have run
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax groups (two threads per row) =====================
        const int t = (int)(warp - 2) >> 3;                       // tile / group index
        const int half = ((int)(warp - 2) & 7) >> 2;              // key-column half handled by this thread
        const uint32_t quarter = warp & 3;                        // TMEM lane quarter (== warp id % 4)
        const int rit = (int)quarter * 32 + (int)lane;            // row inside the tile
        const int row = q0 + t * kAttnTile + rit;
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const uint32_t s_addr = tmem_base + lane_addr + t * 128 + half * 64;         // my 64 score columns
        // my 32 packed-P columns overlay the START OF MY OWN score columns (not the partner's): P of key columns
        // [0,64) lives at S_t + [0,32), P of key columns [64,128) at S_t + [64,96)
        const uint32_t p_addr = tmem_base + lane_addr + t * 128 + half * 64;
        const uint32_t o_addr = tmem_base + lane_addr + 256 + t * 128 + half * 64;   // my 64 output columns
        const uint32_t bar_id = 1 + t * 4 + quarter;              // named barrier of this warp pair (64 threads)
        __nv_bfloat16* dst = attn_out_row(p, b, row) + p.out_col_offset + head * 128 + half * 64;
        if (t == 1 && !tile1) {
            if (row < p.L) {
#pragma unroll
                for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
            }
        } else {
            [[maybe_unused]] float m_run = -INFINITY;
            float l_run = 0.f;
            const float sc = p.scale_log2;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&s_full[t], (uint32_t)j & 1u);
                tc_fence_after();
                const int kv_left = seqlen - j * kAttnTile - half * 64;     // my columns >= kv_left are padding
                float m_new = p.fixed_max;
                if constexpr (!kFixed) {
                // pass 1: row max of my 64 columns (scores are re-read from TMEM in pass 2: keeps the live set at one chunk)
                float m_tile = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t sr[32];
                    tmem_ld_x32(s_addr + c * 32, sr);
                    tmem_wait_ld();
                    if (kv_left >= 64) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, __uint_as_float(sr[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < kv_left) m_tile = fmaxf(m_tile, __uint_as_float(sr[i]));
                    }
                }
                // row max over both halves: exchange with the partner thread (other warp, same lane)
                xch[j & 1][t][half][rit] = m_tile;
                named_bar_sync(bar_id, 64);
                m_tile = fmaxf(m_tile, xch[j & 1][t][half ^ 1][rit]) * sc;   // scaled log2 units
                const bool grow = (m_tile - m_run) > kRescaleThreshold;
                m_new = grow ? m_tile : m_run;
                const float alpha = grow ? ex2_approx(m_run - m_new) : 1.0f;
                if (j > 0 && __any_sync(0xffffffffu, grow)) {
                    mbar_wait(&o_done[t], (uint32_t)(j - 1) & 1u);
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_x32(o_addr + c * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_x32(o_addr + c * 32, o);
                    }
                }
                l_run *= alpha;
                m_run = m_new;
                }
                {
                    const uint64_t sc2 = pack_f32x2(sc, sc), nm2 = pack_f32x2(-m_new, -m_new);
                    uint64_t acc_a = pack_f32x2(0.f, 0.f), acc_b = acc_a;
#pragma unroll
                    for (int c = 0; c < kPC; ++c) {
                        uint32_t sr[kCW];
                        if constexpr (kCW == 32) tmem_ld_x32(s_addr + c * kCW, sr);
                        else tmem_ld_x16(s_addr + c * kCW, sr);
                        tmem_wait_ld();
                        if (kv_left < 64) {
#pragma unroll
                            for (int i = 0; i < kCW; ++i)
                                if (c * kCW + i >= kv_left) sr[i] = 0xff800000u;     // -inf -> p = 0
                        }
                        uint32_t pk[kCW / 2];
#pragma unroll
                        for (int i = 0; i < kCW; i += 2) {
                            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), sc2, nm2);
                            float p0, p1;
                            unpack_f32x2(x2, p0, p1);
                            p0 = ex2_approx(p0);
                            p1 = ex2_approx(p1);
                            if ((i >> 1) & 1) acc_b = add_f32x2(acc_b, pack_f32x2(p0, p1));
                            else acc_a = add_f32x2(acc_a, pack_f32x2(p0, p1));
                            pk[i >> 1] = pack_bf16x2(p0, p1);
                        }
                        // P chunk c overwrites columns [c*kCW/2, +kCW/2) of my own score region: already consumed (c' <= c)
                        if constexpr (kCW == 32) tmem_st_x16(p_addr + c * (kCW / 2), pk);
                        else tmem_st_x8(p_addr + c * (kCW / 2), pk);
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&p_full[kPC * t + c]);   // one mbarrier arrive per warp and chunk
                    }
                    float a0, a1, b0, b1;
                    unpack_f32x2(acc_a, a0, a1);
                    unpack_f32x2(acc_b, b0, b1);
                    l_run += (a0 + b0) + (a1 + b1);
                }
            }
            // ---- epilogue: total row sum = my half + partner's half ----
            xch[n_kv & 1][t][half][rit] = l_run;
            named_bar_sync(bar_id, 64);
            const float inv_l = 1.0f / (l_run + xch[n_kv & 1][t][half ^ 1][rit]);
            mbar_wait(&o_done[t], (uint32_t)(n_kv - 1) & 1u);
            tc_fence_after();
            const bool valid = row < seqlen;
            // kSp: a 32-row piece that straddles two owners' row ranges (rows_per_rank % 32 != 0) keeps the per-thread
            // stores through the peer-mapped pointers -- TMA tile stores take no negative start coordinate
            [[maybe_unused]] const int piece_r0 = q0 + t * kAttnTile + (int)quarter * 32;
            [[maybe_unused]] const bool staged = kSp && (piece_r0 / p.sp_rows == min(piece_r0 + 31, p.L - 1) / p.sp_rows);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t o[32];
                __syncwarp();
                tmem_ld_x32(o_addr + c * 32, o);
                tmem_wait_ld();
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = 0u;
                }
                if (staged || row < p.L) {
                    // kSp: my row of the warp's staging piece inside the idle Q tile (o_done => every QK of this tile has run)
                    [[maybe_unused]] uint8_t* stg_row = smem_q + t * kSlotBytes + half * (kSlotBytes / 2) + rit * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
                        u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
                        u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
                        u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
                        if (staged) *reinterpret_cast<uint4*>(stg_row + (((c * 4 + g) ^ (rit & 7)) << 4)) = u;
                        else *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = u;
                    }
                }
            }
            if constexpr (kSp) {
                // the warp's 32 rows x 64 columns leave as one TMA tile store into the owner's buffer
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && staged && piece_r0 < p.L) {
                    const uint8_t* piece = smem_q + t * kSlotBytes + half * (kSlotBytes / 2) + (int)quarter * 32 * 128;
                    const int owner = piece_r0 / p.sp_rows;
                    tma_store_2d(&spm.m[owner], piece, p.out_col_offset + head * 128 + half * 64, piece_r0 - owner * p.sp_rows);
                    tma_store_commit();
                    tma_store_wait_all();
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
