// attn4_sm100.cuh -- PERSISTENT joint attention forward: one CTA per SM, equal contiguous shares of the (unit, key tile) space.
//
// Same per-tile pipeline as attn3_sm100.cuh (two 128-query tiles per unit, row-split softmax over 16 warps, P in TMEM, TS-mode
// P V, two-instalment P, optional fixed-reference softmax) -- what changes is the schedule.  attn3 launches one CTA per
// (query pair, head, sample): at cfg B that is 16 x 24 = 384 CTAs (372 pair-units of work) on 148 SMs = 2.51 waves executed as
// 3 (84 % before anything else), and every CTA pays its own prologue (TMEM alloc, barrier init, first Q / K fetch) and drain.
// Here the grid is min(#SMs, ...) CTAs.  Full rounds of units are dealt out exactly like the per-pair grid (same L2 behaviour, no
// per-CTA prologue / drain between them); the units of the partial last round are cut along the key tiles into equal contiguous
// shares (see AttnSched).  A CTA's work is a sequence of SEGMENTS (unit, key tiles [kv0, kv1)):
//   * a segment covering all key tiles is finished locally (normalise, store);
//   * a unit cut by a share boundary is produced by several CTAs.  The piece holding key tile 0 -- always the LAST segment of its
//     CTA -- is the finaliser; every other piece -- always the FIRST phase-2 segment of its CTA, which waits on nothing -- dumps its
//     un-normalised (O, l[, m]) to its workspace slot and raises its flag.  The finaliser folds them in its epilogue:
//         fixed-reference softmax:  O = sum O_i, l = sum l_i            (all pieces share the reference exponent)
//         online max:               m = max m_i, O = sum O_i 2^(m_i - m), l likewise
//     Waiters only wait on CTAs with a larger index whose contribution is their first piece: deadlock-free with all CTAs
//     co-resident (grid <= #SMs, one CTA per SM).  Spin loops carry a watchdog that traps instead of hanging the GPU.
// Contract: models/math.py:63-99 for unpadded batches (seqlens == nullptr; right-padded batches and the sequence-parallel
// routing stay on attn3).
#pragma once
#include "attn3_sm100.cuh"
#include "gemm_sm100.cuh"      // ld_acquire_gpu / st_release_gpu

namespace vcb {

constexpr int kAttn4MaxSegs = 64;                            // segments per CTA (the host falls back to attn3 beyond that)
constexpr int kAttn4SmemBytes = kAttn3SmemBytes + 4096 + kAttn4MaxSegs * 32;   // + two more row-exchange buffers + the segment list
constexpr int kAttn4SlotFloats = 2 * 128 * 128 + 2 * 2 * 128 + 2 * 128;   // O [2][32][128] float4 | l [2][2][128] | m [2][128]

struct AttnSkParams {
    float* ws;          // [gridDim.x][kAttn4SlotFloats]
    int* flags;         // [gridDim.x]
    int epoch;          // value that marks "this launch's partial is ready"
    int no_split;       // experiment switch: deal the units of the partial last round out whole (no key-tile split, no partials)
    unsigned long long* timeline;   // optional [gridDim.x][kAttn4MaxSegs + 2] globaltimer stamps: start, end of every segment (debug)
};

// Schedule.  Units (sample, head, query pair) in that order; n_units = samples * heads * pairs.
//   phase 1  the R = n_units / G full rounds exactly like the one-CTA-per-pair grid: round r gives unit r * G + c to CTA c, so the CTAs
//            that run concurrently stream the same ~G / pairs heads' K / V through L2.  (A first version cut the whole (unit, key tile)
//            space into G contiguous ranges: every head's K / V was live at once -- 91 MB at L = 7424 -- and each key-tile step ran
//            20 % slower than in the per-pair grid; measured 534 vs 469 us.  Same lesson as the GEMM's stream-K: split only the tail.)
//   phase 2  the rem = n_units % G units of the partial last round are cut along the key tiles into G2 equal contiguous ranges of the
//            flattened (unit, key tile) space, one per CTA c < G2 (G2 = min(G, steps / 4): every range holds >= 4 key-tile steps).
struct AttnSched {
    int n_pairs, n_kv, n_units, G;
    int R, rem, G2;
    long long U;                                   // key-tile steps of phase 2
    VCB_DEVICE void init(int n_heads, int n_qt, int grid, int no_split = 0) {
        n_pairs = (n_qt + 1) / 2; n_kv = n_qt; n_units = n_heads * n_pairs; G = grid;
        R = n_units / G; rem = n_units - R * G;
        U = (long long)rem * n_kv;
        const long long g2 = U / 4;
        G2 = rem == 0 ? 0 : (int)(g2 < 1 ? 1 : (g2 < G ? g2 : G));
        if (no_split && rem > 0) G2 = rem;             // one whole unit per CTA c < rem: boundary(c) = c * n_kv
    }
    // start of CTA c's phase-2 range (c == G2: the end of the space)
    VCB_DEVICE void boundary(int c, int& unit, int& kv) const {
        const long long x = U * c / G2;
        const int ur = (int)(x / n_kv);
        unit = R * G + ur;
        kv = (int)(x - (long long)ur * n_kv);
    }
};

struct AttnSeg { int unit, kv0, kv1, n_parts; };

// enumerates CTA c's segments: phase 1 units, then the pieces of its phase-2 range.  n_parts (for a cut unit's finaliser) = the
// CTAs right after this one whose range starts inside the unit.
struct AttnSegIter {
    const AttnSched& s;
    int c, r, u, u_s, u_e, k_s, k_e;
    VCB_DEVICE AttnSegIter(const AttnSched& sched, int cta) : s(sched), c(cta), r(0) {
        if (c < s.G2) { s.boundary(c, u_s, k_s); s.boundary(c + 1, u_e, k_e); } else { u_s = 1; u_e = 0; k_s = k_e = 0; }
        u = u_s;
    }
    VCB_DEVICE bool next(AttnSeg& g) {
        if (r < s.R) { g.unit = r * s.G + c; g.kv0 = 0; g.kv1 = s.n_kv; g.n_parts = 0; ++r; return true; }
        while (u <= u_e && u < s.n_units) {
            const int cur = u++;
            const int kv0 = (cur == u_s) ? k_s : 0;
            const int kv1 = (cur == u_e) ? k_e : s.n_kv;
            if (kv0 >= kv1) continue;
            g.unit = cur; g.kv0 = kv0; g.kv1 = kv1; g.n_parts = 0;
            if (kv0 == 0 && kv1 < s.n_kv) {
                for (int pc = c + 1; pc < s.G2; ++pc) {
                    int pu, pk;
                    s.boundary(pc, pu, pk);
                    if (pu != cur) break;
                    ++g.n_parts;
                }
            }
            return true;
        }
        return false;
    }
};
// one entry of a CTA's segment list, computed once by one thread (keeps the 64-bit schedule arithmetic out of the role loops)
struct AttnSegEntry { int b, head, q0, kv0, kv1, tile1, n_parts, unit; };
static_assert(sizeof(AttnSegEntry) == 32, "segment entry");

// CTA `cta`'s segment list -> shared memory; returns the count.  Out of line on purpose: its locals (iterator structs, 64-bit schedule
// arithmetic) get their own stack frame instead of forcing one on the kernel.
__device__ __noinline__ int attn4_build_segments(const AttnSched& sched, int cta, int heads, int seqlen, AttnSegEntry* segs) {
    AttnSegIter it(sched, cta);
    AttnSeg g;
    int n = 0;
    while (n < kAttn4MaxSegs && it.next(g)) {
        AttnSegEntry e;
        const int hi = g.unit / sched.n_pairs;
        e.b = hi / heads;
        e.head = hi - e.b * heads;
        e.q0 = (g.unit - hi * sched.n_pairs) * 2 * kAttnTile;
        e.kv0 = g.kv0; e.kv1 = g.kv1; e.unit = g.unit; e.n_parts = g.n_parts;
        e.tile1 = (e.q0 + kAttnTile) < seqlen ? 1 : 0;
        segs[n++] = e;
    }
    return n;
}

// mbarrier wait with a watchdog: ~seconds of failed try_waits trap the kernel instead of hanging the GPU box
VCB_DEVICE unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Segment-level waits only (q_full / q_empty / o_free): the per-key-tile waits use the plain mbar_wait.  ncu (source page) of the
// first version, where EVERY wait was this loop: the compiler put a YIELD in front of each try_wait, a failed attempt then returned
// after ~70 ns instead of parking the warp, and the s_full wait alone executed 4.4 M try_waits per launch (24 per key-tile step and
// warp) against 0.18 M in the per-pair kernel.
VCB_DEVICE void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// 576 threads: registers are allocated in units of 4 warps, so 18 warps cost as much as 20 and the cap is 96 per thread (a
// __maxnreg__(104) build compiled but could not be launched).  The few spills ptxas reports sit on the per-segment path, not in
// the key-tile loop.
template <bool kFixed>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attn_fwd4_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p, const AttnSkParams skp) {
    constexpr int kPC = 2, kCW = 32;
    const uint32_t warp = warp_id_uniform();
    const uint32_t lane = lane_id();
    const int cta = blockIdx.x, G = gridDim.x;
    pdl_launch_dependents();

    const int n_qt = (p.L + kAttnTile - 1) / kAttnTile;
    AttnSched sched;
    sched.init(p.B * p.H, n_qt, G, skp.no_split);
    const int n_kv_all = sched.n_kv;
    const int seqlen = p.L;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;                                  // 2 x 32 KB
    uint8_t* smem_kv = smem + 2 * kSlotBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kAttn3Slots) * kSlotBytes);
    uint64_t* q_full = bars;                       // [1]   one phase per segment
    uint64_t* q_empty = bars + 1;                  // [1]   all Q K^T of the segment have run: the Q tiles may be overwritten
    uint64_t* kv_full = bars + 2;                  // [slots]
    uint64_t* kv_empty = kv_full + kAttn3Slots;    // [slots]
    uint64_t* s_full = kv_empty + kAttn3Slots;     // [2] per tile
    uint64_t* p_full = s_full + 2;                 // [2 tiles][kPC column chunks]
    uint64_t* o_done = p_full + 2 * kPC;           // [2]
    uint64_t* o_free = o_done + 2;                 // [2]   the segment's epilogue has read O_t out of TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
    // [buffer][tile][column half][row]: buffers 0/1 = per-key-tile row max (parity of the running step), 2/3 = row sum (parity of
    // the segment count)
    float (*xch)[2][2][128] = reinterpret_cast<float (*)[2][2][128]>(smem + (2 + kAttn3Slots) * kSlotBytes + 256);
    AttnSegEntry* segs = reinterpret_cast<AttnSegEntry*>(smem + (2 + kAttn3Slots) * kSlotBytes + 256 + 8192);
    int* n_segs_smem = reinterpret_cast<int*>(tmem_slot + 1);

    if (warp == 2 && lane == 0) *n_segs_smem = attn4_build_segments(sched, cta, p.H, seqlen, segs);
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < kAttn3Slots; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            for (int c = 0; c < kPC; ++c) mbar_init(&p_full[kPC * t + c], 8 /* one arrive per softmax warp */);
            mbar_init(&o_done[t], 1);
            mbar_init(&o_free[t], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int n_segs = *n_segs_smem;
    if (skp.timeline != nullptr && warp == 2 && lane == 0) skp.timeline[(long long)cta * (kAttn4MaxSegs + 2)] = globaltimer_ns();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int seq = 0;
            for (int segn = 0; segn < n_segs; ++segn) {
                const int b = segs[segn].b, head = segs[segn].head, q0 = segs[segn].q0, kv0 = segs[segn].kv0, kv1 = segs[segn].kv1;
                const bool tile1 = segs[segn].tile1 != 0;
                if (segn > 0) mbar_wait_wd(q_empty, (uint32_t)(segn - 1) & 1u);
                const int ntile = tile1 ? 2 : 1;
                mbar_expect_tx(q_full, kSlotBytes * ntile);
                for (int t = 0; t < ntile; ++t) {
                    tma_load_3d<false>(&tmap_qkv, q_full, smem_q + t * kSlotBytes, p.q_col + head * 128, q0 + t * kAttnTile, b, kEvictFirst);
                    tma_load_3d<false>(&tmap_qkv, q_full, smem_q + t * kSlotBytes + kSlotBytes / 2, p.q_col + head * 128 + 64,
                                       q0 + t * kAttnTile, b, kEvictFirst);
                }
                for (int e = 0; e < 2 * (kv1 - kv0); ++e, ++seq) {       // K V K V ...
                    const int slot = seq % kAttn3Slots;
                    const uint32_t ph = (uint32_t)(seq / kAttn3Slots) & 1u;
                    const int j = kv0 + (e >> 1);
                    const int col = ((e & 1) ? p.v_col : p.k_col) + head * 128;
                    mbar_wait(&kv_empty[slot], ph ^ 1);
                    mbar_expect_tx(&kv_full[slot], kSlotBytes);
                    uint8_t* dst = smem_kv + slot * kSlotBytes;
                    tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst, col, j * kAttnTile, b, kEvictLast);
                    tma_load_3d<false>(&tmap_qkv, &kv_full[slot], dst + kSlotBytes / 2, col + 64, j * kAttnTile, b, kEvictLast);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp convergent; one elected lane issues) =====================
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);
        int base = 0;
        int steps0 = 0, steps1 = 0;         // running key-tile steps per query tile (parity of s_full / p_full / o_done)
        int segc0 = 0, segc1 = 0;           // segments in which the tile was active (parity of o_free)
        auto slot_of = [](int seq) { return seq % kAttn3Slots; };
        auto wait_kv = [&](int seq) {
            mbar_wait(&kv_full[slot_of(seq)], (uint32_t)(seq / kAttn3Slots) & 1u);
            tc_fence_after();
        };
        auto issue_qk = [&](int t, int seq) {                  // S_t = Q_t K^T, K in ring entry `seq`
            const uint32_t qa = smem_u32(smem_q + t * kSlotBytes), ka = smem_u32(smem_kv + slot_of(seq) * kSlotBytes);
            const uint64_t qd = make_smem_desc(qa, 16, 1024, kSwizzle128B), kd = make_smem_desc(ka, 16, 1024, kSwizzle128B);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t off = (uint64_t)(((ks >> 2) * (kSlotBytes / 2) + (ks & 3) * 32) >> 4);
                    umma_ss<1>(tmem_base + t * 128, qd + off, kd + off, idesc_qk, ks != 0);
                }
                umma_commit<1>(&s_full[t]);
            }
            __syncwarp();
        };
        // O_t (+)= P_t V, V in ring entry `seq`; first: overwrite O_t; par: parity of tile t's running step (plain values only:
        // anything address-taken here lands in local memory, and with the whole L1 carved out as shared memory a local load is an L2
        // round trip on the S -> P -> PV critical path -- ncu showed 63 K of them and a 20 % slower key-tile step)
        auto issue_pv = [&](int t, int seq, bool first, uint32_t par) {
            const uint32_t va = smem_u32(smem_kv + slot_of(seq) * kSlotBytes);
            const uint64_t vd = make_smem_desc(va, kSlotBytes / 2, 1024, kSwizzle128B);
#pragma unroll
            for (int c = 0; c < kPC; ++c) {
                mbar_wait(&p_full[kPC * t + c], par);
                tc_fence_after();
                if (elect_one()) {
                    constexpr int kPer = kCW / 16;
#pragma unroll
                    for (int i = 0; i < 2 * kPer; ++i) {
                        const int ks = (i / kPer) * 4 + c * kPer + (i % kPer);
                        umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + (ks >> 2) * 64 + (ks & 3) * 8,
                                vd + (uint64_t)(ks * (2048 >> 4)), idesc_pv, (!first || (c | i) != 0) ? 1u : 0u);
                    }
                    if (c == kPC - 1) umma_commit<1>(&o_done[t]);
                }
                __syncwarp();
            }
        };
        for (int segn = 0; segn < n_segs; ++segn) {
            const bool tile1 = segs[segn].tile1 != 0;
            const int n = segs[segn].kv1 - segs[segn].kv0;
            mbar_wait_wd(q_full, (uint32_t)segn & 1u);
            tc_fence_after();
            wait_kv(base);
            issue_qk(0, base);
            if (tile1) issue_qk(1, base);
            if (elect_one()) {
                umma_commit<1>(&kv_empty[slot_of(base)]);
                if (n == 1) umma_commit<1>(q_empty);
            }
            __syncwarp();
            for (int jj = 0; jj < n; ++jj) {
                const bool more = (jj + 1) < n;
                const int ev = base + 2 * jj + 1, ek = base + 2 * jj + 2;
                wait_kv(ev);                                      // V
                if (jj == 0 && segc0 > 0) { mbar_wait_wd(&o_free[0], (uint32_t)(segc0 - 1) & 1u); tc_fence_after(); }
                issue_pv(0, ev, jj == 0, (uint32_t)steps0 & 1u);
                ++steps0;
                if (more) { wait_kv(ek); issue_qk(0, ek); }
                if (tile1) {
                    if (jj == 0 && segc1 > 0) { mbar_wait_wd(&o_free[1], (uint32_t)(segc1 - 1) & 1u); tc_fence_after(); }
                    issue_pv(1, ev, jj == 0, (uint32_t)steps1 & 1u);
                    ++steps1;
                }
                if (elect_one()) umma_commit<1>(&kv_empty[slot_of(ev)]);      // V free once both PVs have run
                __syncwarp();
                if (more) {
                    if (tile1) issue_qk(1, ek);
                    if (elect_one()) {
                        umma_commit<1>(&kv_empty[slot_of(ek)]);               // K free once both QKs have run
                        if (jj + 2 == n) umma_commit<1>(q_empty);             // those were the segment's last Q K^T
                    }
                    __syncwarp();
                }
            }
            base += 2 * n;
            ++segc0;
            if (tile1) ++segc1;
        }
        __syncwarp();
    } else {
        // ===================== softmax groups (two threads per row) =====================
        const int t = (int)(warp - 2) >> 3;                       // tile / group index
        const int half = ((int)(warp - 2) & 7) >> 2;              // key-column half handled by this thread
        const uint32_t quarter = warp & 3;                        // TMEM lane quarter (== warp id % 4)
        const int rit = (int)quarter * 32 + (int)lane;            // row inside the tile
        const uint32_t lane_addr = (quarter * 32u) << 16;
        const uint32_t s_addr = tmem_base + lane_addr + t * 128 + half * 64;
        const uint32_t p_addr = s_addr;                           // packed P overlays the start of my own score columns
        const uint32_t o_addr = tmem_base + lane_addr + 256 + t * 128 + half * 64;
        const uint32_t bar_id = 1 + t * 4 + quarter;              // named barrier of this warp pair (64 threads)
        const float sc = p.scale_log2;
        int step = 0, segc = 0;                                   // running key-tile steps / active segments of MY tile
        for (int segn = 0; segn < n_segs; ++segn) {
            // only the loop bounds stay live across the key-tile loop; the epilogue re-reads the entry from shared memory
            const int kv0 = segs[segn].kv0, kv1 = segs[segn].kv1;
            const bool contributor = kv0 > 0;                     // not the piece holding key tile 0: dump partials
            const bool active = (t == 0) || segs[segn].tile1 != 0;
            if (active) {
                [[maybe_unused]] float m_run = -INFINITY;
                float l_run = 0.f;
                // loop state kept minimal (the kernel sits at the 96-register cap: a spilled loop constant is a local load -- an L2
                // round trip with L1 carved out as shared memory -- between s_full and the first tcgen05.ld of EVERY step)
                int kv_left = seqlen - kv0 * kAttnTile - half * 64;             // my columns >= kv_left are padding
                for (int n_left = kv1 - kv0; n_left > 0; --n_left, ++step, kv_left -= kAttnTile) {
                    mbar_wait(&s_full[t], (uint32_t)step & 1u);
                    tc_fence_after();
                    float m_new = p.fixed_max;
                    if constexpr (!kFixed) {
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
                        xch[step & 1][t][half][rit] = m_tile;
                        named_bar_sync(bar_id, 64);
                        m_tile = fmaxf(m_tile, xch[step & 1][t][half ^ 1][rit]) * sc;   // scaled log2 units
                        const bool grow = (m_tile - m_run) > kRescaleThreshold;
                        m_new = grow ? m_tile : m_run;
                        const float alpha = grow ? ex2_approx(m_run - m_new) : 1.0f;
                        if (n_left != kv1 - kv0 && __any_sync(0xffffffffu, grow)) {
                            mbar_wait(&o_done[t], (uint32_t)(step - 1) & 1u);
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
                            tmem_ld_x32(s_addr + c * kCW, sr);
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
                            tmem_st_x16(p_addr + c * (kCW / 2), pk);
                            tmem_wait_st();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&p_full[kPC * t + c]);
                        }
                        float a0, a1, b0, b1;
                        unpack_f32x2(acc_a, a0, a1);
                        unpack_f32x2(acc_b, b0, b1);
                        l_run += (a0 + b0) + (a1 + b1);
                    }
                }
                // ---------------- segment epilogue ----------------
                mbar_wait_wd(&o_done[t], (uint32_t)(step - 1) & 1u);
                tc_fence_after();
                const int b = segs[segn].b, head = segs[segn].head;
                const int row = segs[segn].q0 + t * kAttnTile + rit;
                const bool finaliser_partial = !contributor && kv1 < n_kv_all;
                float* my_slot = skp.ws + (long long)cta * kAttn4SlotFloats;
                if (contributor) {
                    // un-normalised O (fp32), my half's row sum and the row's reference exponent -> my workspace slot.
                    // O layout [tile][column quad][row]: the lanes of a warp (consecutive rows) write consecutive 16-byte words
                    uint4* o_ws = reinterpret_cast<uint4*>(my_slot) + (long long)t * 32 * 128 + rit;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        __syncwarp();
                        tmem_ld_x32(o_addr + c * 32, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4)
                            o_ws[(half * 16 + c * 8 + q4) * 128] = make_uint4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]);
                    }
                    my_slot[2 * 128 * 128 + (t * 2 + half) * 128 + rit] = l_run;
                    if constexpr (!kFixed) {
                        if (half == 0) my_slot[2 * 128 * 128 + 4 * 128 + t * 128 + rit] = m_run;
                    }
                } else {
                    // fold the other pieces of this unit (following CTAs whose range starts inside it), then normalise and store
                    // contributors are the CTAs right after this one whose range starts inside the unit (ranges are never
                    // empty: the host sizes the grid so that every CTA owns >= 8 half-iterations)
                    float a_own = 1.0f;
                    [[maybe_unused]] float m_fin = m_run;
                    const int n_parts = finaliser_partial ? segs[segn].n_parts : 0;
                    if (n_parts > 0) {
                        if (lane == 0) {
                            for (int pc = cta + 1; pc <= cta + n_parts; ++pc) {
                                uint32_t spins = 0;
                                while (ld_acquire_gpu(skp.flags + pc) != skp.epoch) {
                                    if (++spins > (1u << 28)) __trap();
                                }
                            }
                        }
                        __syncwarp();
                        const float* ps = skp.ws + (long long)(cta + 1) * kAttn4SlotFloats + 2 * 128 * 128;
                        if constexpr (!kFixed) {
                            for (int i = 0; i < n_parts; ++i) m_fin = fmaxf(m_fin, __ldcg(ps + (long long)i * kAttn4SlotFloats + 4 * 128 + t * 128 + rit));
                            a_own = ex2_approx(m_run - m_fin);
                        }
                        l_run *= a_own;
                        for (int i = 0; i < n_parts; ++i) {
                            float ap = 1.0f;
                            if constexpr (!kFixed) ap = ex2_approx(__ldcg(ps + (long long)i * kAttn4SlotFloats + 4 * 128 + t * 128 + rit) - m_fin);
                            l_run = fmaf(__ldcg(ps + (long long)i * kAttn4SlotFloats + (t * 2 + half) * 128 + rit), ap, l_run);
                        }
                    }
                    xch[2 + (segc & 1)][t][half][rit] = l_run;
                    named_bar_sync(bar_id, 64);
                    const float inv_l = 1.0f / (l_run + xch[2 + (segc & 1)][t][half ^ 1][rit]);
                    const bool valid = row < seqlen;
                    __nv_bfloat16* dst = attn_out_row(p, b, row) + p.out_col_offset + head * 128 + half * 64;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        __syncwarp();
                        tmem_ld_x32(o_addr + c * 32, o);
                        tmem_wait_ld();
                        if (n_parts > 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * a_own);
                            for (int pi = 0; pi < n_parts; ++pi) {
                                const float* pslot = skp.ws + (long long)(cta + 1 + pi) * kAttn4SlotFloats;
                                const float4* po = reinterpret_cast<const float4*>(pslot) + (long long)t * 32 * 128 + rit;
                                float ap = 1.0f;
                                if constexpr (!kFixed) ap = ex2_approx(__ldcg(pslot + 2 * 128 * 128 + 4 * 128 + t * 128 + rit) - m_fin);
#pragma unroll
                                for (int q4 = 0; q4 < 8; ++q4) {
                                    const float4 f = __ldcg(po + (half * 16 + c * 8 + q4) * 128);
                                    o[q4 * 4 + 0] = __float_as_uint(fmaf(f.x, ap, __uint_as_float(o[q4 * 4 + 0])));
                                    o[q4 * 4 + 1] = __float_as_uint(fmaf(f.y, ap, __uint_as_float(o[q4 * 4 + 1])));
                                    o[q4 * 4 + 2] = __float_as_uint(fmaf(f.z, ap, __uint_as_float(o[q4 * 4 + 2])));
                                    o[q4 * 4 + 3] = __float_as_uint(fmaf(f.w, ap, __uint_as_float(o[q4 * 4 + 3])));
                                }
                            }
                        }
                        if (!valid) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = 0u;
                        }
                        if (row < p.L) {
#pragma unroll
                            for (int gq = 0; gq < 4; ++gq) {
                                uint4 u;
                                u.x = pack_bf16x2(__uint_as_float(o[gq * 8 + 0]) * inv_l, __uint_as_float(o[gq * 8 + 1]) * inv_l);
                                u.y = pack_bf16x2(__uint_as_float(o[gq * 8 + 2]) * inv_l, __uint_as_float(o[gq * 8 + 3]) * inv_l);
                                u.z = pack_bf16x2(__uint_as_float(o[gq * 8 + 4]) * inv_l, __uint_as_float(o[gq * 8 + 5]) * inv_l);
                                u.w = pack_bf16x2(__uint_as_float(o[gq * 8 + 6]) * inv_l, __uint_as_float(o[gq * 8 + 7]) * inv_l);
                                *reinterpret_cast<uint4*>(dst + c * 32 + gq * 8) = u;
                            }
                        }
                    }
                }
                // O_t has been read out of TMEM: the next segment's first P V may overwrite it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free[t]);
                ++segc;
            }
            if (contributor) {
                // every softmax thread of the CTA has written (and fenced) its part of the slot -> publish
                __threadfence();
                named_bar_sync(9, 512);
                if (warp == 2 && lane == 0) st_release_gpu(skp.flags + cta, skp.epoch);
            }
            if (skp.timeline != nullptr && warp == 2 && lane == 0)
                skp.timeline[(long long)cta * (kAttn4MaxSegs + 2) + 1 + segn] = globaltimer_ns();
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
