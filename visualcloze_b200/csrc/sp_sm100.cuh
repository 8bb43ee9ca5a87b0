// sp_sm100.cuh -- cross-GPU phase barrier of the sequence-parallel mode (include/vcb200.h: vcb_sp_barrier).
//
// The producers before the barrier (QKV GEMM / attention epilogues) write into peer memory with ordinary global stores
// that leave the GPU over NVLink.  Stream order makes those stores "performed" before this kernel starts; a system-scope
// fence + release store then publishes this rank's epoch in every peer's flag array, and an acquire spin waits for every
// peer's epoch in ours.  Flags only grow, so a fast rank publishing epoch e+1 cannot confuse a slow rank waiting for e.
#pragma once
#include "vcb_common.cuh"

namespace vcb {

struct SpFlags { int* f[8]; };     // f[r] = rank r's flag array as mapped in this process

VCB_DEVICE void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
VCB_DEVICE int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
VCB_DEVICE unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void sp_barrier_kernel(SpFlags flags, int world, int rank, int epoch, int* err, unsigned long long timeout_ns) {
    // Launched with programmatic stream serialization: this CTA is resident while the producer kernel still runs and lets the
    // consumer kernel launch right away (it parks in its own griddepcontrol.wait until this grid has completed), so neither
    // launch latency sits on the critical path.  griddepcontrol.wait returns once the producer grid has completed and its
    // memory operations -- including the stores to peer memory -- have been performed.
    pdl_launch_dependents();
    pdl_wait();
    const int t = (int)threadIdx.x;
    if (t >= world || t == rank) return;
    __threadfence_system();
    st_release_sys(flags.f[t] + rank, epoch);                 // "rank has finished phase `epoch`" -> peer t
    const int* mine = flags.f[rank] + t;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mine) - epoch < 0) {                // wrap-safe "peer t has not reached epoch yet"
        if (global_timer_ns() - t0 > timeout_ns) {            // a lost peer must not hang the GPU: report and go on
            atomicExch(err, epoch);
            break;
        }
    }
}

}  // namespace vcb
