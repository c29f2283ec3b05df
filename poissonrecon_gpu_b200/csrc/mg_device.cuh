// Device side of the multi-GPU flag barrier (see MgState in common.cuh).
#pragma once
#include "common.cuh"

namespace prb {

__device__ __forceinline__ void mg_store_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned mg_load_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// executed by ONE thread after every local write that peers will read has been ordered before it
// (grid sync / kernel boundary): announce `epoch` to all peers, wait until all peers announced it
__device__ __forceinline__ void mg_signal_wait(const MgDev& mg, unsigned epoch) {
    if (mg.hdr->error) return;
    __threadfence_system();
    for (int r = 0; r < mg.world; r++)
        if (r != mg.rank) mg_store_release_sys(&mg.peerHdr[r]->flags[mg.rank][0], epoch);
    long long t0 = clock64();
    for (int r = 0; r < mg.world; r++) {
        if (r == mg.rank) continue;
        while ((int)(mg_load_acquire_sys(&mg.hdr->flags[r][0]) - epoch) < 0) {
            if (clock64() - t0 > mg.spinCycles) { mg.hdr->error = 1; return; }
        }
    }
    __threadfence_system();
}

}  // namespace prb
