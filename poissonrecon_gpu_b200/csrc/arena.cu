// Per-context device arena behind DBuf (common.cuh).
//
// Every buffer of a reconstruction lives on ONE stream, so a block freed by the host can be handed
// out again immediately: stream order already guarantees that the previous user's kernels finish
// first.  That makes a plain host-side free-list allocator over a few large cudaMalloc slabs
// sufficient -- and, unlike the driver's stream-ordered pool (cudaMallocAsync), fully
// deterministic: after the first run of a given size the context never calls the driver's
// allocator again (the pool occasionally re-mapped hundreds of MB in the middle of a run, which
// showed up as 100-400 ms stalls of single reconstructions).
#include <cuda_runtime.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "common.cuh"

namespace prb {

namespace {
constexpr size_t kAlign = 512;
constexpr size_t kMinSlab = (size_t)256 << 20;

struct Arena {
    std::vector<std::pair<char*, size_t>> slabs;
    std::map<char*, size_t> freeBlocks;          // address -> bytes (coalesced)
    std::unordered_map<void*, size_t> live;      // allocation -> bytes
    size_t reserved = 0, inUse = 0, peak = 0;
    long driverCalls = 0;

    void* take(size_t bytes) {
        // best fit
        auto best = freeBlocks.end();
        for (auto it = freeBlocks.begin(); it != freeBlocks.end(); ++it)
            if (it->second >= bytes && (best == freeBlocks.end() || it->second < best->second)) best = it;
        if (best == freeBlocks.end()) return nullptr;
        char* p = best->first;
        size_t sz = best->second;
        freeBlocks.erase(best);
        if (sz > bytes) freeBlocks.emplace(p + bytes, sz - bytes);
        live.emplace(p, bytes);
        inUse += bytes;
        peak = std::max(peak, inUse);
        return p;
    }
    int grow(size_t bytes) {
        // a new slab at least as large as everything reserved so far: the slab count stays logarithmic
        size_t want = std::max(std::max(bytes, kMinSlab), std::min(reserved, (size_t)8 << 30));
        want = (want + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        char* p = nullptr;
        cudaError_t e = cudaMalloc((void**)&p, want);
        if (e != cudaSuccess && want > bytes) {          // not enough room for the generous size: ask for the minimum
            cudaGetLastError();
            want = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
            e = cudaMalloc((void**)&p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("device arena: cudaMalloc of " + std::to_string(want >> 20) + " MiB failed (" + cudaGetErrorString(e) + "), " +
                      std::to_string(reserved >> 20) + " MiB already reserved");
            return PRB_ERR_NOMEM;
        }
        driverCalls++;
        slabs.emplace_back(p, want);
        reserved += want;
        give(p, want);
        return PRB_OK;
    }
    void give(char* p, size_t bytes) {
        auto next = freeBlocks.lower_bound(p);
        // merge with the previous block when adjacent AND in the same slab (slabs are never adjacent-merged
        // across cudaMalloc boundaries: a block must not straddle two allocations)
        if (next != freeBlocks.begin()) {
            auto prev = std::prev(next);
            if (prev->first + prev->second == p && same_slab(prev->first, p)) { p = prev->first; bytes += prev->second; freeBlocks.erase(prev); }
        }
        if (next != freeBlocks.end() && p + bytes == next->first && same_slab(p, next->first)) { bytes += next->second; freeBlocks.erase(next); }
        freeBlocks.emplace(p, bytes);
    }
    bool same_slab(const char* a, const char* b) const {
        for (auto& s : slabs)
            if (a >= s.first && a < s.first + s.second) return b >= s.first && b < s.first + s.second;
        return false;
    }
    void destroy() {
        for (auto& s : slabs) cudaFree(s.first);
        slabs.clear(); freeBlocks.clear(); live.clear();
        reserved = inUse = 0;
    }
};

std::mutex g_mu;
std::vector<std::pair<cudaStream_t, Arena*>> g_arenas;

Arena* find(cudaStream_t st) {
    for (auto& a : g_arenas)
        if (a.first == st) return a.second;
    return nullptr;
}
}  // namespace

void arena_register(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!find(st)) g_arenas.emplace_back(st, new Arena());
}

void arena_unregister(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_arenas.size(); i++)
        if (g_arenas[i].first == st) {
            g_arenas[i].second->destroy();
            delete g_arenas[i].second;
            g_arenas.erase(g_arenas.begin() + (long)i);
            return;
        }
}

int arena_alloc(void** out, size_t bytes, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    Arena* a = find(st);
    if (!a) { set_error("device arena: stream not registered"); return PRB_ERR_STATE; }
    bytes = (bytes + kAlign - 1) & ~(kAlign - 1);
    void* p = a->take(bytes);
    if (!p) {
        PRB_TRY(a->grow(bytes));
        p = a->take(bytes);
        if (!p) { set_error("device arena: internal error"); return PRB_ERR_NOMEM; }
    }
    *out = p;
    return PRB_OK;
}

void arena_free(void* p, cudaStream_t st) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    Arena* a = find(st);
    if (!a) return;
    auto it = a->live.find(p);
    if (it == a->live.end()) return;
    size_t bytes = it->second;
    a->live.erase(it);
    a->inUse -= bytes;
    a->give((char*)p, bytes);
}

void arena_stats(cudaStream_t st, size_t* reserved, size_t* peak, long* driverCalls) {
    std::lock_guard<std::mutex> lk(g_mu);
    Arena* a = find(st);
    if (reserved) *reserved = a ? a->reserved : 0;
    if (peak) *peak = a ? a->peak : 0;
    if (driverCalls) *driverCalls = a ? a->driverCalls : 0;
}

}  // namespace prb
