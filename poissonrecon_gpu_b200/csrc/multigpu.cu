// Multi-GPU plumbing: peer-mapped arena (CUDA IPC), epoch-flag barrier, shard plan.
// See the MgState comment in common.cuh for the model.  The reference is single-GPU
// (devID = 0 hard-coded, CG_CUDA.cuh:356); this layer is new.
#include "common.cuh"
#include "mg_device.cuh"
#include <algorithm>
#include <cstring>

namespace prb {

__global__ void k_mg_barrier(MgDev mg, unsigned epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) mg_signal_wait(mg, epoch);
}

int mg_barrier(Context& c) {
    if (!c.mg.active()) return PRB_OK;
    c.mg.epoch++;
    PRB_LAUNCH(c, k_mg_barrier, 1, 32, 0, c.mg.dev(), c.mg.epoch);
    return PRB_OK;
}

struct XchgVals { int v[64]; };
__global__ void k_mg_publish_ints(MgDev mg, int parity, XchgVals vals, int n) {
    const int t = threadIdx.x;
    if (t < n)
        for (int r = 0; r < mg.world; r++) mg.peerHdr[r]->xchg[parity][mg.rank][t] = vals.v[t];
}
int mg_exchange_ints(Context& c, const int* mine, int n, int* all, int stride) {
    if (n < 0 || n > 64) { set_error("mg_exchange_ints: at most 64 values"); return PRB_ERR_ARG; }
    if (stride < 0) stride = n;
    if (!c.mg.active()) { for (int k = 0; k < n; k++) all[k] = mine[k]; return PRB_OK; }
    XchgVals v;
    for (int k = 0; k < 64; k++) v.v[k] = k < n ? mine[k] : 0;
    const int parity = (int)(c.mg.xchgCount++ & 1u);
    PRB_LAUNCH(c, k_mg_publish_ints, 1, 64, 0, c.mg.dev(), parity, v, n);
    PRB_TRY(mg_barrier(c));
    int host[kMaxRanks][64];
    PRB_CUDA(cudaMemcpyAsync(host, &((MgHeader*)c.mg.arena)->xchg[parity][0][0], sizeof(host), cudaMemcpyDeviceToHost, c.stream));
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    for (int r = 0; r < c.mg.world; r++)
        for (int k = 0; k < n; k++) all[r * stride + k] = host[r][k];
    return PRB_OK;
}

// Pull of up to 32 segments from the peers' arenas with the SMs (one launch, all peers' NVLink paths busy at once; a chain of
// cudaMemcpyAsync calls on one stream moves one peer's share after the other: 0.3 GB/ms measured on 8 GPUs).  16-byte words when
// source, destination and length allow, 4-byte words otherwise.
struct PullSegs {
    int n;
    const char* src[32];
    char* dst[32];
    unsigned long long bytes[32], first[33];      // first[k]: first 256-thread tile of segment k (tiles of 4096 x 16 bytes)
};
constexpr unsigned long long kPullTileBytes = 65536;
__global__ void __launch_bounds__(256) k_mg_pull(const __grid_constant__ PullSegs S) {
    const unsigned long long nTiles = S.first[S.n];
    for (unsigned long long t = blockIdx.x; t < nTiles; t += gridDim.x) {
        int k = 0;
        while (k + 1 < S.n && t >= S.first[k + 1]) k++;
        const unsigned long long off = (t - S.first[k]) * kPullTileBytes;
        const unsigned long long len = S.bytes[k] - off < kPullTileBytes ? S.bytes[k] - off : kPullTileBytes;
        const char* src = S.src[k] + off;
        char* dst = S.dst[k] + off;
        if (((((unsigned long long)src) | ((unsigned long long)dst) | len) & 15ull) == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            const unsigned n4 = (unsigned)(len >> 4);
#pragma unroll 4
            for (unsigned i = threadIdx.x; i < n4; i += 256) d4[i] = s4[i];
        } else {
            const unsigned* s1 = reinterpret_cast<const unsigned*>(src);
            unsigned* d1 = reinterpret_cast<unsigned*>(dst);
            const unsigned n1 = (unsigned)(len >> 2);
#pragma unroll 4
            for (unsigned i = threadIdx.x; i < n1; i += 256) d1[i] = s1[i];
        }
    }
}
int mg_pull(Context& c, int n, const void* const* src, void* const* dst, const size_t* bytes) {
    PullSegs S;
    S.n = 0;
    unsigned long long tiles = 0;
    for (int k = 0; k < n; k++) {
        if (!bytes[k]) continue;
        if (S.n == 32) { set_error("mg_pull: too many segments"); return PRB_ERR_ARG; }
        if (bytes[k] & 3) { set_error("mg_pull: lengths are multiples of 4 bytes"); return PRB_ERR_ARG; }
        S.src[S.n] = (const char*)src[k]; S.dst[S.n] = (char*)dst[k]; S.bytes[S.n] = bytes[k]; S.first[S.n] = tiles;
        tiles += (bytes[k] + kPullTileBytes - 1) / kPullTileBytes;
        S.n++;
    }
    if (!S.n) return PRB_OK;
    for (int k = S.n; k <= 32; k++) S.first[k] = tiles;
    for (int k = S.n; k < 32; k++) { S.src[k] = nullptr; S.dst[k] = nullptr; S.bytes[k] = 0; }
    const unsigned grid = (unsigned)std::min<unsigned long long>(tiles, (unsigned long long)c.smCount * 16);
    PRB_LAUNCH(c, k_mg_pull, grid, 256, 0, S);
    return PRB_OK;
}

int mg_allgather(Context& c, size_t arenaOffset, size_t elemBytes, const long long* lo) {
    if (!c.mg.active()) return PRB_OK;
    const int W = c.mg.world, me = c.mg.rank;
    PRB_TRY(mg_barrier(c));
    const void* src[kMaxRanks];
    void* dst[kMaxRanks];
    size_t bytes[kMaxRanks];
    int n = 0;
    for (int qi = 1; qi < W; qi++) {
        const int q = (me + qi) % W;
        const size_t a = (size_t)lo[q] * elemBytes, b = (size_t)lo[q + 1] * elemBytes;
        src[n] = c.mg.peer[q] + arenaOffset + a; dst[n] = c.mg.arena + arenaOffset + a; bytes[n] = b - a; n++;
    }
    PRB_TRY(mg_pull(c, n, src, dst, bytes));
    PRB_TRY(mg_barrier(c));
    return PRB_OK;
}

}  // namespace prb

using namespace prb;

extern "C" {

int prb_mg_init(prb_context* h, int rank, int world, int64_t arena_bytes, void* handle_out /* 64 bytes */) {
    if (!h || !handle_out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || arena_bytes < (int64_t)kMgHeaderBytes) {
        set_error("prb_mg_init: bad argument (world <= 8, arena >= 16 KiB)");
        return PRB_ERR_ARG;
    }
    Context& c = h->c;
    PRB_DEVICE(c);
    if (c.mg.arena) { set_error("prb_mg_init: already initialised"); return PRB_ERR_STATE; }
    PRB_CUDA(cudaMalloc((void**)&c.mg.arena, (size_t)arena_bytes));
    PRB_CUDA(cudaMemset(c.mg.arena, 0, kMgHeaderBytes));
    c.mg.arenaBytes = (size_t)arena_bytes;
    c.mg.rank = rank;
    c.mg.world = world;
    c.mg.peer[rank] = c.mg.arena;
    c.mg.peerOpen[rank] = false;
    c.mg.reset_allocs();
    cudaIpcMemHandle_t hd;
    PRB_CUDA(cudaIpcGetMemHandle(&hd, c.mg.arena));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &hd, 64);
    return PRB_OK;
}

int prb_mg_set_peer(prb_context* h, int peer_rank, const void* handle /* 64 bytes */) {
    if (!h || !handle) return PRB_ERR_ARG;
    Context& c = h->c;
    if (!c.mg.arena || peer_rank < 0 || peer_rank >= c.mg.world) { set_error("prb_mg_set_peer: bad rank or prb_mg_init not called"); return PRB_ERR_ARG; }
    if (peer_rank == c.mg.rank) return PRB_OK;
    PRB_DEVICE(c);
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle, 64);
    void* p = nullptr;
    PRB_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c.mg.peer[peer_rank] = (char*)p;
    c.mg.peerOpen[peer_rank] = true;
    return PRB_OK;
}

int prb_mg_barrier(prb_context* h) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    for (int r = 0; r < c.mg.world; r++)
        if (!c.mg.peer[r]) { set_error("prb_mg_barrier: peer arena not set"); return PRB_ERR_STATE; }
    PRB_TRY(mg_barrier(c));
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    int err = 0;
    PRB_CUDA(cudaMemcpy(&err, &((MgHeader*)c.mg.arena)->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { set_error("multi-GPU barrier timed out waiting for a peer"); return PRB_ERR_CUDA; }
    return PRB_OK;
}

// Host-only: which rank runs which refinement pass (pass i = `count[i]` roots at depth `depth[i]`, maxDepth `depth_max`); the deal every
// rank derives for itself in prb_extract.
int prb_mg_deal_passes(int depth_max, int n_passes, const int32_t* depth, const int32_t* count, int world, int32_t* owner_out) {
    if (n_passes < 0 || world < 1 || world > kMaxRanks || (n_passes > 0 && (!depth || !count || !owner_out))) { set_error("prb_mg_deal_passes: bad argument"); return PRB_ERR_ARG; }
    deal_passes(depth_max, n_passes, depth, count, world, owner_out);
    return PRB_OK;
}

// Host-only shard plan (no GPU needed): splits `count` units into `world` contiguous chunks of
// (nearly) equal size; out[r] = first unit of rank r, out[world] = count.  Every rank computes the
// same plan from the same replicated counts.
int prb_mg_plan(int64_t count, int world, int64_t* out) {
    if (!out || world < 1 || world > kMaxRanks || count < 0) { set_error("prb_mg_plan: bad argument"); return PRB_ERR_ARG; }
    for (int r = 0; r <= world; r++) out[r] = (count * r) / world;
    return PRB_OK;
}

}  // extern "C"
