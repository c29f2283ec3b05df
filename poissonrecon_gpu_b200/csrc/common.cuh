// Shared declarations of the sm_100a pipeline: context, device buffers, launch helpers and the
// small closed-form tables (neighbour LUTs, cube corner / edge numbering) used by every stage.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "prb.h"
#include "bspline_host.h"

namespace prb {

typedef long long i64;
typedef unsigned long long u64;

void set_error(const std::string& msg);

#define PRB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            prb::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return PRB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)
#define PRB_TRY(call)                 \
    do {                              \
        int r__ = (call);             \
        if (r__ != PRB_OK) return r__; \
    } while (0)

constexpr int kMaxDepth = 12;
constexpr int kSMs = 148;   // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

// Device buffers come from a per-context arena (arena.cpp): a host-side free-list allocator over a
// few large cudaMalloc slabs, keyed by the context's stream.  All work of a context is ordered on
// that one stream, so freed blocks are reusable at once; after the first run of a given size no
// driver allocation happens any more (deterministic step times).
void arena_register(cudaStream_t st);
void arena_unregister(cudaStream_t st);
int arena_alloc(void** out, size_t bytes, cudaStream_t st);
void arena_free(void* p, cudaStream_t st);
void arena_stats(cudaStream_t st, size_t* reserved, size_t* peak, long* driverCalls);

// Move-only owner of an arena block: an early return (PRB_TRY / PRB_CUDA) gives the block back to the arena.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), s(o.s), cap(o.cap) { o.p = nullptr; o.n = 0; o.cap = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; cap = o.cap; o.p = nullptr; o.n = 0; o.cap = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    int alloc(size_t count, cudaStream_t st) {
        release();
        s = st;
        n = count;
        if (count == 0) { p = nullptr; return PRB_OK; }
        return arena_alloc((void**)&p, count * sizeof(T), st);
    }
    void release() {
        if (p) arena_free(p, s);
        p = nullptr;
        n = 0;
        cap = 0;
    }
    // grow-only reuse (workspace buffers kept by the context between passes and runs)
    size_t cap = 0;
    int ensure(size_t count, cudaStream_t st) {
        if (p && count <= cap) { n = count; return PRB_OK; }
        if (p) arena_free(p, s);
        p = nullptr;
        s = st;
        cap = count + count / 8 + 256;
        n = count;
        return arena_alloc((void**)&p, cap * sizeof(T), st);
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Growable pinned host buffer (D2H target of the mesh: pageable memory would halve PCIe throughput).
template <class T>
struct HBuf {
    T* p = nullptr;
    size_t cap = 0;
    unsigned generation = 0;      // bumped by every reallocation (the new block may well sit at the old address)
    int reserve(size_t count) {
        if (count <= cap) return PRB_OK;
        generation++;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = count + count / 4 + 1024;
        PRB_CUDA(cudaMallocHost((void**)&p, want * sizeof(T)));
        cap = want;
        return PRB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// ---------------------------------------------------------------------------------------------
// Multi-GPU (one process per GPU, all GPUs of one NVLink/NVSwitch box).  The octree topology is
// replicated on every rank (it is built from the same samples, bit-identically); the heavy field
// stages are sharded by Morton range and exchange data through a PEER-MAPPED ARENA: one
// cudaMalloc block per rank, exported with cudaIpcGetMemHandle and opened by every other rank,
// in which all ranks make the same sequence of sub-allocations (same sizes -> same offsets), so
// `peerArena[r] + offset` addresses rank r's copy of a buffer.  Kernels read halo data with plain
// loads / cp.async on those peer pointers (NVLink), and ranks synchronise with epoch flags in
// the arena header -- no host round trip, no NCCL call on the data path.
constexpr int kMaxRanks = 8;
constexpr size_t kMgHeaderBytes = 16384;     // flags + dot-product slots
struct MgCgLine {                             // one 128-byte line per (barrier parity, sender): the sender's per-depth partial sums and,
    double v[14];                             // written last with release semantics, the barrier epoch they belong to
    unsigned epoch;
    unsigned pad[3];
};
struct MgHeader {
    unsigned flags[kMaxRanks][32];            // flags[r][0]: last epoch rank r has arrived at (one 128-B line per writer)
    double slots[2][kMaxRanks][32];           // partial sums of the stages between kernels (iso value, mesh totals) [parity][rank][slot]
    MgCgLine cg[2][kMaxRanks];                // in-kernel barriers of the CG solve (solver.cu cg_sync)
    int xchg[2][kMaxRanks][64];               // small integer all-gathers between kernels (mg_exchange_ints), double buffered
    int error;                                // set by a kernel whose peer wait timed out
};
static_assert(sizeof(MgCgLine) == 128, "one line per sender");
static_assert(sizeof(MgHeader) <= kMgHeaderBytes, "arena header too small");
struct MgDev {                                // passed by value to kernels
    int rank, world;
    MgHeader* hdr;                            // own header
    MgHeader* peerHdr[kMaxRanks];
    long long spinCycles;                     // give-up limit of a peer wait, in SM clock cycles
};
struct MgState {
    int rank = 0, world = 1;
    char* arena = nullptr;
    size_t arenaBytes = 0, used = 0;
    char* peer[kMaxRanks] = {nullptr};        // peer[rank] == arena
    bool peerOpen[kMaxRanks] = {false};
    unsigned epoch = 0;                       // last epoch used (host-tracked, identical on all ranks)
    unsigned xchgCount = 0;                   // number of mg_exchange_ints calls so far (buffer parity)
    unsigned cgEpoch = 0;                     // the same for the in-kernel barriers of the CG solve (their own flag words)
    int minShardRows = 65536;                 // depths with fewer rows stay replicated
    long long spinCycles = 4000000000ll;      // ~2 s at 2 GHz (option "mg_timeout_ms"): a rank that died must not hang the others (and the box)
    bool active() const { return world > 1; }
    void reset_allocs() { used = kMgHeaderBytes; }
    template <class T>
    T* alloc(size_t count, size_t* offset) {  // 256-byte aligned bump allocation; nullptr when the arena is full
        size_t a = (used + 255) & ~(size_t)255, b = a + count * sizeof(T);
        if (!arena || b > arenaBytes) return nullptr;
        used = b;
        if (offset) *offset = a;
        return (T*)(arena + a);
    }
    MgDev dev() const {
        MgDev d;
        d.rank = rank; d.world = world; d.hdr = (MgHeader*)arena; d.spinCycles = spinCycles;
        for (int r = 0; r < kMaxRanks; r++) d.peerHdr[r] = (MgHeader*)peer[r];
        return d;
    }
};
int mg_barrier(struct Context& c);            // device-side flag barrier on the context stream (all ranks must call it)
// all-gather of a buffer that lives at the same arena offset on every rank and whose element range [lo[r], lo[r+1]) was produced by
// rank r: barrier, pull the other ranks' ranges over NVLink (each rank starts with its successor), barrier
// all-gather of up to 64 ints per rank through the arena headers: all[r * stride + k] = value k of rank r (all ranks must call it)
int mg_exchange_ints(struct Context& c, const int* mine, int n, int* all, int stride = -1);
void deal_passes(int D, int n, const int* depth, const int* count, int world, int* owner);     // mc.cu
int mg_pull(struct Context& c, int n, const void* const* src, void* const* dst, const size_t* bytes);     // one-launch pull of segments from peer arenas
int mg_allgather(struct Context& c, size_t arenaOffset, size_t elemBytes, const long long* lo /* [world + 1] */);

// One pass (the main depth-D pass or a refinement pass) of mesh output.
struct PassRecord { int kind, nv, nt; };   // kind: 0 main, 1 coarse (single root), 2 batched per depth

// look-back scan state of a context (scan.cuh)
struct ScanWork {
    unsigned long long* desc = nullptr;      // [maxTiles]
    unsigned* ticket = nullptr;              // [1] + total [1] (as int)
    size_t maxTiles = 0;
    unsigned epoch = 0;
};
struct Context {
    int device = 0, D = 0;
    ScanWork scanWork;             // scan.cuh: look-back descriptors + ticket (views of scanDesc / scanTicket)
    DBuf<unsigned long long> scanDesc;
    DBuf<unsigned> scanTicket;
    int* hScanTotal = nullptr;     // pinned host word the scan totals are copied into
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[10];
    int stage = 0;                 // 0 none, 1 points, 2 octree, 3 splat, 4 solve, 5 extract
    int launches = 0;
    double cgTol = 1e-5;
    int cgMaxIter = 10000;
    int cgZigzag = 1;
    int cgTiming = 0;              // 1: CTA 0 of the CG kernel records the time it spends in every phase and barrier ("cg_phase_ns")
    long long cgPhaseNs[8] = {0};
    int cgBulk = 1;                // CG streaming phases through TMA bulk copies (solver.cu stream_pairs_bulk); 0: per-thread cp.async ring
    int refineBoundCheck = 0;      // 1: evaluate every refinement brick and verify the certified signs (tests)
    long long boundChecked = 0, boundEvaluated = 0;
    int doRefine = 1;
    int refineImplicit = 1;        // 0: materialised virtual subtrees for every pass (debug / cross-check)
    int smCount = kSMs;
    size_t deviceMemBytes = 0;
    // ---- samples
    i64 N = 0;
    DBuf<float> rawP, rawN;        // file coordinates, [N][3]
    float *rawPp = nullptr, *rawNp = nullptr;   // where they live: rawP.p / rawN.p, or the peer-mapped arena (prb_set_points_sharded)
    bool rawSharded = false;       // multi-GPU: only this rank's slice has been uploaded; stage_octree gathers the others over NVLink
    size_t mgRawPOff = 0, mgRawNOff = 0, mgVOff = 0;
    float* Vp = nullptr;           // vector field of the current run (V.p, or inside the arena when splat is sharded)
    DBuf<float> P, Nr;             // Morton-sorted, normalised samples / rescaled normals [N][3]
    DBuf<u64> sortedKey;           // Morton key per sorted sample
    DBuf<int> sortedIdx;           // sorted position -> input index
    DBuf<int> p2n;                 // sorted sample -> depth-D slot (local)
    float center[3] = {0, 0, 0}, scale = 1;
    // ---- octree: one global node index space, depth slabs at base[d] (SoA)
    int M = 0;
    int base[kMaxDepth + 2] = {0}, cnt[kMaxDepth + 1] = {0};
    DBuf<int> dBase;               // device copy of base[0..D+1]
    DBuf<u64> key;
    DBuf<int> parent, child0, pidx, pnum, didx, dnum;
    DBuf<int> neighs;              // [M][27]
    DBuf<ushort4> offs;            // per node (ox, oy, oz, depth)
    DBuf<int> sgTab;               // [nSg][64] super-group table: first row of every block of the 4x4x4 cube (octree.cu k_sg_table)
    DBuf<int> sgTab4;              // [nSg][8][12] the same in the staging order of the CG kernel (solver.cu k_sg_table4)
    int nSg = 0;
    // ---- tables
    BSplineTables tab;
    DBuf<float> dMaxDepthFn, dBaseFn, dDfT, dStencil;
    DBuf<int> dDfOffset;
    DBuf<double> dFfX, dD2X;       // cross-depth 1-D integrals (cascadic mode)
    DBuf<int> dCrossOff;
    DBuf<float> bCas;              // right-hand side of the cascadic mode (divg stays the reference's divergence)
    int cascadic = 0;              // OPT-IN, outside reference parity: coarse-to-fine coupling of the depths (solver.cu k_cascadic_rhs)
    DBuf<float> dBvAnc, dBvOwn, dBvCell, dBvGrid;    // base-function values at cell corners per (depth, ancestor level) (mc.cu k_build_bv)
    int bvAncOff[kMaxDepth + 1] = {0}, bvOwnOff[kMaxDepth + 1] = {0};
    // ---- fields
    DBuf<float> V;                 // [M_D][3]
    DBuf<float> divg, x;          // x is padded: node i lives at xv[i] = x.p[7 + i] (sibling blocks 32-byte aligned)
    float* xv = nullptr;
    float* divgv = nullptr;        // divg.p + 7, same padding
    int divMode = 1;               // 1: block-table / profile divergence (field.cu); 0: first-version kernels through the 27-neighbour rows (cross-check)
    DBuf<float> pointValue;
    float iso = 0;
    float isoPlain = 0, isoWeighted = 0;   // the reference's mean of chi over the samples / the density-weighted mean (opt-in, "iso_density_weighted")
    int isoDensityWeighted = 0;
    int cgIters[kMaxDepth + 1] = {0};
    i64 cgRowIters = 0;
    // ---- mesh
    DBuf<float> meshV;             // device
    DBuf<int> meshT;
    i64 nMeshV = 0, nMeshT = 0;
    HBuf<float> hMeshV;            // pinned host copies (prb_get_mesh)
    HBuf<int> hMeshT;
    bool hMeshValid = false;
    // OPTIONAL early device -> host copy of the main marching-cubes piece (it leads the mesh and is final before the refinement passes
    // start): runs on its own stream under the passes; prb_get_mesh then only copies what the passes added.  Off by default: measured
    // on scan5m_d10 it saves the 1.2 ms download but the passes' ~50 short host round trips queue behind the bulk PCIe traffic and
    // lose as much (36.8 vs 36.7 ms end to end, profiles/r02/e2e_early_copy_ab.txt)
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evMainPiece = nullptr, evEarlyCopy = nullptr, evPositions = nullptr, evNormals = nullptr;
    bool normalsPending = false;  // the upload of the normals runs on copyStream behind the positions; the sort's gather pass waits for evNormals
    i64 earlyV = 0, earlyT = 0;   // vertices / triangles already in hMeshV / hMeshT
    int earlyMeshCopy = 0;        // option "early_mesh_copy"
    std::vector<PassRecord> passes;            // every pass of the reconstruction (all ranks), in output order
    struct PieceRecord { long long pass, vBase, nv, tBase, nt; };
    std::vector<PieceRecord> layout;           // the pieces of the mesh THIS context holds: global vertex / triangle offsets (1 GPU: all passes)
    i64 nGlobalV = 0, nGlobalT = 0;            // size of the whole mesh (== nMeshV / nMeshT on one GPU)
    std::vector<int> subdivide;    // host copy of the refined leaves (node ids)
    DBuf<float> vval;              // [M][8] corner values (valid at the owner's slot)
    // grow-only workspace of the refinement passes (kept across runs)
    DBuf<float> wsVal7, wsLow;
    DBuf<unsigned char> wsCat, wsNtri;
    DBuf<unsigned short> wsEmask, wsVpre;
    DBuf<int> wsVbase, wsTbase;
    prb_stats stats;
    // ---- optional sub-stage timeline (prb_set_option "detail", 1): events recorded at named points of the run; the intervals between
    // consecutive marks are reported by prb_get_array("detail_ms") / ("detail_names", NUL-separated)
    int detail = 0;
    std::vector<cudaEvent_t> detailEv;
    std::vector<std::string> detailName;
    size_t detailUsed = 0;
    // ---- multi-GPU
    MgState mg;
    int shardFrom = 0;                         // first sharded depth (D+1: none); set by stage_octree
    int sgLo[kMaxDepth + 2][kMaxRanks + 1];    // super-group range of every rank at every depth
    int rowLo[kMaxDepth + 2][kMaxRanks + 1];   // node range of every rank at every depth (sharded depths), else [base, base+cnt) for rank 0..
    float* mgP = nullptr;                      // CG direction vector inside the arena (padded like x)
    float* mgX = nullptr;                      // solution inside the arena
    size_t mgPOff = 0, mgXOff = 0;
    float* mgVval = nullptr;                   // corner values [M][8] inside the arena
    size_t mgVvalOff = 0;
    float* vvalPtr = nullptr;                  // where the corner values of the current run live (vval.p or mgVval)
};

inline void mark(Context& c, const char* name) {
    if (!c.detail) return;
    if (c.detailUsed == c.detailEv.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; c.detailEv.push_back(e); c.detailName.emplace_back(); }
    c.detailName[c.detailUsed] = name;
    cudaEventRecord(c.detailEv[c.detailUsed++], c.stream);
}

// stages (implemented in the .cu files)
int stage_octree(Context& c);
int stage_splat(Context& c);
int stage_divergence(Context& c);
int stage_solve(Context& c);
int stage_iso(Context& c);
int stage_extract(Context& c);
int upload_tables(Context& c);
int build_cg_table(Context& c);       // solver.cu: sgTab -> sgTab4

// exclusive scan of n ints on the context stream; returns the grand total through *total_host
// (synchronises the stream) when total_host != nullptr.
int exclusive_scan(Context& c, const int* in, int* out, i64 n, i64* total_host, int hostSlot = 0);
int ensure_copy_stream(Context& c);           // second stream + its events (uploads / downloads that run under compute)

#define PRB_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);         \
        (ctx).launches++;                                                       \
    } while (0)

inline int div_up(i64 a, i64 b) { return (int)((a + b - 1) / b); }
// grid for a grid-stride kernel: enough CTAs to fill the machine, a multiple of the SM count
inline int grid_for(const Context& c, i64 n, int block, int perSM = 8) {
    i64 need = (n + block - 1) / block;
    i64 cap = (i64)c.smCount * perSM;
    if (need <= cap) return need > 0 ? (int)need : 1;
    return (int)cap;
}

// ---------------------------------------------------------------------------------------------
// Closed forms shared by host and device.
// Neighbour slot j = 9(dx+1)+3(dy+1)+(dz+1); child code c = x<<2|y<<1|z (reference LUTparent /
// LUTchild, main.cu:80-99: per axis t = bit + dir, parent dir = floor(t/2), child bit = t&1).
__host__ __device__ inline void lut_parent_child(int c, int j, int& pj, int& cc) {
    int dx = j / 9 - 1, dy = (j / 3) % 3 - 1, dz = j % 3 - 1;
    int tx = ((c >> 2) & 1) + dx, ty = ((c >> 1) & 1) + dy, tz = (c & 1) + dz;
    int px = tx < 0 ? 0 : (tx > 1 ? 2 : 1), py = ty < 0 ? 0 : (ty > 1 ? 2 : 1), pz = tz < 0 ? 0 : (tz > 1 ? 2 : 1);
    pj = px * 9 + py * 3 + pz;
    cc = ((tx & 1) << 2) | ((ty & 1) << 1) | (tz & 1);
}
// cube corner numbering of the MC tables ("ring" order, main.cu:2444-2455)
__host__ __device__ inline int ring_index(int x, int y, int z) {
    int r = x | (z << 2);
    if (y) r += (r & 1) ? 1 : 3;
    return r;
}
// inverse: ring index -> bit code x|y<<1|z<<2
__host__ __device__ inline int ring_to_bits(int r) {
    const int t[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    return t[r];
}
// edge e = orientation<<2 | off0 | off1<<1 (off0/off1: the two other axes, ascending; main.cu:1823-1840);
// offset of the edge along axis a, or -1 along its own axis
__host__ __device__ inline int edge_off(int e, int a) {
    int o = e >> 2;
    if (a == o) return -1;
    int dim = (a == 0) ? 0 : (a == 1 ? (o != 0) : 1);   // number of axes below a other than o
    return (e >> dim) & 1;
}
// the two axes other than o, ascending
__host__ __device__ inline void other_axes(int o, int& a0, int& a1) {
    a0 = (o == 0) ? 1 : 0;
    a1 = (o == 2) ? 1 : 2;
}

// Every entry point runs on the context's device and gives the caller's current device back on return.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        ok = cudaSetDevice(device) == cudaSuccess;
        if (!ok) { cudaGetLastError(); set_error("cudaSetDevice failed"); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define PRB_DEVICE(c) DeviceGuard guard__((c).device); if (!guard__.ok) return PRB_ERR_CUDA

}  // namespace prb

struct prb_context {
    prb::Context c;
};
