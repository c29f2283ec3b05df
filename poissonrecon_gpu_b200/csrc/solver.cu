// Per-depth conjugate gradient, all depths in ONE persistent cooperative kernel.
//
// Replaces GenerateSingleNodeLaplacian + scan + 2x copy_if CSR assembly (main.cu:1143-1329) and
// the cooperative CG sample kernel (CG_CUDA.cuh:186-324, 344-509).  The reference solves the
// D+1 independent same-depth systems one after another, each with its own CSR matrix in managed
// memory, 7 grid syncs per iteration and host loops over managed arrays around the launch.
// Here:
//   * the matrix is never formed: rows are the translation-invariant 27-point stencil
//     (4 distinct values per depth) applied through a per-super-group table: the up to 64 rows
//     under one node Q (8 sibling groups) and all their neighbours live in the 4x4x4 cube of
//     8-row blocks around Q's children;
//   * the depths are independent, so they all iterate in lock-step inside one
//     launch: one iteration of the kernel = one CG iteration of every still-active depth, with
//     per-depth alpha / beta / residual and per-depth stopping.  Grid syncs per solve drop from
//     7 * sum_d iters_d to 3 * max_d iters_d;
//   * SpMV: the super-groups of all active depths form one flat tile list; a warp takes 4
//     consecutive super-groups (one per quarter-warp) per step.  The p values a super-group
//     touches are a 6x8x8 box = 96 half-blocks of 16 bytes; they are copied global -> shared with
//     cp.async into a double buffer one step ahead of the compute (their addresses come from the
//     table, which is loaded two steps ahead), so the gather latency is off the critical path.
//     Every lane then produces EIGHT rows (an x pair times the four z of its column) from a
//     4x3x8 window read with 64-bit LDS: 6 shared-memory floats per row instead of 18 with one
//     z pair per lane -- the SpMV is bound by shared-memory wavefronts, not by issue slots or
//     HBM.  The half-block layout P = xx + 10 bz + 36 by (16-byte units, odd cube stride) makes
//     both the LDS.64 window reads (per half-warp) and the cp.async stores (per quarter-warp)
//     bank-conflict free;
//   * the row sum runs over the neighbour slots in order j = 0..26 with FMAs, exactly the
//     reference's CSR order (absent neighbours contribute an exact +0), so A*p is bit-identical;
//     dots are float products accumulated in double (CG_CUDA.cuh:217-220); alpha, beta are float;
//   * p is double buffered, so x += alpha_{k-1} p_{k-1} is deferred by one iteration and rides along
//     with the SpMV of iteration k (one float4 batch per step, loaded before and finished after the
//     step's FMAs); the two streaming phases that remain (p_k = r + beta p_{k-1};  r -= alpha Ap, r.r)
//     read their inputs through a per-thread cp.async ring in the idle SpMV buffers; every phase sweeps
//     memory in the direction opposite to the previous one (L2 reuse across the grid syncs).
// Algorithmic bytes per row per iteration: 57.5 B
//   (p=r+beta*p: 12, SpMV: 13.5 + 4 + 4, x/r update: 24).
#include "common.cuh"
#include "mg_device.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace prb {

constexpr int kCgWarps = 12;
constexpr int kCgBlock = kCgWarps * 32;
// shared box of one super-group: half-block (xx, by, bz) -- the four p values (y&1, z&1) at
// x = xx of block (xx>>1, by, bz) -- sits at 16-byte unit P = xx + 10 bz + 36 by (xx = 1..6).
// P mod 8 = xx + 2 bz + 4 by: the 16 lanes of a half-warp (two cubes, shifted by one unit) read 16
// different 8-byte bank pairs, and the staging rounds below write 8 different 16-byte bank groups.
constexpr int kUnitBz = 10, kUnitBy = 36, kCubeUnits = 145, kCubeFloats = 4 * kCubeUnits;
constexpr int kWarpBufFloats = 4 * kCubeFloats;          // 4 cubes per warp-step
constexpr int kRounds = 12;                               // 96 half-blocks / 8 lanes
constexpr int kPrefetchSteps = 2;                         // L2 prefetch distance of the SpMV table, in steps beyond the register pipeline

// Staging plan of a quarter-warp: in round r (12 rounds x 8 lanes = the 96 half-blocks) lane li
// copies one 16-byte half-block.  The cost of the gather is the number of distinct 128-byte
// lines an LDGSTS instruction touches (LSU tag stage), so every round takes its 8 half-blocks from
// as few of Q's 27 neighbours as possible (the children of one neighbour are contiguous in memory):
//   rounds 0..7  full blocks (bx = 1, 2): the lane PAIR (li, li^1) copies the two x halves of one
//                block (one full 32-byte sector): r0 / r1 the pair's own rows' blocks bz = 1 / 2 (centre
//                neighbour), r2 / r3 the y faces, r4 / r5 the z faces, r6 / r7 the x-parallel edges;
//   rounds 8..11 edge blocks (bx = 0: upper half, bx = 3: lower half), one lane each: x faces,
//                z-parallel edges, y-parallel edges, corners.
// The table holds the 64 block bases in that order: [pair][8] then [lane][4].
__host__ __device__ constexpr int plan_block(int li, int r) {      // bx * 16 + by * 4 + bz
    const int m = li >> 1, a = li >> 2, b = (li >> 1) & 1, c = li & 1;
    int bx = 0, by = 0, bz = 0;
    switch (r) {
        case 0: bx = 1 + a; by = 1 + b; bz = 1; break;
        case 1: bx = 1 + a; by = 1 + b; bz = 2; break;
        case 2: bx = 1 + (m >> 1); by = 0; bz = 1 + (m & 1); break;
        case 3: bx = 1 + (m >> 1); by = 3; bz = 1 + (m & 1); break;
        case 4: bx = 1 + (m >> 1); by = 1 + (m & 1); bz = 0; break;
        case 5: bx = 1 + (m >> 1); by = 1 + (m & 1); bz = 3; break;
        case 6: bx = 1 + (m & 1); by = 0; bz = 3 * (m >> 1); break;
        case 7: bx = 1 + (m & 1); by = 3; bz = 3 * (m >> 1); break;
        case 8: bx = 3 * a; by = 1 + b; bz = 1 + c; break;
        case 9: bx = 3 * a; by = 3 * b; bz = 1 + c; break;
        case 10: bx = 3 * a; by = 1 + c; bz = 3 * b; break;
        default: bx = 3 * a; by = 3 * b; bz = 3 * c; break;
    }
    return bx * 16 + by * 4 + bz;
}
__host__ __device__ constexpr int plan_half(int li, int r) { return r < 8 ? (li & 1) : ((li >> 2) ? 0 : 1); }
constexpr bool plan_ok() {
    bool seen[8][4][4] = {};
    for (int r = 0; r < kRounds; r++)
        for (int li = 0; li < 8; li++) {
            int u = plan_block(li, r), xx = 2 * (u >> 4) + plan_half(li, r), by = (u >> 2) & 3, bz = u & 3;
            if (xx < 1 || xx > 6 || seen[xx][by][bz]) return false;
            seen[xx][by][bz] = true;
        }
    return true;
}
static_assert(plan_ok(), "staging plan must cover the 96 half-blocks exactly once");

constexpr int kAbsent = -7;    // table entry of a missing block (the address p - 7 stays valid and 32-byte aligned; it is never read)

// sgTab4[sg][64]: block bases in staging order, kAbsent when the block does not exist.
__global__ void __launch_bounds__(256) k_sg_table4(const int* __restrict__ sgTab, i64 nSg, int* __restrict__ tab4) {
    const i64 total = nSg * 64;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const int k = (int)(t & 63);
        const int u = k < 32 ? plan_block(2 * (k >> 3), k & 7) : plan_block((k - 32) >> 2, 8 + ((k - 32) & 3));
        const int b = sgTab[(t & ~(i64)63) + u];
        tab4[t] = b >= 0 ? b : kAbsent;
    }
}

struct CgParams {
    int D;
    int sgStart[kMaxDepth + 2];   // first super-group of depth d (d = 1..D), sgStart[D+1] = total
    const int* tab4;
    int nSg;
    int zigzag;                    // sweep memory in alternating directions phase by phase (L2 reuse across phase boundaries)
    int bulk;                      // streaming phases through TMA bulk copies + mbarriers (1) or per-thread cp.async (0)
    const float* stencil;          // [D+1][27]
    const float* b;                // divergence
    // vectors indexed by node id; node 1 sits on a 32-byte boundary (pointer = allocation + 7)
    float* x;
    float* r;
    float* p;                      // direction vector, double buffered: iteration k lives in p + (k & 1) * pStride
    i64 pStride;
    float* Ap;
    double* dots;                  // [2 buffers][2 kinds][16] + [16] for the initial r.r
    int* itersOut;                 // [D+1]
    float* resOut;                 // [D+1] final r.r
    float tol2;
    int maxIter;
    // ---- multi-GPU (world == 1: everything below is unused).  Depths >= shardFrom are split by
    // super-group range; a rank updates only its rows and reads the p blocks of other ranks
    // straight from their (peer-mapped) arenas; dot products and phase boundaries go through
    // the arena header (mg_device.cuh).  Shallower depths are solved redundantly on every rank
    // with rank 0's dot products so that all ranks take identical steps.
    int world, rank, shardFrom;
    int sg0[kMaxDepth + 2], sg1[kMaxDepth + 2];       // this rank's super-group range per depth
    int row0[kMaxDepth + 2], row1[kMaxDepth + 2];     // this rank's node range per depth
    int rowLo[kMaxDepth + 2][kMaxRanks + 1];          // node ranges of all ranks (sharded depths)
    const float* peerP[kMaxRanks];
    MgDev mg;
    unsigned epoch0;
    unsigned* barCount;            // arrival counter / release flag of cg_sync (zeroed before the launch)
    unsigned* barRelease;
    int onlyDepth;                 // -1: all depths in lock-step (the reference's independent systems); d: solve depth d alone (cascadic mode)
    long long* phaseNs;            // optional [8]: time CTA 0 spent in phase C / its barrier / SpMV / barrier / phase B / barrier (ns, %globaltimer)
};

// Grid-wide (world == 1) or box-wide barrier between the phases of an iteration.  Every CTA arrives on a counter; the LAST one
// to arrive does the cross-GPU part -- warp 0, lane r <-> rank r: the lane writes this rank's per-depth partial sums (kind >= 0:
// `local[1..D]`, complete because every CTA has arrived) and then, with release semantics, the barrier epoch into ITS line of
// rank r's header, and polls the line rank r writes into this rank's header; all peers in parallel, one NVLink round trip -- and
// then releases the other CTAs through a flag.  One counter round trip instead of two cooperative grid syncs around a
// single-thread exchange.  Lines alternate with the epoch parity: a rank can run at most one barrier ahead of the slowest one.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
template <bool MG>
__device__ __forceinline__ void cg_sync(const CgParams& P, unsigned& epoch, unsigned& gen, int kind, const double* local) {
    __syncthreads();
    gen++;
    if (MG) epoch++;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int last = 0;
        if (lane == 0) {
            __threadfence();
            last = atomicAdd(P.barCount, 1u) == gen * gridDim.x - 1u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (MG) {
                // lane 0 observed every CTA's arrival (each one fenced its writes first); the shuffle above orders that before the
                // other lanes' stores, and release / acquire are cumulative: the epoch store below publishes this rank's whole
                // phase to the peer, and the release of the local CTAs after the polls publishes the peers' phases to them
                if (lane < P.world && !P.mg.hdr->error) {
                    MgCgLine* out = &P.mg.peerHdr[lane]->cg[epoch & 1u][P.rank];
                    if (kind >= 0) {
                        double2* o2 = reinterpret_cast<double2*>(out->v);              // v[0] is unused: pairs (0,1), (2,3), ...
                        for (int d = 0; d <= P.D; d += 2) o2[d >> 1] = make_double2(__ldcg(local + d), __ldcg(local + d + 1));
                    }
                    mg_store_release_sys(&out->epoch, epoch);
                    const unsigned* in = &P.mg.hdr->cg[epoch & 1u][lane].epoch;
                    const long long t0 = clock64();
                    while ((int)(mg_load_acquire_sys(in) - epoch) < 0)
                        if (clock64() - t0 > P.mg.spinCycles) { P.mg.hdr->error = 1; break; }
                }
                __syncwarp();
            }
            if (lane == 0) st_release_gpu(P.barRelease, gen);
        } else if (lane == 0) {
            while ((int)(ld_acquire_gpu(P.barRelease) - gen) < 0) {}
        }
        __syncwarp();
    }
    __syncthreads();
}
// total of a per-depth dot product after cg_sync (`epoch`: the barrier that carried it)
template <bool MG>
__device__ __forceinline__ double cg_total(const CgParams& P, unsigned epoch, int d, const double* local) {
    if (!MG) return __ldcg(local + d);
    const volatile MgCgLine* L = &P.mg.hdr->cg[epoch & 1u][0];
    if (d < P.shardFrom) return L[0].v[d];
    double t = 0.0;
    for (int r = 0; r < P.world; r++) t += L[r].v[d];
    return t;
}

__device__ __forceinline__ void cp_async16(unsigned smemAddr, const void* gptr, int srcBytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smemAddr), "l"(gptr), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// warp-level: sum `part` over the warp and add it to the block accumulator of depth d
__device__ __forceinline__ void warp_add(double part, double* sAcc, int d, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    if (lane == 0 && part != 0.0) atomicAdd(&sAcc[d], part);
}

// Streaming phases: two input vectors are read through a per-thread ring of kStreamQ 16-byte
// cp.async copies into the warp's (otherwise idle) SpMV buffers, so that kStreamQ x 2 x 512 bytes
// per warp are in flight at any time (twice what register staging allowed with 12 warps of 168
// registers).  Work unit = one 32-lane batch of float4; the batches of all active depths form
// one flat list (pre[d] = first batch of depth d), swept forwards or backwards (zig-zag).
// A thread reads back exactly the 16 bytes it copied, so no warp-level synchronisation is needed.
constexpr int kStreamQ = 12;
struct BatchCursor {
    int d, lo, hi, r0, n4;          // current depth, its batch range, its first row and float4 count
    __device__ __forceinline__ void init(const int* pre, const int* row0, const int* row1) {
        d = 1; lo = pre[1]; hi = pre[2]; r0 = row0[1]; n4 = (row1[1] - row0[1]) >> 2;
    }
    // -> float index of this lane's float4 in batch u of the flat list (-1: past the end of the depth's rows)
    __device__ __forceinline__ int locate(const int* pre, const int* row0, const int* row1, int u, int lane) {
        if (u >= hi || u < lo) {
            while (u >= hi) { d++; lo = hi; hi = pre[d + 1]; }
            while (u < lo) { d--; hi = lo; lo = pre[d]; }
            r0 = row0[d]; n4 = (row1[d] - r0) >> 2;
        }
        const int c4 = (u - lo) * 32 + lane;
        return c4 < n4 ? r0 + 4 * c4 : -1;
    }
};
template <class F>
__device__ __forceinline__ void stream_pairs(const int* pre, int total, const int* row0, const int* row1, const float* __restrict__ A0, const float* __restrict__ A1,
                                             bool rev, float* ring /* this warp's smem */, int gwarp, int nwarps, int lane, F&& consume) {
    const unsigned ringS = (unsigned)__cvta_generic_to_shared(ring);
    BatchCursor ci, cc;
    ci.init(pre, row0, row1); cc.init(pre, row0, row1);
    int bi = gwarp;                                    // next batch to issue
    auto issue = [&](int slot) {
        if (bi < total) {
            const int i = ci.locate(pre, row0, row1, rev ? total - 1 - bi : bi, lane);
            if (i >= 0) {
                const unsigned dst = ringS + 16u * (unsigned)(slot * 32 + lane);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(A0 + i) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * 32u * kStreamQ), "l"(A1 + i) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        bi += nwarps;
    };
#pragma unroll
    for (int q = 0; q < kStreamQ; q++) issue(q);
    int slot = 0;
    for (int b = gwarp; b < total; b += nwarps) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kStreamQ - 1) : "memory");
        const int i = cc.locate(pre, row0, row1, rev ? total - 1 - b : b, lane);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (i >= 0) {
            v0 = *reinterpret_cast<const float4*>(ring + 4 * (slot * 32 + lane));
            v1 = *reinterpret_cast<const float4*>(ring + 4 * ((kStreamQ + slot) * 32 + lane));
        }
        consume(cc.d, i, v0, v1);           // called by ALL lanes (i < 0: no element), the depth is warp-uniform
        issue(slot);
        slot = slot + 1 == kStreamQ ? 0 : slot + 1;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// The same streaming phases through the TMA engine (default): the two vectors of a batch are contiguous 512-byte runs of a depth
// slab, so ONE lane issues two bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) per batch into the warp's ring slot
// and arms the slot's mbarrier with the byte count (expect_tx); the warp waits on the barrier's phase parity before reading the
// slot.  32 LDGSTS instructions with 32 address computations each become two instructions issued by one thread.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
template <class F>
__device__ __forceinline__ void stream_pairs_bulk(const int* pre, int total, const int* row0, const int* row1, const float* __restrict__ A0, const float* __restrict__ A1,
                                                  bool rev, float* ring /* this warp's smem */, unsigned barS /* this warp's kStreamQ mbarriers */, unsigned& phaseBits,
                                                  int gwarp, int nwarps, int lane, F&& consume) {
    const unsigned ringS = (unsigned)__cvta_generic_to_shared(ring);
    // what the generic proxy (LDGSTS / LDS of the previous phase) did to these buffers is ordered before the async-proxy writes below
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    BatchCursor ci, cc;
    ci.init(pre, row0, row1); cc.init(pre, row0, row1);
    int bi = gwarp;                                    // next batch to issue
    auto issue = [&](int slot) {
        if (bi < total) {
            const int u = rev ? total - 1 - bi : bi;
            const int i0 = ci.locate(pre, row0, row1, u, 0);                       // first float of the batch (a batch is never empty)
            const int cnt4 = min(32, ci.n4 - (u - ci.lo) * 32);
            if (lane == 0) {
                const unsigned bytes = 16u * (unsigned)cnt4, bar = barS + 8u * (unsigned)slot;
                mbar_expect_tx(bar, 2u * bytes);
                bulk_g2s(ringS + 512u * (unsigned)slot, A0 + i0, bytes, bar);
                bulk_g2s(ringS + 512u * (unsigned)(kStreamQ + slot), A1 + i0, bytes, bar);
            }
        }
        bi += nwarps;
    };
#pragma unroll 1
    for (int q = 0; q < kStreamQ; q++) issue(q);
    int slot = 0;
    for (int b = gwarp; b < total; b += nwarps) {
        mbar_wait(barS + 8u * (unsigned)slot, (phaseBits >> slot) & 1u);
        phaseBits ^= 1u << slot;
        const int i = cc.locate(pre, row0, row1, rev ? total - 1 - b : b, lane);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (i >= 0) {
            v0 = *reinterpret_cast<const float4*>(ring + 4 * (slot * 32 + lane));
            v1 = *reinterpret_cast<const float4*>(ring + 4 * ((kStreamQ + slot) * 32 + lane));
        }
        consume(cc.d, i, v0, v1);           // called by ALL lanes (i < 0: no element), the depth is warp-uniform
        __syncwarp();                       // every lane has read the slot before it is refilled
        issue(slot);
        slot = slot + 1 == kStreamQ ? 0 : slot + 1;
    }
}

extern __shared__ __align__(16) float sDyn[];   // [kCgWarps][2][kWarpBufFloats]

template <bool MG>
__global__ void __launch_bounds__(kCgBlock, 1) k_cg_all_depths(const __grid_constant__ CgParams P) {
    __shared__ __align__(16) float sSt[kMaxDepth + 1][4];
    __shared__ double sAcc[kMaxDepth + 1];
    __shared__ float sR1[kMaxDepth + 1], sAlpha[kMaxDepth + 1], sBeta[kMaxDepth + 1];
    __shared__ int sActive[kMaxDepth + 1], sIter[kMaxDepth + 1];
    __shared__ int sStep[kMaxDepth + 3];      // sStep[d] = first flat SpMV step of depth d (active depths only), sStep[D+1] = total
    __shared__ int sPend[kMaxDepth + 1];      // x of depth d still lacks alpha_{k-1} p_{k-1} (applied under the next SpMV)
    __shared__ int sXup[kMaxDepth + 3];       // sXup[d] = first flat 32-lane batch of the pending x updates of depth d
    __shared__ int sAct4[kMaxDepth + 3];      // the same over the ACTIVE depths (streaming phases)
    __shared__ __align__(8) unsigned long long sBar[kCgWarps][kStreamQ];   // mbarriers of the bulk-copy ring, one per warp and slot
    const int D = P.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = blockIdx.x * kCgWarps + warp, nwarps = gridDim.x * kCgWarps;
    const int gthread = blockIdx.x * kCgBlock + tid, nthreads = gridDim.x * kCgBlock;
    if (tid < (D + 1) * 4) {
        // one representative per number of off-centre axes: j = 13, 14, 17, 26
        const int rep[4] = {13, 14, 17, 26};
        sSt[tid >> 2][tid & 3] = P.stencil[(tid >> 2) * 27 + rep[tid & 3]];
    }
    if (tid <= D) sAcc[tid] = 0.0;
    const unsigned barS = (unsigned)__cvta_generic_to_shared(&sBar[warp][0]);
    unsigned phaseBits = 0u;                  // parity of the next completion of every ring slot (warp-uniform)
    if (P.bulk && lane == 0) {
        for (int q = 0; q < kStreamQ; q++) mbar_init(barS + 8u * (unsigned)q, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---- depth 0: a 1x1 system, solved by one thread with the same recurrences
    if (blockIdx.x == 0 && tid == 0 && P.onlyDepth <= 0) {
        float a00 = sSt[0][0];
        float x0 = 0.f, r = P.b[0], p = 0.f, r0 = 0.f;
        float r1 = (float)(double)(r * r);
        int k = 1;
        while (r1 > P.tol2 && k <= P.maxIter) {
            if (k > 1) { float be = r1 / r0; p = __fadd_rn(r, __fmul_rn(be, p)); } else p = r;
            float Ap = __fmul_rn(a00, p);
            double dd = (double)(p * Ap);
            float al = (float)((double)r1 / dd);
            x0 = __fmaf_rn(al, p, x0);
            r = __fmaf_rn(-al, Ap, r);
            r0 = r1;
            r1 = (float)(double)(r * r);
            k++;
        }
        P.x[0] = x0;
        P.itersOut[0] = k - 1;
        P.resOut[0] = r1;
    }
    // ---- init: x = 0, r = b, p = 0, r1 = r.r per depth
    {
        // b (the divergence) is an unpadded array, so its quads are not 16-byte aligned: scalar loads
        for (int d = 1; d <= D; d++) {
            if (P.onlyDepth >= 0 && d != P.onlyDepth) continue;        // (the other depths keep their solution)
            const int i0 = P.row0[d], i1 = P.row1[d];
            double part = 0.0;
            for (int i = i0 + gthread; i < i1; i += nthreads) {
                float bv = P.b[i];
                P.x[i] = 0.f; P.r[i] = bv; P.p[i] = 0.f;                // p_0 = 0 (buffer 0); iteration 1 writes buffer 1
                part += (double)(bv * bv);
            }
            warp_add(part, sAcc, d, lane);
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&P.dots[64 + tid], sAcc[tid]);   // dedicated init buffer
    }
    unsigned epoch = P.epoch0, gen = 0;
    cg_sync<MG>(P, epoch, gen, 0, P.dots + 64);
    if (tid >= 1 && tid <= D) {
        float r1 = (float)cg_total<MG>(P, epoch, tid, P.dots + 64);
        sR1[tid] = r1; sIter[tid] = 1; sBeta[tid] = 0.f; sAlpha[tid] = 0.f; sPend[tid] = 0;
        sActive[tid] = (r1 > P.tol2 && 1 <= P.maxIter && (P.onlyDepth < 0 || tid == P.onlyDepth)) ? 1 : 0;
    }
    __syncthreads();

    // ---- per-lane constants of the SpMV
    // quarter-warp q works on cube q; within it lane li = (Xp, Y) produces the rows
    // x = 2 + 2 Xp + {0, 1}, y = 2 + Y, z = 2..5 of the 8x8x8 node cube
    const int q = lane >> 3, li = lane & 7, Xp = li >> 2, Y = li & 3;
    float* const wbuf = sDyn + (size_t)warp * 2 * kWarpBufFloats;
    const unsigned cubeS = (unsigned)__cvta_generic_to_shared(wbuf) + 4u * (unsigned)(q * kCubeFloats);   // byte address of this lane's cube, buffer 0
    const float* const cube0 = wbuf + q * kCubeFloats;
    const int Ylo = Y & 1;
    int dstOff[kRounds];       // float offset of the lane's half-block of round r inside the cube
#pragma unroll
    for (int r = 0; r < kRounds; r++) {
        const int u = plan_block(li, r);
        dstOff[r] = 4 * ((2 * (u >> 4) + plan_half(li, r)) + kUnitBz * (u & 3) + kUnitBy * ((u >> 2) & 3));
    }
    const int hFull = 4 * (li & 1), hEdge = (li >> 2) ? 0 : 4;    // source offset of the x half copied in rounds 0..7 / 8..11
    // window column (dxp, dy): xx = 1 + 2 Xp + dxp, yy = 1 + Y + dy, z pairs bz = 0..3 at stride 4 kUnitBz floats
    int colOff[3];
#pragma unroll
    for (int dy = 0; dy < 3; dy++) {
        int yy = 1 + Y + dy;
        colOff[dy] = 4 * ((1 + 2 * Xp) + kUnitBy * (yy >> 1)) + 2 * (yy & 1);
    }
    const bool upper = lane >= 16;
    int colA[3], colB[3];      // z = 1 (pair bz = 0, odd word) and z = 6 (pair bz = 3, even word), swapped in the upper half-warp
#pragma unroll
    for (int dy = 0; dy < 3; dy++) {
        colA[dy] = colOff[dy] + (upper ? 12 * kUnitBz : 1);
        colB[dy] = colOff[dy] + (upper ? 1 : 12 * kUnitBz);
    }
    const int outOff = 2 * Ylo;          // the lane's rows (y = Y & 1) start 2 (Y & 1) floats after the block base

    long long tPrev = 0;
    const bool timing = P.phaseNs != nullptr && blockIdx.x == 0 && tid == 0;
    auto tick = [&](int slot) {
        if (!timing) return;
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
        if (slot >= 0) P.phaseNs[slot] += t - tPrev;
        tPrev = t;
    };
    tick(-1);
    int phase = 0;
    int it = 1;
    for (;; it++) {
        int anyActive = 0;
        for (int d = 1; d <= D; d++) anyActive |= sActive[d];
        if (!anyActive) break;
        if (tid <= D) sAcc[tid] = 0.0;
        if (tid == 0) {
            int acc = 0;
            sStep[0] = 0; sStep[1] = 0;
            for (int d = 1; d <= D; d++) {
                if (sActive[d]) acc += (P.sg1[d] - P.sg0[d] + 3) >> 2;
                sStep[d + 1] = acc;
            }
            acc = 0;
            sAct4[0] = 0; sAct4[1] = 0;
            for (int d = 1; d <= D; d++) {
                if (sActive[d]) acc += (((P.row1[d] - P.row0[d]) >> 2) + 31) >> 5;
                sAct4[d + 1] = acc;
            }
            acc = 0;
            sXup[0] = 0; sXup[1] = 0;
            for (int d = 1; d <= D; d++) {
                if (sPend[d]) acc += (((P.row1[d] - P.row0[d]) >> 2) + 31) >> 5;
                sXup[d + 1] = acc;
            }
        }
        __syncthreads();
        const int cur = it & 1, nxt = cur ^ 1;
        float* const pNew = P.p + (i64)cur * P.pStride;            // p_k
        const float* const pOld = P.p + (i64)nxt * P.pStride;      // p_{k-1}
        double* dPAp = P.dots + cur * 32;         // kind 0
        double* dRRn = P.dots + cur * 32 + 16;    // kind 1 (this iteration's new r.r)

        // ---------------- phase C: p = r + beta p   (beta = 0 and p = 0 in the first iteration)
        const bool revC = P.zigzag && (phase++ & 1) != 0;
        {
            auto stepC = [&](int d, int i, const float4& rv, const float4& pv) {
                if (i < 0) return;
                const float be = sBeta[d];
                float4 o;
                o.x = __fadd_rn(rv.x, __fmul_rn(be, pv.x));
                o.y = __fadd_rn(rv.y, __fmul_rn(be, pv.y));
                o.z = __fadd_rn(rv.z, __fmul_rn(be, pv.z));
                o.w = __fadd_rn(rv.w, __fmul_rn(be, pv.w));
                *reinterpret_cast<float4*>(pNew + i) = o;
            };
            if (P.bulk) stream_pairs_bulk(sAct4, sAct4[D + 1], P.row0, P.row1, P.r, pOld, revC, wbuf, barS, phaseBits, gwarp, nwarps, lane, stepC);
            else stream_pairs(sAct4, sAct4[D + 1], P.row0, P.row1, P.r, pOld, revC, wbuf, gwarp, nwarps, lane, stepC);
        }
        tick(0);
        cg_sync<MG>(P, epoch, gen, -1, nullptr);
        tick(1);
        // ---------------- phase A: Ap = A p ; p.Ap over the flat step list of all active depths
        {
            const int total = sStep[D + 1];
            const bool rev = P.zigzag && (phase++ & 1) != 0;       // zig-zag: every phase sweeps memory in the direction opposite to the previous one
            // locate flat step u: depth and this quarter-warp's super-group (-1: past the end of the depth's
            // range).  (d, lo, hi) is carried along: consecutive steps of a warp move monotonically
            int ld = 1, llo = sStep[1], lhi = sStep[2];
            auto locate = [&](int t, int& sg) {
                const int u = rev ? total - 1 - t : t;
                while (u >= lhi) { ld++; llo = lhi; lhi = sStep[ld + 1]; }
                while (u < llo) { ld--; lhi = llo; llo = sStep[ld]; }
                sg = P.sg0[ld] + 4 * (u - llo) + q;
                if (sg >= P.sg1[ld]) sg = -1;
            };
            auto load_tab = [&](int sg, int (&e)[kRounds]) {
                if (sg >= 0) {
                    const int* row = P.tab4 + (i64)sg * 64;
                    const int4* tf = reinterpret_cast<const int4*>(row + (li >> 1) * 8);
                    const int4 a = __ldcs(tf), b = __ldcs(tf + 1), c = __ldcs(reinterpret_cast<const int4*>(row + 32 + li * 4));   // streamed once per sweep
                    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = b.x; e[5] = b.y; e[6] = b.z; e[7] = b.w; e[8] = c.x; e[9] = c.y; e[10] = c.z; e[11] = c.w;
                    if (kPrefetchSteps > 0 && (li & 3) == 0) {      // the table of a later step is on its way to L2 meanwhile
                        i64 sgp = (i64)sg + (rev ? -4 : 4) * (i64)kPrefetchSteps * nwarps;
                        sgp = sgp < 0 ? 0 : (sgp >= P.nSg ? P.nSg - 1 : sgp);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tab4 + sgp * 64 + li * 8));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kRounds; k++) e[k] = kAbsent;
                }
            };
            // source of a half-block: this rank's p, or the owner's through its peer-mapped arena
            auto stage = [&](const int (&e)[kRounds], int d, int buf) {
                const bool remote = MG && d >= P.shardFrom;
                const unsigned dst = cubeS + 4u * (unsigned)(buf * kWarpBufFloats);
#pragma unroll
                for (int r = 0; r < kRounds; r++) {
                    const int b = e[r];
                    const float* src = pNew + (b + (r >= 8 ? hEdge : hFull));
                    if (MG && remote && b >= 0 && (b < P.row0[d] || b >= P.row1[d])) {
                        int rk = 0;
                        while (rk + 1 < P.world && b >= P.rowLo[d][rk + 1]) rk++;
                        src = P.peerP[rk] + (i64)cur * P.pStride + (b + (r >= 8 ? hEdge : hFull));
                    }
                    cp_async16(dst + 4u * (unsigned)dstOff[r], src, b >= 0 ? 16 : 0);     // missing block: zero fill, no memory access
                }
            };
            // pending x updates (x += alpha_{k-1} p_{k-1}) ride along: one 32-lane batch of float4 per SpMV step,
            // loaded before the step's compute and finished after it, so their memory latency is hidden by
            // the FMAs; what is left (more batches than steps) is swept after the loop
            const int xTotal = sXup[D + 1];
            int xd = 1, xlo = sXup[1], xhi = sXup[2];
            int xb = gwarp;                       // this warp's next batch (flat index over the pending depths)
            auto x_locate = [&](int b, int& d, int& i) {        // -> depth and float index of this lane's float4 (i < 0: past the end)
                const int u = rev ? xTotal - 1 - b : b;
                while (u >= xhi) { xd++; xlo = xhi; xhi = sXup[xd + 1]; }
                while (u < xlo) { xd--; xhi = xlo; xlo = sXup[xd]; }
                d = xd;
                const int c4 = (u - xlo) * 32 + lane;
                i = c4 < ((P.row1[xd] - P.row0[xd]) >> 2) ? P.row0[xd] + 4 * c4 : -1;
            };
            double part = 0.0;
            int partDepth = 0;
            int t = gwarp;
            int eB[kRounds], eC[kRounds];
            int dA = 1, dB = 1, sgq;
            int rb0 = kAbsent, rb1 = kAbsent;          // bases of the blocks bz = 1 / bz = 2 of the lane's rows (current tile)
            if (t < total) {
                locate(t, sgq);
                dA = ld;
                load_tab(sgq, eB);
                stage(eB, dA, 0);
                rb0 = eB[0]; rb1 = eB[1];
            }
            cp_async_commit();
            bool haveB = t + nwarps < total;
            if (haveB) { locate(t + nwarps, sgq); dB = ld; load_tab(sgq, eB); }
            int buf = 0;
            for (; t < total; t += nwarps, buf ^= 1) {
                // next tile's copies into the other buffer, the table of the tile after that into registers
                bool haveC = false;
                int dC = dB;
                if (haveB) {
                    stage(eB, dB, buf ^ 1);
                    haveC = t + 2 * nwarps < total;
                    if (haveC) { locate(t + 2 * nwarps, sgq); dC = ld; load_tab(sgq, eC); }
                }
                cp_async_commit();
                int xi = -1, xdep = 1;
                float4 xpv = make_float4(0.f, 0.f, 0.f, 0.f), xxv = xpv;
                if (xb < xTotal) {
                    x_locate(xb, xdep, xi);
                    if (xi >= 0) { xpv = *reinterpret_cast<const float4*>(pOld + xi); xxv = *reinterpret_cast<const float4*>(P.x + xi); }
                }
                if (dA != partDepth) {               // flush the dot-product partial when the depth changes
                    if (partDepth) warp_add(part, sAcc, partDepth, lane);
                    part = 0.0;
                    partDepth = dA;
                }
                const float4 sv = *reinterpret_cast<const float4*>(&sSt[dA][0]);
                cp_async_wait<1>();
                __syncwarp();
                {
                    const float* cb = cube0 + buf * kWarpBufFloats;
                    float acc[2][4], pc[2][4];
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int z = 0; z < 4; z++) { acc[a][z] = 0.f; pc[a][z] = 0.f; }
#pragma unroll
                    for (int dxp = 0; dxp < 4; dxp++)
#pragma unroll
                        for (int dy = 0; dy < 3; dy++) {
                            // z = 2..5 as two 64-bit loads (conflict free per half-warp); z = 1 and z = 6 as
                            // 32-bit loads: the lower half-warp reads z = 1 while the upper one reads z = 6 and
                            // vice versa, so the 32 lanes always hit banks of both parities (1 wavefront each)
                            const float* col = cb + colOff[dy] + 4 * dxp;
                            float w[8];
                            const float ea = cb[colA[dy] + 4 * dxp], eb = cb[colB[dy] + 4 * dxp];
                            const float2 v12 = *reinterpret_cast<const float2*>(col + 4 * kUnitBz);
                            const float2 v34 = *reinterpret_cast<const float2*>(col + 8 * kUnitBz);
                            w[0] = 0.f; w[7] = 0.f;
                            w[1] = upper ? eb : ea; w[6] = upper ? ea : eb;
                            w[2] = v12.x; w[3] = v12.y; w[4] = v34.x; w[5] = v34.y;
#pragma unroll
                            for (int xr = 0; xr < 2; xr++) {
                                const int dx = dxp - 1 - xr;
                                if (dx < -1 || dx > 1) continue;
                                const int ty = (dx != 0) + (dy != 1);
                                const float se = ty == 0 ? sv.y : (ty == 1 ? sv.z : sv.w);     // dz != 0
                                const float sc = ty == 0 ? sv.x : (ty == 1 ? sv.y : sv.z);     // dz == 0
#pragma unroll
                                for (int zr = 0; zr < 4; zr++) {
                                    float a = acc[xr][zr];
                                    a = __fmaf_rn(se, w[1 + zr], a);
                                    a = __fmaf_rn(sc, w[2 + zr], a);
                                    a = __fmaf_rn(se, w[3 + zr], a);
                                    acc[xr][zr] = a;
                                    if (dx == 0 && dy == 1) pc[xr][zr] = w[2 + zr];
                                }
                            }
                        }
                    if (rb0 >= 0) {
                        float* o = P.Ap + rb0 + outOff;
                        *reinterpret_cast<float2*>(o) = make_float2(acc[0][0], acc[0][1]);
                        *reinterpret_cast<float2*>(o + 4) = make_float2(acc[1][0], acc[1][1]);
                        part += (double)(pc[0][0] * acc[0][0]);
                        part += (double)(pc[0][1] * acc[0][1]);
                        part += (double)(pc[1][0] * acc[1][0]);
                        part += (double)(pc[1][1] * acc[1][1]);
                    }
                    if (rb1 >= 0) {
                        float* o = P.Ap + rb1 + outOff;
                        *reinterpret_cast<float2*>(o) = make_float2(acc[0][2], acc[0][3]);
                        *reinterpret_cast<float2*>(o + 4) = make_float2(acc[1][2], acc[1][3]);
                        part += (double)(pc[0][2] * acc[0][2]);
                        part += (double)(pc[0][3] * acc[0][3]);
                        part += (double)(pc[1][2] * acc[1][2]);
                        part += (double)(pc[1][3] * acc[1][3]);
                    }
                }
                if (xi >= 0) {
                    const float al = sAlpha[xdep];
                    xxv.x = __fmaf_rn(al, xpv.x, xxv.x); xxv.y = __fmaf_rn(al, xpv.y, xxv.y); xxv.z = __fmaf_rn(al, xpv.z, xxv.z); xxv.w = __fmaf_rn(al, xpv.w, xxv.w);
                    *reinterpret_cast<float4*>(P.x + xi) = xxv;
                }
                xb += nwarps;
                __syncwarp();      // all lanes are done with this buffer before the next copies land in it
                rb0 = eB[0]; rb1 = eB[1];
                dA = dB; dB = dC;
                haveB = haveC;
#pragma unroll
                for (int k = 0; k < kRounds; k++) eB[k] = eC[k];
            }
            cp_async_wait<0>();
            if (partDepth) warp_add(part, sAcc, partDepth, lane);
            for (; xb < xTotal; xb += nwarps) {      // x batches beyond this warp's SpMV steps
                int xi, xdep;
                x_locate(xb, xdep, xi);
                if (xi >= 0) {
                    const float al = sAlpha[xdep];
                    const float4 pv = *reinterpret_cast<const float4*>(pOld + xi);
                    float4 xv = *reinterpret_cast<const float4*>(P.x + xi);
                    xv.x = __fmaf_rn(al, pv.x, xv.x); xv.y = __fmaf_rn(al, pv.y, xv.y); xv.z = __fmaf_rn(al, pv.z, xv.z); xv.w = __fmaf_rn(al, pv.w, xv.w);
                    *reinterpret_cast<float4*>(P.x + xi) = xv;
                }
            }
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dPAp[tid], sAcc[tid]);
        tick(2);
        cg_sync<MG>(P, epoch, gen, 0, dPAp);
        tick(3);
        // both accumulators of the NEXT iteration are zeroed here: every block has passed this
        // iteration's syncs, hence finished reading them after the previous iteration's syncs
        if (blockIdx.x == 0 && tid < 32) P.dots[nxt * 32 + tid] = 0.0;
        if (tid <= D) sAcc[tid] = 0.0;
        if (tid >= 1 && tid <= D && sActive[tid]) sAlpha[tid] = (float)((double)sR1[tid] / cg_total<MG>(P, epoch, tid, dPAp));
        __syncthreads();
        // ---------------- phase B: r -= alpha Ap ; r.r   (x += alpha p is applied under the next SpMV, or by the sweep after the loop)
        const bool revB = P.zigzag && (phase++ & 1) != 0;
        {
            double part = 0.0;
            int partDepth = 0;
            auto stepB = [&](int d, int i, const float4& av, const float4& rv) {
                if (d != partDepth) {           // (warp-uniform: a batch never mixes depths)
                    if (partDepth) warp_add(part, sAcc, partDepth, lane);
                    part = 0.0;
                    partDepth = d;
                }
                if (i < 0) return;
                const float al = sAlpha[d];
                float4 ro;
                ro.x = __fmaf_rn(-al, av.x, rv.x); ro.y = __fmaf_rn(-al, av.y, rv.y); ro.z = __fmaf_rn(-al, av.z, rv.z); ro.w = __fmaf_rn(-al, av.w, rv.w);
                *reinterpret_cast<float4*>(P.r + i) = ro;
                part += (double)(ro.x * ro.x);
                part += (double)(ro.y * ro.y);
                part += (double)(ro.z * ro.z);
                part += (double)(ro.w * ro.w);
            };
            if (P.bulk) stream_pairs_bulk(sAct4, sAct4[D + 1], P.row0, P.row1, P.Ap, P.r, revB, wbuf, barS, phaseBits, gwarp, nwarps, lane, stepB);
            else stream_pairs(sAct4, sAct4[D + 1], P.row0, P.row1, P.Ap, P.r, revB, wbuf, gwarp, nwarps, lane, stepB);
            if (partDepth) warp_add(part, sAcc, partDepth, lane);
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dRRn[tid], sAcc[tid]);
        tick(4);
        cg_sync<MG>(P, epoch, gen, 1, dRRn);
        tick(5);
        if (tid >= 1 && tid <= D && sActive[tid]) {
            float r0 = sR1[tid], r1 = (float)cg_total<MG>(P, epoch, tid, dRRn);
            sR1[tid] = r1;
            int k = sIter[tid] + 1;
            sIter[tid] = k;
            sBeta[tid] = r1 / r0;
            sActive[tid] = (r1 > P.tol2 && k <= P.maxIter) ? 1 : 0;
            sPend[tid] = 1;                 // this iteration's alpha p is still to be added to x
        } else if (tid >= 1 && tid <= D) {
            sPend[tid] = 0;                 // a depth that had stopped before: its last update went in under this iteration's SpMV
        }
        __syncthreads();
    }
    // ---- the last alpha p of every depth that was still iterating (p_k sits in the buffer of the last iteration)
    {
        const int itLast = it - 1;
        const float* pLast = P.p + (i64)(itLast & 1) * P.pStride;
        for (int d = 1; d <= D; d++) {
            if (!sPend[d]) continue;
            const float al = sAlpha[d];
            const int i0 = P.row0[d], n4 = (P.row1[d] - i0) >> 2;
            for (int c = gthread; c < n4; c += nthreads) {
                const int ik = i0 + 4 * c;
                const float4 pv = *reinterpret_cast<const float4*>(pLast + ik);
                float4 xv = *reinterpret_cast<const float4*>(P.x + ik);
                xv.x = __fmaf_rn(al, pv.x, xv.x); xv.y = __fmaf_rn(al, pv.y, xv.y); xv.z = __fmaf_rn(al, pv.z, xv.z); xv.w = __fmaf_rn(al, pv.w, xv.w);
                *reinterpret_cast<float4*>(P.x + ik) = xv;
            }
        }
    }
    if (blockIdx.x == 0 && tid >= 1 && tid <= D && (P.onlyDepth < 0 || tid == P.onlyDepth)) { P.itersOut[tid] = sIter[tid] - 1; P.resOut[tid] = sR1[tid]; }
    if (blockIdx.x == 0 && tid == 0) P.itersOut[15] = (int)epoch;
}

// OPT-IN cascadic mode (the reference solves its D+1 systems independently, main.cu:1237-1329): before depth d is
// solved, the contribution of the coarser solutions is taken out of its right-hand side,
//   b'_o = b_o - sum_{e < d} sum_{n in N27(ancestor_e(o))} L(o, n) x_n,   L = D2x FFy FFz + FFx D2y FFz + FFx FFy D2z
// with the cross-depth 1-D integrals of bspline_host (ffX / d2X, indexed by u = off_o - 2^(d-e) (off_n - 1) per axis; computed in double
// and narrowed to float like the reference's same-depth entries, main.cu:1199-1207).  One thread per node of depth d.
__global__ void __launch_bounds__(256) k_cascadic_rhs(int d, int base, int count, const int* __restrict__ parent, const int* __restrict__ neighs, const ushort4* __restrict__ offs,
                                                      const double* __restrict__ ffX, const double* __restrict__ d2X, const int* __restrict__ crossOff /* row d of the table */,
                                                      const float* __restrict__ x, const float* __restrict__ bIn, float* __restrict__ bOut) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < count; l += gridDim.x * blockDim.x) {
        const int o = base + l;
        const ushort4 oo = offs[o];
        double acc = 0.0;
        int a = parent[o];
        for (int e = d - 1; e >= 0 && a >= 0; --e) {
            const int k = 1 << (d - e);
            const ushort4 ao = offs[a];
            const int rx = (int)oo.x - (int)ao.x * k, ry = (int)oo.y - (int)ao.y * k, rz = (int)oo.z - (int)ao.z * k;      // in [0, k)
            const double* F = ffX + crossOff[e];
            const double* S = d2X + crossOff[e];
            double fx[3], fy[3], fz[3], sx[3], sy[3], sz[3];
#pragma unroll
            for (int t = 0; t < 3; t++) {            // neighbour direction t - 1 along the axis: u = r + (1 - (t - 1)) k
                const int ux = rx + (2 - t) * k, uy = ry + (2 - t) * k, uz = rz + (2 - t) * k;
                fx[t] = F[ux]; fy[t] = F[uy]; fz[t] = F[uz];
                sx[t] = S[ux]; sy[t] = S[uy]; sz[t] = S[uz];
            }
            const int* nb = neighs + 27 * (i64)a;
#pragma unroll
            for (int j = 0; j < 27; j++) {
                const int n = nb[j];
                if (n < 0) continue;
                const int jx = j / 9, jy = (j / 3) % 3, jz = j % 3;
                const float L = (float)(sx[jx] * fy[jy] * fz[jz] + fx[jx] * sy[jy] * fz[jz] + fx[jx] * fy[jy] * sz[jz]);
                acc += (double)(L * x[n]);
            }
            a = parent[a];
        }
        bOut[o] = bIn[o] - (float)acc;
    }
}

int build_cg_table(Context& c) {
    PRB_TRY(c.sgTab4.alloc(64 * (size_t)c.nSg, c.stream));
    PRB_LAUNCH(c, k_sg_table4, grid_for(c, (i64)c.nSg * 64, 256), 256, 0, c.sgTab.p, (i64)c.nSg, c.sgTab4.p);
    return PRB_OK;
}

int stage_solve(Context& c) {
    const int D = c.D, M = c.M;
    cudaStream_t st = c.stream;
    // vectors are padded by 7 floats so that node 1 (the first sibling block) is 32-byte aligned
    const size_t padN = ((size_t)M + 8 + 7) & ~(size_t)7;      // a multiple of 8: the second p buffer keeps the 32-byte alignment of the sibling blocks
    const bool mg = c.mg.active();
    float* pBuf = nullptr;
    DBuf<float> r, p, Ap, resOut;
    if (mg) {
        // x and p live in the peer-mapped arena: p is read by the other ranks during the solve,
        // x is collected from them afterwards
        for (int q = 0; q < c.mg.world; q++)
            if (!c.mg.peer[q]) { set_error("multi-GPU: peer arenas not exchanged (prb_mg_set_peer)"); return PRB_ERR_STATE; }
        if (!c.mgX) {
            c.mgX = c.mg.alloc<float>(padN, &c.mgXOff);
            c.mgP = c.mg.alloc<float>(2 * padN, &c.mgPOff);
            if (!c.mgX || !c.mgP) { set_error("multi-GPU arena too small for the CG vectors (prb_mg_init arena_bytes)"); return PRB_ERR_NOMEM; }
        }
        c.xv = c.mgX + 7;
        pBuf = c.mgP;
    } else {
        PRB_TRY(c.x.alloc(padN, st));
        c.xv = c.x.p + 7;
        PRB_TRY(p.alloc(2 * padN, st));
        pBuf = p.p;
    }
    DBuf<double> dots;
    DBuf<int> itersOut;
    PRB_TRY(r.alloc(padN, st));
    PRB_TRY(Ap.alloc(padN, st));
    PRB_TRY(dots.alloc(96 + 32, st));                   // + the arrival counter and the release flag of cg_sync, one 128-byte line each
    PRB_TRY(itersOut.alloc(16, st));
    PRB_TRY(resOut.alloc(16, st));
    PRB_CUDA(cudaMemsetAsync(dots.p, 0, (96 + 32) * sizeof(double), st));
    PRB_CUDA(cudaMemsetAsync(itersOut.p, 0, 16 * sizeof(int), st));
    CgParams P;
    P.D = D;
    // super-groups: sg 0 = depth 1; depth d >= 2 owns the super-groups 1 + (sibling groups of depth d-1)
    P.sgStart[0] = 0;
    P.sgStart[1] = 0;
    for (int d = 2; d <= D + 1; d++) P.sgStart[d] = 1 + (c.base[d - 1] - 1) / 8;
    P.tab4 = c.sgTab4.p; P.nSg = c.nSg; P.zigzag = c.cgZigzag; P.bulk = c.cgBulk; P.stencil = c.dStencil.p; P.b = c.divgv;
    P.x = c.xv; P.r = r.p + 7; P.p = pBuf + 7; P.pStride = (i64)padN; P.Ap = Ap.p + 7;
    P.world = c.mg.world; P.rank = c.mg.rank; P.shardFrom = mg ? c.shardFrom : D + 1;
    for (int d = 0; d <= D + 1; d++) {
        const bool sh = mg && d >= c.shardFrom && d <= D;
        P.sg0[d] = sh ? c.sgLo[d][c.mg.rank] : P.sgStart[d];
        P.sg1[d] = sh ? c.sgLo[d][c.mg.rank + 1] : (d <= D ? P.sgStart[d + 1] : P.sgStart[d]);
        P.row0[d] = sh ? c.rowLo[d][c.mg.rank] : (d <= D ? c.base[d] : 0);
        P.row1[d] = sh ? c.rowLo[d][c.mg.rank + 1] : (d <= D ? c.base[d + 1] : 0);
        for (int q = 0; q <= kMaxRanks; q++) P.rowLo[d][q] = c.rowLo[d][q];
    }
    for (int q = 0; q < kMaxRanks; q++) P.peerP[q] = (mg && q < c.mg.world) ? (const float*)(c.mg.peer[q] + c.mgPOff) + 7 : nullptr;
    P.mg = c.mg.dev();
    if (mg) PRB_TRY(mg_barrier(c));        // every rank's previous use of the arena buffers is over before anyone writes p / x again
    P.epoch0 = c.mg.cgEpoch;
    P.barCount = reinterpret_cast<unsigned*>(dots.p + 96);
    P.barRelease = P.barCount + 32;       // (its own 128-byte line)
    P.dots = dots.p; P.itersOut = itersOut.p; P.resOut = resOut.p;
    DBuf<long long> phaseNs;
    P.phaseNs = nullptr;
    if (c.cgTiming) {
        PRB_TRY(phaseNs.alloc(8, st));
        PRB_CUDA(cudaMemsetAsync(phaseNs.p, 0, 8 * sizeof(long long), st));
        P.phaseNs = phaseNs.p;
    }
    float tol = (float)c.cgTol;
    P.tol2 = tol * tol;
    P.maxIter = c.cgMaxIter;
    const size_t dynSmem = (size_t)kCgWarps * 2 * kWarpBufFloats * sizeof(float);
    const void* kern = mg ? (const void*)k_cg_all_depths<true> : (const void*)k_cg_all_depths<false>;
    PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
    PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int perSM = 0;
    PRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kCgBlock, dynSmem));
    if (perSM < 1) { set_error("CG kernel does not fit on an SM"); return PRB_ERR_CUDA; }
    int gridSize = c.smCount * perSM;
    i64 maxTiles = (P.sgStart[D + 1] + 3) / 4 + D;      // warp-steps of one SpMV sweep (4 super-groups each)
    i64 needBlocks = (maxTiles + kCgWarps - 1) / kCgWarps;
    if (gridSize > needBlocks) gridSize = (int)(((needBlocks + c.smCount - 1) / c.smCount) * c.smCount);   // small problems: fewer CTAs, cheaper grid syncs
    if (gridSize > c.smCount * perSM) gridSize = c.smCount * perSM;
    if (gridSize < 1) gridSize = 1;
    mark(c, "solve:setup");
    void* args[] = {(void*)&P};
    P.onlyDepth = -1;
    if (!c.cascadic) {
        PRB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(gridSize), dim3(kCgBlock), args, dynSmem, st));
        c.launches++;
    } else {
        // cascadic mode: depth by depth, coarse to fine; the right-hand side of depth d first loses what the coarser solutions explain
        if (mg) { set_error("the cascadic mode is single-GPU"); return PRB_ERR_STATE; }
        PRB_TRY(c.bCas.alloc((size_t)M + 16, st));
        float* bc = c.bCas.p + 7;
        PRB_CUDA(cudaMemcpyAsync(bc, c.divgv, sizeof(float) * (size_t)M, cudaMemcpyDeviceToDevice, st));
        P.b = bc;
        for (int d = 0; d <= D; d++) {
            if (d >= 1)
                PRB_LAUNCH(c, k_cascadic_rhs, grid_for(c, c.cnt[d], 256), 256, 0, d, c.base[d], c.cnt[d], c.parent.p, c.neighs.p, c.offs.p, c.dFfX.p, c.dD2X.p,
                           c.dCrossOff.p + (size_t)d * (D + 1), c.xv, c.divgv, bc);
            PRB_CUDA(cudaMemsetAsync(dots.p, 0, (96 + 32) * sizeof(double), st));
            P.onlyDepth = d;
            PRB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(gridSize), dim3(kCgBlock), args, dynSmem, st));
            c.launches++;
        }
    }
    mark(c, "solve:kernel");
    int hIters[16];
    if (c.cgTiming) PRB_CUDA(cudaMemcpyAsync(c.cgPhaseNs, phaseNs.p, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaMemcpyAsync(hIters, itersOut.p, sizeof(hIters), cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    c.cgRowIters = 0;
    for (int d = 0; d <= D; d++) { c.cgIters[d] = hIters[d]; c.cgRowIters += (i64)(P.row1[d] - P.row0[d]) * hIters[d]; }
    if (mg) {
        c.mg.cgEpoch = (unsigned)hIters[15];
        // collect the other ranks' parts of the solution (pull over NVLink), then let nobody run ahead
        PRB_TRY(mg_barrier(c));
        {
            const void* src[32];
            void* dst[32];
            size_t bytes[32];
            int n = 0;
            for (int qi = 1; qi < c.mg.world; qi++) {     // start with the next rank: the peers are not all pulled from in the same order
                const int q = (c.mg.rank + qi) % c.mg.world;
                const float* px = (const float*)(c.mg.peer[q] + c.mgXOff) + 7;
                // one segment per peer when the depths' ranges would not fit the segment list (they are disjoint: fall back to a launch per peer)
                for (int d = c.shardFrom; d <= D; d++) {
                    const size_t cnt = (size_t)(c.rowLo[d][q + 1] - c.rowLo[d][q]);
                    if (!cnt) continue;
                    if (n == 32) { PRB_TRY(mg_pull(c, n, src, dst, bytes)); n = 0; }
                    src[n] = px + c.rowLo[d][q]; dst[n] = c.xv + c.rowLo[d][q]; bytes[n] = cnt * sizeof(float); n++;
                }
            }
            PRB_TRY(mg_pull(c, n, src, dst, bytes));
        }
        PRB_TRY(mg_barrier(c));
        int err = 0;
        PRB_CUDA(cudaMemcpyAsync(&err, &((MgHeader*)c.mg.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        if (err) { set_error("multi-GPU solve: timed out waiting for a peer"); return PRB_ERR_CUDA; }
    }
    mark(c, "solve:gathered");
    r.release(); p.release(); Ap.release(); dots.release(); itersOut.release(); resOut.release();
    return PRB_OK;
}

}  // namespace prb
