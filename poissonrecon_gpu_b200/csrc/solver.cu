// Per-depth conjugate gradient, all depths in ONE persistent cooperative kernel.
//
// Replaces GenerateSingleNodeLaplacian + scan + 2x copy_if CSR assembly (main.cu:1143-1329) and
// the cooperative CG sample kernel (CG_CUDA.cuh:186-324, 344-509).  The reference solves the
// D+1 independent same-depth systems one after another, each with its own CSR matrix in managed
// memory, 7 grid syncs per iteration and host loops over managed arrays around the launch.
// Here:
//   * the matrix is never formed: rows are the translation-invariant 27-point stencil
//     (4 distinct values per depth) applied through the super-group table sgTab[sg][64]
//     (octree.cu k_sg_table): the up to 64 rows under one node Q (8 sibling groups) and all
//     their neighbours live in the 4x4x4 cube of 8-row blocks around Q's children
//     (256 B of topology per <= 64 rows instead of 216 B per ROW of CSR);
//   * the depths are independent, so they all iterate in lock-step inside one
//     launch: one iteration of the kernel = one CG iteration of every still-active depth, with
//     per-depth alpha / beta / residual and per-depth stopping.  Grid syncs per solve drop from
//     7 * sum_d iters_d to 3 * max_d iters_d;
//   * SpMV: a warp owns a super-group per step.  The 64 blocks (an 8x8x8 cube of p values, 2 KB)
//     are copied global -> shared with 16-byte cp.async into a double buffer, one tile ahead of
//     the compute, and the table two tiles ahead, so the gather latency is off the critical path;
//     every lane then produces two rows (a z pair) from a 3x3x4 register window read with
//     conflict-free LDS (padded block-major layout; the two half-warps walk z in opposite
//     directions so that they always hit different banks);
//   * the row sum runs over the neighbour slots in order j = 0..26 with FMAs, exactly the
//     reference's CSR order (absent neighbours contribute an exact +0), so A*p is bit-identical;
//     dots are float products accumulated in double (CG_CUDA.cuh:217-220); alpha, beta are float.
// Algorithmic bytes per row per iteration: 57.5 B
//   (p=r+beta*p: 12, SpMV: 13.5 + 4 + 4, x/r update: 24).
#include "common.cuh"
#include "mg_device.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace prb {

constexpr int kCgBlock = 256;
constexpr int kCgWarps = kCgBlock / 32;
// shared cube of one super-group: block (bx,by,bz) at bx*kSX + by*kSY + bz*8 floats, 8 floats per
// block in child-code order; the paddings make both the 16-byte staging stores and the
// per-lane window loads bank-conflict free (see k_cg_all_depths)
constexpr int kSX = 200, kSY = 48, kCube = 4 * kSX;

struct CgParams {
    int D;
    int gbase[kMaxDepth + 2];     // first sibling group of depth d (groups cover nodes 1..M-1), gbase[D+1] = total
    int sgStart[kMaxDepth + 2];   // first super-group of depth d (d = 1..D), sgStart[D+1] = total
    const int* sgTab;
    const float* stencil;          // [D+1][27]
    const float* b;                // divergence
    // vectors indexed by node id; node 1 sits on a 32-byte boundary (pointer = allocation + 7)
    float* x;
    float* r;
    float* p;
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
};

// grid-wide (world == 1) or box-wide barrier.  kind: -1 none, 0/1 = publish this rank's per-depth
// partial sums `local[1..D]` to every rank's slot table before signalling.
template <bool MG>
__device__ __forceinline__ void cg_sync(cg::grid_group& grid, const CgParams& P, unsigned& epoch, int parity, int kind, const double* local) {
    grid.sync();
    if (!MG) return;
    epoch++;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (kind >= 0)
            for (int r = 0; r < P.world; r++)
                for (int d = 1; d <= P.D; d++) P.mg.peerHdr[r]->slots[parity][P.rank][kind * 16 + d] = local[d];
        mg_signal_wait(P.mg, epoch);
    }
    grid.sync();
}
// total of a per-depth dot product after cg_sync
template <bool MG>
__device__ __forceinline__ double cg_total(const CgParams& P, int parity, int kind, int d, const double* local) {
    if (!MG) return local[d];
    const volatile double* s = &P.mg.hdr->slots[parity][0][kind * 16 + d];
    if (d < P.shardFrom) return s[0];
    double t = 0.0;
    for (int r = 0; r < P.world; r++) t += s[r * 32];
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

extern __shared__ __align__(16) float sDyn[];   // [kCgWarps][2][kCube]

template <bool MG>
__global__ void __launch_bounds__(kCgBlock) k_cg_all_depths(const __grid_constant__ CgParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sSt[kMaxDepth + 1][4];
    __shared__ double sAcc[kMaxDepth + 1];
    __shared__ float sR1[kMaxDepth + 1], sAlpha[kMaxDepth + 1], sBeta[kMaxDepth + 1];
    __shared__ int sActive[kMaxDepth + 1], sIter[kMaxDepth + 1];
    const int D = P.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = blockIdx.x * kCgWarps + warp, nwarps = gridDim.x * kCgWarps;
    const int gthread = blockIdx.x * kCgBlock + tid, nthreads = gridDim.x * kCgBlock;
    if (tid < (D + 1) * 4) {
        // one representative per number of off-centre axes: j = 13, 14, 17, 26
        const int rep[4] = {13, 14, 17, 26};
        sSt[tid >> 2][tid & 3] = P.stencil[(tid >> 2) * 27 + rep[tid & 3]];
    }
    if (tid <= D) sAcc[tid] = 0.0;
    __syncthreads();

    // ---- depth 0: a 1x1 system, solved by one thread with the same recurrences
    if (blockIdx.x == 0 && tid == 0) {
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
            const int i0 = P.row0[d], i1 = P.row1[d];
            double part = 0.0;
            for (int i = i0 + gthread; i < i1; i += nthreads) {
                float bv = P.b[i];
                P.x[i] = 0.f; P.r[i] = bv; P.p[i] = 0.f;
                part += (double)(bv * bv);
            }
            warp_add(part, sAcc, d, lane);
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&P.dots[64 + tid], sAcc[tid]);   // dedicated init buffer
    }
    unsigned epoch = P.epoch0;
    cg_sync<MG>(grid, P, epoch, 0, 0, P.dots + 64);
    if (tid >= 1 && tid <= D) {
        float r1 = (float)cg_total<MG>(P, 0, 0, tid, P.dots + 64);
        sR1[tid] = r1; sIter[tid] = 1; sBeta[tid] = 0.f; sAlpha[tid] = 0.f;
        sActive[tid] = (r1 > P.tol2 && 1 <= P.maxIter) ? 1 : 0;
    }
    __syncthreads();

    // ---- per-lane constants of the SpMV
    float* const cube0 = sDyn + (size_t)warp * 2 * kCube;
    const unsigned cubeS = (unsigned)__cvta_generic_to_shared(cube0);
    // staging: copy task t = lane + 32k (k = 0..3) moves half-block (blk = t>>1, half = t&1)
    int stOff[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int t = lane + 32 * k, blk = t >> 1, half = t & 1;
        stOff[k] = (blk >> 4) * kSX + ((blk >> 2) & 3) * kSY + (blk & 3) * 8 + half * 4;
    }
    // compute: lane -> (kz, X, Y); it produces the rows at cube node (2+X, 2+Y, 2+2kz + {0,1})
    const int kz = lane >> 4, X = (lane >> 2) & 3, Y = lane & 3;
    const int pb = (1 + (X >> 1)) * 16 + (1 + (Y >> 1)) * 4 + (1 + kz);      // cube block of the rows' parent
    const int rowIn = ((X & 1) << 2) | ((Y & 1) << 1);                         // child code of the z pair's first row
    int ax[3], ay[3], zo[4];
#pragma unroll
    for (int t = 0; t < 3; t++) {
        int xx = 1 + X + t, yy = 1 + Y + t;
        ax[t] = (xx >> 1) * kSX + ((xx & 1) << 2);
        ay[t] = (yy >> 1) * kSY + ((yy & 1) << 1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        // half-warp 0 reads z = z0-1 .. z0+2 upwards, half-warp 1 downwards: at every step the two
        // halves touch z of opposite parity, i.e. different banks
        int z = kz ? (2 + 2 * kz + 2 - i) : (1 + i);
        zo[i] = (z >> 1) * 8 + (z & 1);
    }

    for (int it = 1;; it++) {
        int anyActive = 0;
        for (int d = 1; d <= D; d++) anyActive |= sActive[d];
        if (!anyActive) break;
        if (tid <= D) sAcc[tid] = 0.0;
        __syncthreads();
        const int cur = it & 1, nxt = cur ^ 1;
        double* dPAp = P.dots + cur * 32;         // kind 0
        double* dRRn = P.dots + cur * 32 + 16;    // kind 1 (this iteration's new r.r)

        // ---------------- phase C: p = r + beta p   (beta = 0 and p = 0 in the first iteration)
        for (int d = 1; d <= D; d++) {
            if (!sActive[d]) continue;
            const float be = sBeta[d];
            const int i1 = P.row1[d];
            for (int i = P.row0[d] + 4 * gthread; i < i1; i += 4 * nthreads) {
                float4 rv = *reinterpret_cast<const float4*>(P.r + i);
                float4 pv = *reinterpret_cast<const float4*>(P.p + i);
                pv.x = __fadd_rn(rv.x, __fmul_rn(be, pv.x));
                pv.y = __fadd_rn(rv.y, __fmul_rn(be, pv.y));
                pv.z = __fadd_rn(rv.z, __fmul_rn(be, pv.z));
                pv.w = __fadd_rn(rv.w, __fmul_rn(be, pv.w));
                *reinterpret_cast<float4*>(P.p + i) = pv;
            }
        }
        cg_sync<MG>(grid, P, epoch, cur, -1, nullptr);
        // ---------------- phase A: Ap = A p ; p.Ap
        for (int d = 1; d <= D; d++) {
            if (!sActive[d]) continue;
            const float s0 = sSt[d][0], s1 = sSt[d][1], s2 = sSt[d][2], s3 = sSt[d][3];
            const int t1 = P.sg1[d];
            int t = P.sg0[d] + gwarp;
            const bool remote = MG && d >= P.shardFrom;
            const int myLo = P.row0[d], myHi = P.row1[d];
            // source of a block: this rank's p, or the owner's through its peer-mapped arena
            auto src = [&](int base) -> const float* {
                if (!remote || (base >= myLo && base < myHi)) return P.p + base;
                int r = 0;
                while (r + 1 < P.world && base >= P.rowLo[d][r + 1]) r++;
                return P.peerP[r] + base;
            };
            double part = 0.0;
            // table registers of the current tile (A), the next (B) and the one after (C)
            int a0 = -1, a1 = -1, b0 = -1, b1 = -1;
            if (t < t1) { a0 = P.sgTab[64 * (i64)t + lane]; a1 = P.sgTab[64 * (i64)t + 32 + lane]; }
            if (t + nwarps < t1) { b0 = P.sgTab[64 * (i64)(t + nwarps) + lane]; b1 = P.sgTab[64 * (i64)(t + nwarps) + 32 + lane]; }
            int buf = 0;
            if (t < t1) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int base = __shfl_sync(0xffffffffu, k < 2 ? a0 : a1, ((lane >> 1) + 16 * k) & 31);
                    cp_async16(cubeS + 4u * (unsigned)stOff[k], base >= 0 ? (const void*)(src(base) + 4 * (lane & 1)) : (const void*)(P.p + 1), base >= 0 ? 16 : 0);
                }
            }
            cp_async_commit();
            for (; t < t1; t += nwarps, buf ^= 1) {
                // next tile's copies into the other buffer, the table of the tile after that into registers
                int c0 = -1, c1 = -1;
                if (t + nwarps < t1) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        int base = __shfl_sync(0xffffffffu, k < 2 ? b0 : b1, ((lane >> 1) + 16 * k) & 31);
                        cp_async16(cubeS + 4u * (unsigned)((buf ^ 1) * kCube + stOff[k]), base >= 0 ? (const void*)(src(base) + 4 * (lane & 1)) : (const void*)(P.p + 1),
                                   base >= 0 ? 16 : 0);
                    }
                    if (t + 2 * nwarps < t1) { c0 = P.sgTab[64 * (i64)(t + 2 * nwarps) + lane]; c1 = P.sgTab[64 * (i64)(t + 2 * nwarps) + 32 + lane]; }
                }
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const int r0 = __shfl_sync(0xffffffffu, a0, pb & 31), r1 = __shfl_sync(0xffffffffu, a1, pb & 31);
                const int rowBase = pb < 32 ? r0 : r1;
                if (rowBase >= 0) {
                    const float* cb = cube0 + buf * kCube;
                    float acc0 = 0.f, acc1 = 0.f, pc0 = 0.f, pc1 = 0.f;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
#pragma unroll
                        for (int dy = 0; dy < 3; dy++) {
                            const float* q = cb + ax[dx] + ay[dy];
                            float t0 = q[zo[0]], t1v = q[zo[1]], t2 = q[zo[2]], t3 = q[zo[3]];
                            const float w0 = kz ? t3 : t0, w1 = kz ? t2 : t1v, w2 = kz ? t1v : t2, w3 = kz ? t0 : t3;
                            const int ty = (dx != 1) + (dy != 1);
                            const float se = ty == 0 ? s1 : (ty == 1 ? s2 : s3);     // dz != 0
                            const float sc = ty == 0 ? s0 : (ty == 1 ? s1 : s2);     // dz == 0
                            acc0 = __fmaf_rn(se, w0, acc0); acc0 = __fmaf_rn(sc, w1, acc0); acc0 = __fmaf_rn(se, w2, acc0);
                            acc1 = __fmaf_rn(se, w1, acc1); acc1 = __fmaf_rn(sc, w2, acc1); acc1 = __fmaf_rn(se, w3, acc1);
                            if (dx == 1 && dy == 1) { pc0 = w1; pc1 = w2; }
                        }
                    *reinterpret_cast<float2*>(P.Ap + rowBase + rowIn) = make_float2(acc0, acc1);
                    part += (double)(pc0 * acc0);
                    part += (double)(pc1 * acc1);
                }
                __syncwarp();      // all lanes are done with this buffer before the next copies land in it
                a0 = b0; a1 = b1; b0 = c0; b1 = c1;
            }
            cp_async_wait<0>();
            warp_add(part, sAcc, d, lane);
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dPAp[tid], sAcc[tid]);
        cg_sync<MG>(grid, P, epoch, cur, 0, dPAp);
        // both accumulators of the NEXT iteration are zeroed here: every block has passed this
        // iteration's syncs, hence finished reading them after the previous iteration's syncs
        if (blockIdx.x == 0 && tid < 32) P.dots[nxt * 32 + tid] = 0.0;
        if (tid <= D) sAcc[tid] = 0.0;
        if (tid >= 1 && tid <= D && sActive[tid]) sAlpha[tid] = (float)((double)sR1[tid] / cg_total<MG>(P, cur, 0, tid, dPAp));
        __syncthreads();
        // ---------------- phase B: x += alpha p ; r -= alpha Ap ; r.r
        for (int d = 1; d <= D; d++) {
            if (!sActive[d]) continue;
            const float al = sAlpha[d];
            const int i1 = P.row1[d];
            double part = 0.0;
            for (int i = P.row0[d] + 4 * gthread; i < i1; i += 4 * nthreads) {
                float4 pv = *reinterpret_cast<const float4*>(P.p + i);
                float4 av = *reinterpret_cast<const float4*>(P.Ap + i);
                float4 xv = *reinterpret_cast<const float4*>(P.x + i);
                float4 rv = *reinterpret_cast<const float4*>(P.r + i);
                xv.x = __fmaf_rn(al, pv.x, xv.x); xv.y = __fmaf_rn(al, pv.y, xv.y); xv.z = __fmaf_rn(al, pv.z, xv.z); xv.w = __fmaf_rn(al, pv.w, xv.w);
                rv.x = __fmaf_rn(-al, av.x, rv.x); rv.y = __fmaf_rn(-al, av.y, rv.y); rv.z = __fmaf_rn(-al, av.z, rv.z); rv.w = __fmaf_rn(-al, av.w, rv.w);
                *reinterpret_cast<float4*>(P.x + i) = xv;
                *reinterpret_cast<float4*>(P.r + i) = rv;
                part += (double)(rv.x * rv.x);
                part += (double)(rv.y * rv.y);
                part += (double)(rv.z * rv.z);
                part += (double)(rv.w * rv.w);
            }
            warp_add(part, sAcc, d, lane);
        }
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dRRn[tid], sAcc[tid]);
        cg_sync<MG>(grid, P, epoch, cur, 1, dRRn);
        if (tid >= 1 && tid <= D && sActive[tid]) {
            float r0 = sR1[tid], r1 = (float)cg_total<MG>(P, cur, 1, tid, dRRn);
            sR1[tid] = r1;
            int k = sIter[tid] + 1;
            sIter[tid] = k;
            sBeta[tid] = r1 / r0;
            sActive[tid] = (r1 > P.tol2 && k <= P.maxIter) ? 1 : 0;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid >= 1 && tid <= D) { P.itersOut[tid] = sIter[tid] - 1; P.resOut[tid] = sR1[tid]; }
    if (blockIdx.x == 0 && tid == 0) P.itersOut[15] = (int)epoch;
}

int stage_solve(Context& c) {
    const int D = c.D, M = c.M;
    cudaStream_t st = c.stream;
    // vectors are padded by 7 floats so that node 1 (the first sibling block) is 32-byte aligned
    const size_t padN = (size_t)M + 8;
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
            c.mgP = c.mg.alloc<float>(padN, &c.mgPOff);
            if (!c.mgX || !c.mgP) { set_error("multi-GPU arena too small for the CG vectors (prb_mg_init arena_bytes)"); return PRB_ERR_NOMEM; }
        }
        c.xv = c.mgX + 7;
        pBuf = c.mgP;
    } else {
        PRB_TRY(c.x.alloc(padN, st));
        c.xv = c.x.p + 7;
        PRB_TRY(p.alloc(padN, st));
        pBuf = p.p;
    }
    DBuf<double> dots;
    DBuf<int> itersOut;
    PRB_TRY(r.alloc(padN, st));
    PRB_TRY(Ap.alloc(padN, st));
    PRB_TRY(dots.alloc(96, st));
    PRB_TRY(itersOut.alloc(16, st));
    PRB_TRY(resOut.alloc(16, st));
    PRB_CUDA(cudaMemsetAsync(dots.p, 0, 96 * sizeof(double), st));
    PRB_CUDA(cudaMemsetAsync(itersOut.p, 0, 16 * sizeof(int), st));
    CgParams P;
    P.D = D;
    for (int d = 1; d <= D + 1; d++) P.gbase[d] = (c.base[d] - 1) / 8;
    P.gbase[0] = 0;
    // super-groups: sg 0 = depth 1; depth d >= 2 owns the super-groups 1 + (sibling groups of depth d-1)
    P.sgStart[0] = 0;
    P.sgStart[1] = 0;
    for (int d = 2; d <= D + 1; d++) P.sgStart[d] = 1 + P.gbase[d - 1];
    P.sgTab = c.sgTab.p; P.stencil = c.dStencil.p; P.b = c.divg.p;
    P.x = c.xv; P.r = r.p + 7; P.p = pBuf + 7; P.Ap = Ap.p + 7;
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
    P.epoch0 = c.mg.epoch;
    if (mg) PRB_TRY(mg_barrier(c));        // every rank's previous use of the arena buffers is over before anyone writes p / x again
    P.epoch0 = c.mg.epoch;
    P.dots = dots.p; P.itersOut = itersOut.p; P.resOut = resOut.p;
    float tol = (float)c.cgTol;
    P.tol2 = tol * tol;
    P.maxIter = c.cgMaxIter;
    const size_t dynSmem = (size_t)kCgWarps * 2 * kCube * sizeof(float);
    const void* kern = mg ? (const void*)k_cg_all_depths<true> : (const void*)k_cg_all_depths<false>;
    PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem));
    PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int perSM = 0;
    PRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kCgBlock, dynSmem));
    if (perSM < 1) { set_error("CG kernel does not fit on an SM"); return PRB_ERR_CUDA; }
    int gridSize = c.smCount * perSM;
    i64 maxTiles = P.sgStart[D + 1];
    i64 needBlocks = (maxTiles + kCgWarps - 1) / kCgWarps;
    if (gridSize > needBlocks) gridSize = (int)(((needBlocks + c.smCount - 1) / c.smCount) * c.smCount);   // small problems: fewer CTAs, cheaper grid syncs
    if (gridSize > c.smCount * perSM) gridSize = c.smCount * perSM;
    if (gridSize < 1) gridSize = 1;
    void* args[] = {(void*)&P};
    PRB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(gridSize), dim3(kCgBlock), args, dynSmem, st));
    c.launches++;
    int hIters[16];
    PRB_CUDA(cudaMemcpyAsync(hIters, itersOut.p, sizeof(hIters), cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    c.cgRowIters = 0;
    for (int d = 0; d <= D; d++) { c.cgIters[d] = hIters[d]; c.cgRowIters += (i64)(P.row1[d] - P.row0[d]) * hIters[d]; }
    if (mg) {
        c.mg.epoch = (unsigned)hIters[15];
        // collect the other ranks' parts of the solution (pull over NVLink), then let nobody run ahead
        PRB_TRY(mg_barrier(c));
        for (int q = 0; q < c.mg.world; q++) {
            if (q == c.mg.rank) continue;
            const float* px = (const float*)(c.mg.peer[q] + c.mgXOff) + 7;
            for (int d = c.shardFrom; d <= D; d++) {
                size_t n = (size_t)(c.rowLo[d][q + 1] - c.rowLo[d][q]);
                if (n) PRB_CUDA(cudaMemcpyAsync(c.xv + c.rowLo[d][q], px + c.rowLo[d][q], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
        }
        PRB_TRY(mg_barrier(c));
        int err = 0;
        PRB_CUDA(cudaMemcpyAsync(&err, &((MgHeader*)c.mg.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        if (err) { set_error("multi-GPU solve: timed out waiting for a peer"); return PRB_ERR_CUDA; }
    }
    r.release(); p.release(); Ap.release(); dots.release(); itersOut.release(); resOut.release();
    return PRB_OK;
}

}  // namespace prb
