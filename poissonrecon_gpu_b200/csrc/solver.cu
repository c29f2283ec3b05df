// Per-depth conjugate gradient, all depths in ONE persistent cooperative kernel.
//
// Replaces GenerateSingleNodeLaplacian + scan + 2x copy_if CSR assembly (main.cu:1143-1329) and
// the cooperative CG sample kernel (CG_CUDA.cuh:186-324, 344-509).  The reference solves the
// D+1 independent same-depth systems one after another, each with its own CSR matrix in managed
// memory, 7 grid syncs per iteration and host loops over managed arrays around the launch.
// Here:
//   * the matrix is never formed: rows are the translation-invariant 27-point stencil
//     (4 distinct values per depth, stencil[d][27]) applied through the sibling-block table
//     nbBase[group][27] (13.5 B/row instead of 216 B/row of CSR);
//   * the depths are independent, so they all iterate in lock-step inside one
//     launch: one iteration of the kernel = one CG iteration of every still-active depth, with
//     per-depth alpha / beta / residual and per-depth stopping.  Grid syncs per solve drop from
//     7 * sum_d iters_d to 2 * max_d iters_d;
//   * p = r + beta*p is recomputed on the fly while the neighbour blocks are staged into shared
//     memory (double-buffered p), which removes the third pass and its sync;
//   * the row sum runs over present neighbours in slot order j = 0..26 with FMAs, exactly the
//     reference's CSR order, so A*p is bit-identical; dots are float products accumulated in
//     double (CG_CUDA.cuh:217-220); alpha, beta are float.
// Algorithmic bytes per row per iteration: 57.5 B.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace prb {

constexpr int kCgBlock = 256;
constexpr int kGroupsPerTile = kCgBlock / 8;      // 32 sibling groups = 256 rows per tile

struct CgParams {
    int D;
    int gbase[kMaxDepth + 2];     // first group of depth d (groups cover nodes 1..M-1), gbase[D+1] = total
    const int* nbBase;
    const float* stencil;          // [D+1][27]
    const float* b;                // divergence
    float* x;
    float* r;
    float* p0;
    float* p1;
    float* Ap;
    double* dots;                  // [2 buffers][2 kinds][16] + [16] for the initial r.r
    int* itersOut;                 // [D+1]
    float* resOut;                 // [D+1] final r.r
    float tol2;
    int maxIter;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (kCgBlock >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;   // valid in thread 0
}

__global__ void __launch_bounds__(kCgBlock) k_cg_all_depths(CgParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float sval[kGroupsPerTile][27][8];
    __shared__ int sbase[kGroupsPerTile][27];
    __shared__ float sSt[kMaxDepth + 1][27];
    __shared__ double red[32];
    __shared__ float sR1[kMaxDepth + 1], sR0[kMaxDepth + 1], sAlpha[kMaxDepth + 1], sBeta[kMaxDepth + 1];
    __shared__ int sActive[kMaxDepth + 1], sIter[kMaxDepth + 1];
    __shared__ int sTileStart[kMaxDepth + 2];     // prefix of tiles over active depths
    const int D = P.D, tid = threadIdx.x;
    for (int t = tid; t < (D + 1) * 27; t += kCgBlock) sSt[t / 27][t % 27] = P.stencil[t];
    __syncthreads();

    // ---- depth 0: a 1x1 system, solved by one thread with the same recurrences
    if (blockIdx.x == 0 && tid == 0) {
        float a00 = sSt[0][13];
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
        double acc = 0.0;
        int curD = -1;
        __shared__ double sAcc[kMaxDepth + 1];
        if (tid <= D) sAcc[tid] = 0.0;
        __syncthreads();
        int totalRows = 8 * P.gbase[D + 1];
        for (int rowi = blockIdx.x * kCgBlock + tid; rowi < totalRows; rowi += gridDim.x * kCgBlock) {
            int i = 1 + rowi, G = rowi >> 3, d = 1;
            while (G >= P.gbase[d + 1]) d++;
            if (d != curD) { if (curD >= 0) atomicAdd(&sAcc[curD], acc); acc = 0.0; curD = d; }
            float bv = P.b[i];
            P.x[i] = 0.f; P.r[i] = bv; P.p0[i] = 0.f; P.p1[i] = 0.f;
            acc += (double)(bv * bv);
        }
        if (curD >= 0) atomicAdd(&sAcc[curD], acc);
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&P.dots[64 + tid], sAcc[tid]);   // dedicated init buffer
    }
    grid.sync();
    if (tid >= 1 && tid <= D) {
        float r1 = (float)P.dots[64 + tid];
        sR1[tid] = r1; sR0[tid] = 0.f; sIter[tid] = 1; sBeta[tid] = 0.f; sAlpha[tid] = 0.f;
        sActive[tid] = (r1 > P.tol2 && 1 <= P.maxIter) ? 1 : 0;
    }
    __syncthreads();

    for (int it = 1;; it++) {
        // every block derives the same schedule from the same per-depth state
        int anyActive = 0;
        if (tid == 0) {
            int acc = 0;
            for (int d = 1; d <= D; d++) {
                sTileStart[d] = acc;
                if (sActive[d]) acc += (P.gbase[d + 1] - P.gbase[d] + kGroupsPerTile - 1) / kGroupsPerTile;
            }
            sTileStart[D + 1] = acc;
        }
        __syncthreads();
        const int nTiles = sTileStart[D + 1];
        anyActive = nTiles > 0;
        if (!anyActive) break;
        const int cur = it & 1, nxt = cur ^ 1;
        double* dPAp = P.dots + cur * 32;         // kind 0
        double* dRRn = P.dots + cur * 32 + 16;    // kind 1 (this iteration's new r.r)
        const float* pOld = (it & 1) ? P.p0 : P.p1;
        float* pNew = (it & 1) ? P.p1 : P.p0;
        // ---------------- phase A: p = r + beta p ; Ap = A p ; p.Ap
        {
            __shared__ double sAcc[kMaxDepth + 1];
            if (tid <= D) sAcc[tid] = 0.0;
            __syncthreads();
            for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
                int d = 1;
                while (!(sActive[d] && tile >= sTileStart[d] && tile < sTileStart[d + 1])) d++;
                const int g0 = P.gbase[d] + (tile - sTileStart[d]) * kGroupsPerTile;
                const int ng = min(kGroupsPerTile, P.gbase[d + 1] - g0);
                const float beta = sBeta[d];
                for (int t = tid; t < ng * 27; t += kCgBlock) (&sbase[0][0])[t] = P.nbBase[27 * (i64)g0 + t];
                __syncthreads();
                // stage the 27 neighbour blocks of every group: thread (g, e) walks the 27 blocks, so
                // 8 lanes read one 32-byte sector per load and all 27 (x2) loads are independent
                {
                    const int g = tid >> 3, e = tid & 7;
                    if (g < ng) {
#pragma unroll
                        for (int blk = 0; blk < 27; blk++) {
                            int b = sbase[g][blk];
                            float val = 0.f;
                            if (b >= 0) val = __fadd_rn(P.r[b + e], __fmul_rn(beta, pOld[b + e]));
                            sval[g][blk][e] = val;
                        }
                    }
                }
                __syncthreads();
                double part = 0.0;
                int g = tid >> 3, c = tid & 7;
                if (g < ng) {
                    const int cx = (c >> 2) & 1, cy = (c >> 1) & 1, cz = c & 1;
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 27; j++) {
                        const int tx = cx + j / 9 - 1, ty = cy + (j / 3) % 3 - 1, tz = cz + j % 3 - 1;
                        const int blk = (tx < 0 ? 0 : (tx > 1 ? 2 : 1)) * 9 + (ty < 0 ? 0 : (ty > 1 ? 2 : 1)) * 3 + (tz < 0 ? 0 : (tz > 1 ? 2 : 1));
                        const int e = ((tx & 1) << 2) | ((ty & 1) << 1) | (tz & 1);
                        if (sbase[g][blk] >= 0) acc = __fmaf_rn(sSt[d][j], sval[g][blk][e], acc);
                    }
                    int i = 1 + 8 * (g0 + g) + c;
                    float pv = sval[g][13][c];
                    P.Ap[i] = acc;
                    pNew[i] = pv;
                    part = (double)(pv * acc);
                }
                double tot = block_sum(part, red);
                if (tid == 0) sAcc[d] += tot;
                __syncthreads();
            }
            __syncthreads();
            if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dPAp[tid], sAcc[tid]);
        }
        grid.sync();
        // both accumulators of the NEXT iteration are zeroed here: every block has passed this
        // iteration's first sync, hence finished reading them after the previous iteration's syncs
        if (blockIdx.x == 0 && tid < 32) P.dots[nxt * 32 + tid] = 0.0;
        if (tid >= 1 && tid <= D && sActive[tid]) sAlpha[tid] = (float)((double)sR1[tid] / dPAp[tid]);
        __syncthreads();
        // ---------------- phase B: x += alpha p ; r -= alpha Ap ; r.r
        {
            __shared__ double sAcc[kMaxDepth + 1];
            if (tid <= D) sAcc[tid] = 0.0;
            __syncthreads();
            for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
                int d = 1;
                while (!(sActive[d] && tile >= sTileStart[d] && tile < sTileStart[d + 1])) d++;
                const int g0 = P.gbase[d] + (tile - sTileStart[d]) * kGroupsPerTile;
                const int ng = min(kGroupsPerTile, P.gbase[d + 1] - g0);
                const float al = sAlpha[d];
                double part = 0.0;
                if (tid < ng * 8) {
                    int i = 1 + 8 * g0 + tid;
                    float pv = pNew[i], av = P.Ap[i];
                    P.x[i] = __fmaf_rn(al, pv, P.x[i]);
                    float rv = __fmaf_rn(-al, av, P.r[i]);
                    P.r[i] = rv;
                    part = (double)(rv * rv);
                }
                double tot = block_sum(part, red);
                if (tid == 0) sAcc[d] += tot;
                __syncthreads();
            }
            __syncthreads();
            if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dRRn[tid], sAcc[tid]);
        }
        grid.sync();
        if (tid >= 1 && tid <= D && sActive[tid]) {
            float r0 = sR1[tid], r1 = (float)dRRn[tid];
            sR0[tid] = r0; sR1[tid] = r1;
            int k = sIter[tid] + 1;
            sIter[tid] = k;
            sBeta[tid] = r1 / r0;
            sActive[tid] = (r1 > P.tol2 && k <= P.maxIter) ? 1 : 0;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid >= 1 && tid <= D) { P.itersOut[tid] = sIter[tid] - 1; P.resOut[tid] = sR1[tid]; }
}

int stage_solve(Context& c) {
    const int D = c.D, M = c.M;
    cudaStream_t st = c.stream;
    PRB_TRY(c.x.alloc((size_t)M, st));
    DBuf<float> r, p0, p1, Ap, resOut;
    DBuf<double> dots;
    DBuf<int> itersOut;
    PRB_TRY(r.alloc((size_t)M, st));
    PRB_TRY(p0.alloc((size_t)M, st));
    PRB_TRY(p1.alloc((size_t)M, st));
    PRB_TRY(Ap.alloc((size_t)M, st));
    PRB_TRY(dots.alloc(96, st));
    PRB_TRY(itersOut.alloc(16, st));
    PRB_TRY(resOut.alloc(16, st));
    PRB_CUDA(cudaMemsetAsync(dots.p, 0, 96 * sizeof(double), st));
    PRB_CUDA(cudaMemsetAsync(itersOut.p, 0, 16 * sizeof(int), st));
    CgParams P;
    P.D = D;
    for (int d = 1; d <= D + 1; d++) P.gbase[d] = (c.base[d] - 1) / 8;
    P.gbase[0] = 0;
    P.nbBase = c.nbBase.p; P.stencil = c.dStencil.p; P.b = c.divg.p; P.x = c.x.p; P.r = r.p; P.p0 = p0.p; P.p1 = p1.p; P.Ap = Ap.p;
    P.dots = dots.p; P.itersOut = itersOut.p; P.resOut = resOut.p;
    float tol = (float)c.cgTol;
    P.tol2 = tol * tol;
    P.maxIter = c.cgMaxIter;
    int perSM = 0;
    PRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_cg_all_depths, kCgBlock, 0));
    if (perSM < 1) { set_error("CG kernel does not fit on an SM"); return PRB_ERR_CUDA; }
    int gridSize = c.smCount * perSM;
    int maxTiles = 0;
    for (int d = 1; d <= D; d++) maxTiles += (P.gbase[d + 1] - P.gbase[d] + kGroupsPerTile - 1) / kGroupsPerTile;
    if (gridSize > maxTiles) gridSize = ((maxTiles + c.smCount - 1) / c.smCount) * c.smCount;   // small problems: fewer CTAs, cheaper grid syncs
    if (gridSize > c.smCount * perSM) gridSize = c.smCount * perSM;
    if (gridSize < 1) gridSize = 1;
    void* args[] = {(void*)&P};
    PRB_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_all_depths, dim3(gridSize), dim3(kCgBlock), args, 0, st));
    c.launches++;
    int hIters[16];
    PRB_CUDA(cudaMemcpyAsync(hIters, itersOut.p, sizeof(hIters), cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    c.cgRowIters = 0;
    for (int d = 0; d <= D; d++) { c.cgIters[d] = hIters[d]; c.cgRowIters += (i64)c.cnt[d] * hIters[d]; }
    r.release(); p0.release(); p1.release(); Ap.release(); dots.release(); itersOut.release(); resOut.release();
    return PRB_OK;
}

}  // namespace prb
