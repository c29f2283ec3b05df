// Per-depth conjugate gradient, all depths in ONE persistent cooperative kernel.
//
// Replaces GenerateSingleNodeLaplacian + scan + 2x copy_if CSR assembly (main.cu:1143-1329) and
// the cooperative CG sample kernel (CG_CUDA.cuh:186-324, 344-509).  The reference solves the
// D+1 independent same-depth systems one after another, each with its own CSR matrix in managed
// memory, 7 grid syncs per iteration and host loops over managed arrays around the launch.
// Here:
//   * the matrix is never formed: rows are the translation-invariant 27-point stencil
//     (4 distinct values per depth) applied through the sibling-block table nbBase[group][27]
//     (13.5 B/row instead of 216 B/row of CSR);
//   * the depths are independent, so they all iterate in lock-step inside one
//     launch: one iteration of the kernel = one CG iteration of every still-active depth, with
//     per-depth alpha / beta / residual and per-depth stopping.  Grid syncs per solve drop from
//     7 * sum_d iters_d to 3 * max_d iters_d;
//   * SpMV is warp-centric: a warp owns 4 sibling groups (32 rows), stages their 27 neighbour
//     blocks with 128-bit loads (a sibling block of 8 floats is one aligned 32-byte sector) into
//     a 6x6x6 shared-memory cube per group, and every row then reads its 3x3x3 window at
//     compile-time offsets from one base address (bank-conflict free) -- no block barriers;
//   * the row sum runs over the neighbour slots in order j = 0..26 with FMAs, exactly the
//     reference's CSR order (absent neighbours contribute an exact +0), so A*p is bit-identical;
//     dots are float products accumulated in double (CG_CUDA.cuh:217-220); alpha, beta are float.
// Algorithmic bytes per row per iteration: 57.5 B
//   (p=r+beta*p: 12, SpMV: 13.5 + 4 + 4, x/r update: 24).
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace prb {

constexpr int kCgBlock = 256;
constexpr int kCgWarps = kCgBlock / 32;
constexpr int kGroupsPerWarp = 4;                 // 32 rows per warp tile

struct CgParams {
    int D;
    int gbase[kMaxDepth + 2];     // first group of depth d (groups cover nodes 1..M-1), gbase[D+1] = total
    const int* nbBase;
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
};

// (blk, half) -> offset inside the 6x6x6 cube of the first of the 4 floats of that half block
__constant__ unsigned short cCubeOff[54];

__global__ void __launch_bounds__(kCgBlock) k_cg_all_depths(CgParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) float sCube[kCgWarps][kGroupsPerWarp][216];
    __shared__ float sSt[kMaxDepth + 1][4];
    __shared__ double sAcc[kMaxDepth + 1];
    __shared__ float sR1[kMaxDepth + 1], sAlpha[kMaxDepth + 1], sBeta[kMaxDepth + 1];
    __shared__ int sActive[kMaxDepth + 1], sIter[kMaxDepth + 1];
    __shared__ int sTileStart[kMaxDepth + 2];     // prefix of warp tiles over active depths
    __shared__ int sGrpStart[kMaxDepth + 2];      // prefix of sibling groups over active depths
    const int D = P.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = blockIdx.x * kCgWarps + warp, nwarps = gridDim.x * kCgWarps;
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
        double acc = 0.0;
        int curD = -1;
        const int totalRows = 8 * P.gbase[D + 1];
        for (int rowi = blockIdx.x * kCgBlock + tid; rowi < totalRows; rowi += gridDim.x * kCgBlock) {
            int i = 1 + rowi, G = rowi >> 3, d = 1;
            while (G >= P.gbase[d + 1]) d++;
            if (d != curD) { if (curD >= 0) atomicAdd(&sAcc[curD], acc); acc = 0.0; curD = d; }
            float bv = P.b[i];
            P.x[i] = 0.f; P.r[i] = bv; P.p[i] = 0.f;
            acc += (double)(bv * bv);
        }
        if (curD >= 0) atomicAdd(&sAcc[curD], acc);
        __syncthreads();
        if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&P.dots[64 + tid], sAcc[tid]);   // dedicated init buffer
    }
    grid.sync();
    if (tid >= 1 && tid <= D) {
        float r1 = (float)P.dots[64 + tid];
        sR1[tid] = r1; sIter[tid] = 1; sBeta[tid] = 0.f; sAlpha[tid] = 0.f;
        sActive[tid] = (r1 > P.tol2 && 1 <= P.maxIter) ? 1 : 0;
    }
    __syncthreads();

    const int gi = lane >> 3, cc = lane & 7;
    const int cubeBase = (2 + ((cc >> 2) & 1)) * 36 + (2 + ((cc >> 1) & 1)) * 6 + (2 + (cc & 1));
    float* myCube = &sCube[warp][0][0];
    // staging task t = lane + 32k handles half-block (blk = t>>1, half = t&1); its cube offset
    // depends on the lane only, so it is computed once (a __constant__ table indexed by lane would
    // serialise: the constant cache serves one address per cycle)
    int cubeOff[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int t = lane + 32 * k, blk = t >> 1, half = t & 1;
        cubeOff[k] = (2 * (blk / 9) + half) * 36 + (2 * ((blk / 3) % 3)) * 6 + 2 * (blk % 3);
    }

    for (int it = 1;; it++) {
        // every block derives the same schedule from the same per-depth state
        if (tid == 0) {
            int acc = 0, accG = 0;
            for (int d = 1; d <= D; d++) {
                sTileStart[d] = acc;
                sGrpStart[d] = accG;
                if (sActive[d]) { acc += (P.gbase[d + 1] - P.gbase[d] + kGroupsPerWarp - 1) / kGroupsPerWarp; accG += P.gbase[d + 1] - P.gbase[d]; }
            }
            sTileStart[D + 1] = acc;
            sGrpStart[D + 1] = accG;
        }
        if (tid <= D) sAcc[tid] = 0.0;
        __syncthreads();
        const int nTiles = sTileStart[D + 1];
        if (nTiles == 0) break;
        const int cur = it & 1, nxt = cur ^ 1;
        double* dPAp = P.dots + cur * 32;         // kind 0
        double* dRRn = P.dots + cur * 32 + 16;    // kind 1 (this iteration's new r.r)

        // ---------------- phase C: p = r + beta p   (beta = 0 and p = 0 in the first iteration)
        // streaming over the active depth slabs, 4 rows (one 128-bit access) per thread
        {
            const int nQuads = 2 * sGrpStart[D + 1];
            int d = 0, lo = 0, hi = 0, gb = 0;
            float be = 0.f;
            for (int q = blockIdx.x * kCgBlock + tid; q < nQuads; q += gridDim.x * kCgBlock) {
                int ag = q >> 1;
                while (ag >= hi) { d++; lo = sGrpStart[d]; hi = sGrpStart[d + 1]; gb = P.gbase[d]; be = sBeta[d]; }   // inactive depths have lo == hi
                int i = 1 + 8 * (gb + ag - lo) + 4 * (q & 1);
                float4 rv = *reinterpret_cast<const float4*>(P.r + i);
                float4 pv = *reinterpret_cast<const float4*>(P.p + i);
                pv.x = __fadd_rn(rv.x, __fmul_rn(be, pv.x));
                pv.y = __fadd_rn(rv.y, __fmul_rn(be, pv.y));
                pv.z = __fadd_rn(rv.z, __fmul_rn(be, pv.z));
                pv.w = __fadd_rn(rv.w, __fmul_rn(be, pv.w));
                *reinterpret_cast<float4*>(P.p + i) = pv;
            }
        }
        grid.sync();
        // ---------------- phase A: Ap = A p ; p.Ap
        {
            double part = 0.0;
            int curD = -1;
            int d = 0, lo = 0, hi = 0, gb = 0, ge = 0;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int tile = gwarp; tile < nTiles; tile += nwarps) {
                while (tile >= hi) {
                    d++; lo = sTileStart[d]; hi = sTileStart[d + 1]; gb = P.gbase[d]; ge = P.gbase[d + 1];
                    s0 = sSt[d][0]; s1 = sSt[d][1]; s2 = sSt[d][2]; s3 = sSt[d][3];
                }
                const int g0 = gb + (tile - lo) * kGroupsPerWarp;
                const int ng = min(kGroupsPerWarp, ge - g0);
                if (d != curD) {
                    if (curD >= 0) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
                        if (lane == 0) atomicAdd(&sAcc[curD], part);
                    }
                    part = 0.0;
                    curD = d;
                }
                __syncwarp();
                {
                    // all 8 base loads first, then all 8 128-bit value loads, then the stores:
                    // 8 independent requests in flight per lane instead of a base->value chain per group
                    const int* nbp = P.nbBase + 27 * (i64)g0;
                    int bb[kGroupsPerWarp][2];
#pragma unroll
                    for (int g = 0; g < kGroupsPerWarp; g++)
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            int t = lane + 32 * k;
                            bb[g][k] = (g < ng && t < 54) ? nbp[27 * g + (t >> 1)] : -1;
                        }
                    float4 vv[kGroupsPerWarp][2];
#pragma unroll
                    for (int g = 0; g < kGroupsPerWarp; g++)
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            vv[g][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (bb[g][k] >= 0) vv[g][k] = *reinterpret_cast<const float4*>(P.p + bb[g][k] + 4 * (lane & 1));
                        }
#pragma unroll
                    for (int g = 0; g < kGroupsPerWarp; g++)
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            int t = lane + 32 * k;
                            if (g < ng && t < 54) {
                                int off = cubeOff[k];
                                float* cube = myCube + g * 216;
                                *reinterpret_cast<float2*>(cube + off) = make_float2(vv[g][k].x, vv[g][k].y);
                                *reinterpret_cast<float2*>(cube + off + 6) = make_float2(vv[g][k].z, vv[g][k].w);
                            }
                        }
                }
                __syncwarp();
                if (gi < ng) {
                    const float* cb = myCube + gi * 216 + cubeBase;
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 27; j++) {
                        const int dx = j / 9 - 1, dy = (j / 3) % 3 - 1, dz = j % 3 - 1;
                        const int ty = (dx != 0) + (dy != 0) + (dz != 0);
                        const float sv = ty == 0 ? s0 : (ty == 1 ? s1 : (ty == 2 ? s2 : s3));
                        acc = __fmaf_rn(sv, cb[dx * 36 + dy * 6 + dz], acc);
                    }
                    int i = 1 + 8 * (g0 + gi) + cc;
                    P.Ap[i] = acc;
                    part += (double)(cb[0] * acc);
                }
            }
            if (curD >= 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
                if (lane == 0) atomicAdd(&sAcc[curD], part);
            }
            __syncthreads();
            if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dPAp[tid], sAcc[tid]);
        }
        grid.sync();
        // both accumulators of the NEXT iteration are zeroed here: every block has passed this
        // iteration's syncs, hence finished reading them after the previous iteration's syncs
        if (blockIdx.x == 0 && tid < 32) P.dots[nxt * 32 + tid] = 0.0;
        if (tid <= D) sAcc[tid] = 0.0;
        if (tid >= 1 && tid <= D && sActive[tid]) sAlpha[tid] = (float)((double)sR1[tid] / dPAp[tid]);
        __syncthreads();
        // ---------------- phase B: x += alpha p ; r -= alpha Ap ; r.r
        {
            double part = 0.0;
            int curD = -1;
            const int nQuads = 2 * sGrpStart[D + 1];
            int d = 0, lo = 0, hi = 0, gb = 0;
            float al = 0.f;
            for (int q = blockIdx.x * kCgBlock + tid; q < nQuads; q += gridDim.x * kCgBlock) {
                int ag = q >> 1;
                while (ag >= hi) { d++; lo = sGrpStart[d]; hi = sGrpStart[d + 1]; gb = P.gbase[d]; al = sAlpha[d]; }
                if (d != curD) { if (curD >= 0) atomicAdd(&sAcc[curD], part); part = 0.0; curD = d; }
                int i = 1 + 8 * (gb + ag - lo) + 4 * (q & 1);
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
            {
                // one shared-memory atomic per warp when the whole warp ended in the same depth
                int d0 = __shfl_sync(0xffffffffu, curD, 0);
                if (__all_sync(0xffffffffu, curD == d0)) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
                    if (lane == 0 && d0 >= 0) atomicAdd(&sAcc[d0], part);
                } else if (curD >= 0) atomicAdd(&sAcc[curD], part);
            }
            __syncthreads();
            if (tid >= 1 && tid <= D && sAcc[tid] != 0.0) atomicAdd(&dRRn[tid], sAcc[tid]);
        }
        grid.sync();
        if (tid >= 1 && tid <= D && sActive[tid]) {
            float r0 = sR1[tid], r1 = (float)dRRn[tid];
            sR1[tid] = r1;
            int k = sIter[tid] + 1;
            sIter[tid] = k;
            sBeta[tid] = r1 / r0;
            sActive[tid] = (r1 > P.tol2 && k <= P.maxIter) ? 1 : 0;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid >= 1 && tid <= D) { P.itersOut[tid] = sIter[tid] - 1; P.resOut[tid] = sR1[tid]; }
}

static bool g_cubeReady = false;

int stage_solve(Context& c) {
    const int D = c.D, M = c.M;
    cudaStream_t st = c.stream;
    if (!g_cubeReady) {
        unsigned short h[54];
        for (int t = 0; t < 54; t++) {
            int blk = t >> 1, half = t & 1;
            int bx = blk / 9, by = (blk / 3) % 3, bz = blk % 3;
            h[t] = (unsigned short)((2 * bx + half) * 36 + (2 * by) * 6 + 2 * bz);
        }
        PRB_CUDA(cudaMemcpyToSymbol(cCubeOff, h, sizeof(h)));
        g_cubeReady = true;
    }
    // vectors are padded by 7 floats so that node 1 (the first sibling block) is 32-byte aligned
    const size_t padN = (size_t)M + 8;
    PRB_TRY(c.x.alloc(padN, st));
    c.xv = c.x.p + 7;
    DBuf<float> r, p, Ap, resOut;
    DBuf<double> dots;
    DBuf<int> itersOut;
    PRB_TRY(r.alloc(padN, st));
    PRB_TRY(p.alloc(padN, st));
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
    P.nbBase = c.nbBase.p; P.stencil = c.dStencil.p; P.b = c.divg.p;
    P.x = c.xv; P.r = r.p + 7; P.p = p.p + 7; P.Ap = Ap.p + 7;
    P.dots = dots.p; P.itersOut = itersOut.p; P.resOut = resOut.p;
    float tol = (float)c.cgTol;
    P.tol2 = tol * tol;
    P.maxIter = c.cgMaxIter;
    int perSM = 0;
    PRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_cg_all_depths, kCgBlock, 0));
    if (perSM < 1) { set_error("CG kernel does not fit on an SM"); return PRB_ERR_CUDA; }
    int gridSize = c.smCount * perSM;
    i64 maxTiles = 0;
    for (int d = 1; d <= D; d++) maxTiles += (P.gbase[d + 1] - P.gbase[d] + kGroupsPerWarp - 1) / kGroupsPerWarp;
    i64 needBlocks = (maxTiles + kCgWarps - 1) / kCgWarps;
    if (gridSize > needBlocks) gridSize = (int)(((needBlocks + c.smCount - 1) / c.smCount) * c.smCount);   // small problems: fewer CTAs, cheaper grid syncs
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
    r.release(); p.release(); Ap.release(); dots.release(); itersOut.release(); resOut.release();
    return PRB_OK;
}

}  // namespace prb
