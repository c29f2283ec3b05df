// Vector-field splat (A7) and divergence (A8).
//
// Splat replaces computeVectorField / F_center_width_Point (main.cu:913-964): every depth-D
// slot gathers B-hat((c_o - q)/w) * n_q over the samples of its 27 neighbours, in neighbour
// order then sample order, with the reference's float evaluation of the shifted piecewise
// polynomial (ConfirmedPPolynomial.cuh:35-48, 79-91) so the weights carry the same rounding.
//
// Divergence replaces precomputeEncodedFunctionIdxOfNode + computeEncodedFinerNodesDivergence
// + the per-node host loop for depths 0-4 (main.cu:985-1141, 3383-3462).  The res x res double
// table dot_F_DF is replaced by the translation-invariant rows dfT[d][t] (bspline_host.h) and
// the per-node host loop by a block-per-node reduction; nodes of the two finest depths are
// summed by one thread in the reference's order (bit-identical), coarser nodes by a warp or a
// block with double partial sums.
#include "common.cuh"

namespace prb {

// value at xe of B-hat(width 2^-D) shifted to t; fn = 4 pieces x (c0,c1,c2,start)
__device__ __forceinline__ float shifted_bspline(const float* __restrict__ fn, float t, float xe) {
    float res = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float start = __fadd_rn(fn[4 * i + 3], t);
        if (!(xe > start)) break;
        float a[3] = {fn[4 * i], fn[4 * i + 1], fn[4 * i + 2]};
        float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k <= 2; k++) {
            float temp = 1.f;
#pragma unroll
            for (int j = k; j >= 0; j--) {
                c[j] = __fmaf_rn(a[k], temp, c[j]);
                temp = __fmul_rn(temp, __fmul_rn(-t, (float)j));
                temp = __fdiv_rn(temp, (float)(k - j + 1));
            }
        }
        float pw = 1.f, v = 0.f;
#pragma unroll
        for (int j = 0; j <= 2; j++) {
            v = __fmaf_rn(pw, c[j], v);
            pw = __fmul_rn(pw, xe);
        }
        res = __fadd_rn(res, v);
    }
    return res;
}

__global__ void __launch_bounds__(128) k_splat(const float* __restrict__ fn, const float* __restrict__ P, const float* __restrict__ Nr,
                                               const int* __restrict__ neighs, const int* __restrict__ pidx, const int* __restrict__ pnum,
                                               const ushort4* __restrict__ offs, int baseD, int countD, float width, float* __restrict__ V) {
    __shared__ float sfn[16];
    if (threadIdx.x < 16) sfn[threadIdx.x] = fn[threadIdx.x];
    __syncthreads();
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < countD; l += gridDim.x * blockDim.x) {
        int i = baseD + l;
        ushort4 o = offs[i];
        float oc[3] = {(float)((0.5 + o.x) * (double)width), (float)((0.5 + o.y) * (double)width), (float)((0.5 + o.z) * (double)width)};
        float val[3] = {0.f, 0.f, 0.f};
        const int* nb = neighs + 27 * (i64)i;
        for (int j = 0; j < 27; j++) {
            int n = nb[j];
            if (n < 0) continue;
            int p0 = pidx[n], pn = pnum[n];
            for (int k = 0; k < pn; k++) {
                i64 q = 3 * (i64)(p0 + k);
                float wx = shifted_bspline(sfn, P[q], oc[0]);
                float wy = shifted_bspline(sfn, P[q + 1], oc[1]);
                float wz = shifted_bspline(sfn, P[q + 2], oc[2]);
                float w = __fmul_rn(__fmul_rn(wx, wy), wz);
                val[0] = __fmaf_rn(w, Nr[q], val[0]);
                val[1] = __fmaf_rn(w, Nr[q + 1], val[1]);
                val[2] = __fmaf_rn(w, Nr[q + 2], val[2]);
            }
        }
        V[3 * (i64)l] = val[0];
        V[3 * (i64)l + 1] = val[1];
        V[3 * (i64)l + 2] = val[2];
    }
}

// G cooperating threads per node: 1 (two finest depths), 32 (warp) or blockDim (coarse depths)
template <int G>
__global__ void __launch_bounds__(256) k_divergence(const float* __restrict__ V, const ushort4* __restrict__ offs, const int* __restrict__ neighs,
                                                    const int* __restrict__ didx, const int* __restrict__ dnum, const float* __restrict__ dfRow,
                                                    int base, int count, int baseD, int k /* 2^(D-d) */, float* __restrict__ divg) {
    const int groupsPerBlock = (G == 1) ? blockDim.x : (G == 32 ? blockDim.x / 32 : 1);
    const int gid = (G == 1) ? threadIdx.x : (G == 32 ? threadIdx.x >> 5 : 0);
    const int lane = (G == 1) ? 0 : (G == 32 ? (threadIdx.x & 31) : threadIdx.x);
    const int gsz = (G == 1) ? 1 : (G == 32 ? 32 : blockDim.x);
    __shared__ double red[32];
    for (int l0 = blockIdx.x * groupsPerBlock; l0 < count; l0 += gridDim.x * groupsPerBlock) {
        int l = l0 + gid;
        double val = 0.0;
        if (l < count) {
            int i = base + l;
            ushort4 o = offs[i];
            int bx = k * ((int)o.x - 1), by = k * ((int)o.y - 1), bz = k * ((int)o.z - 1);
            const int* nb = neighs + 27 * (i64)i;
            for (int j = 0; j < 27; j++) {
                int n = nb[j];
                if (n < 0) continue;
                int s0 = didx[n], sn = dnum[n];
                for (int q = lane; q < sn; q += gsz) {
                    int s = s0 + q;
                    ushort4 so = offs[baseD + s];
                    float u0 = dfRow[(int)so.x - bx], u1 = dfRow[(int)so.y - by], u2 = dfRow[(int)so.z - bz];
                    float dp = __fmul_rn(V[3 * (i64)s], u0);                  // DotProduct (main.cu:966-972)
                    dp = __fmaf_rn(V[3 * (i64)s + 1], u1, dp);
                    dp = __fmaf_rn(V[3 * (i64)s + 2], u2, dp);
                    val += (double)dp;
                }
            }
        }
        if (G == 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
        } else if (G > 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
            __syncthreads();
            if (threadIdx.x < 32) {
                val = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
            }
        }
        if (lane == 0 && l < count) divg[base + l] = (float)val;
    }
}

// ---- coarse depths: scatter form.  A coarse node has up to 27 * 8^(D-d) terms, far too many for
// one thread block (the reference runs depths 0-4 as a host loop with one launch set per node,
// main.cu:3419-3458).  Work item = (node n, chunk of <= kChunk of the depth-D slots under n);
// every slot contributes to the <= 27 neighbours o of n, so a block accumulates 27 partial sums
// in double, reduces them and issues at most 27 atomicAdd(double).
constexpr int kChunk = 2048;
__global__ void __launch_bounds__(256) k_count_items(const int* __restrict__ dnum, int nNodes, int* __restrict__ items) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += gridDim.x * blockDim.x) items[i] = (dnum[i] + kChunk - 1) / kChunk;
}
__global__ void __launch_bounds__(256) k_divergence_scatter(const float* __restrict__ V, const ushort4* __restrict__ offs, const int* __restrict__ neighs,
                                                            const int* __restrict__ didx, const int* __restrict__ dnum, const float* __restrict__ dfT,
                                                            const int* __restrict__ dfOffset, const int* __restrict__ itemBase, int nNodes, int baseD, int D,
                                                            double* __restrict__ accum) {
    __shared__ int sNode;
    __shared__ int sNb[27];
    __shared__ double sRed[8][27];
    if (threadIdx.x == 0) {
        int lo = 0, hi = nNodes;                      // last node with itemBase <= blockIdx.x
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (itemBase[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
        sNode = lo;
    }
    __syncthreads();
    const int n = sNode;
    if (threadIdx.x < 27) sNb[threadIdx.x] = neighs[27 * (i64)n + threadIdx.x];
    const ushort4 on = offs[n];
    const int d = on.w, k = 1 << (D - d);
    const float* row = dfT + dfOffset[d];
    const int chunk = (int)blockIdx.x - itemBase[n];
    const int s0 = didx[n] + chunk * kChunk;
    const int cnt = min(kChunk, dnum[n] - chunk * kChunk);
    double acc[27];
#pragma unroll
    for (int j = 0; j < 27; j++) acc[j] = 0.0;
    const int bx = k * ((int)on.x - 1), by = k * ((int)on.y - 1), bz = k * ((int)on.z - 1);
    for (int q = threadIdx.x; q < cnt; q += 256) {
        int s = s0 + q;
        ushort4 so = offs[baseD + s];
        float v0 = V[3 * (i64)s], v1 = V[3 * (i64)s + 1], v2 = V[3 * (i64)s + 2];
        float ux[3], uy[3], uz[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {                 // neighbour o = n + (t-1) per axis
            ux[t] = row[(int)so.x - bx - k * (t - 1)];
            uy[t] = row[(int)so.y - by - k * (t - 1)];
            uz[t] = row[(int)so.z - bz - k * (t - 1)];
        }
#pragma unroll
        for (int j = 0; j < 27; j++) {
            float dp = __fmul_rn(v0, ux[j / 9]);
            dp = __fmaf_rn(v1, uy[(j / 3) % 3], dp);
            dp = __fmaf_rn(v2, uz[j % 3], dp);
            acc[j] += (double)dp;
        }
    }
#pragma unroll
    for (int j = 0; j < 27; j++) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 27 && sNb[threadIdx.x] >= 0) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += sRed[w][threadIdx.x];
        atomicAdd(&accum[sNb[threadIdx.x]], v);
    }
}
__global__ void __launch_bounds__(256) k_divergence_finish(const double* __restrict__ accum, int n, float* __restrict__ divg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) divg[i] = (float)accum[i];
}

int stage_splat(Context& c) {
    const int D = c.D;
    cudaStream_t st = c.stream;
    PRB_TRY(c.V.alloc(3 * (size_t)c.cnt[D], st));
    PRB_TRY(c.divg.alloc((size_t)c.M, st));
    float width = (float)(1.0 / (1 << D));
    PRB_LAUNCH(c, k_splat, grid_for(c, c.cnt[D], 128, 16), 128, 0, c.dMaxDepthFn.p, c.P.p, c.Nr.p, c.neighs.p, c.pidx.p, c.pnum.p, c.offs.p,
               c.base[D], c.cnt[D], width, c.V.p);
    PRB_CUDA(cudaEventRecord(c.ev[3], st));
    PRB_TRY(stage_divergence(c));
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

int stage_divergence(Context& c) {
    const int D = c.D;
    cudaStream_t st = c.stream;
    // depths with up to 8^4 slots under a node and deeper trees of slots go through the scatter
    // kernel; the three finest depths are gathered per node (thread / warp), which for the two
    // finest reproduces the reference's summation order exactly.
    int dc = D - 4;                                   // last depth handled by the scatter kernel
    if (dc >= 0) {
        const int nCoarse = c.base[dc + 1];
        DBuf<int> items, itemBase;
        DBuf<double> accum;
        PRB_TRY(items.alloc((size_t)nCoarse, st));
        PRB_TRY(itemBase.alloc((size_t)nCoarse, st));
        PRB_TRY(accum.alloc((size_t)nCoarse, st));
        PRB_CUDA(cudaMemsetAsync(accum.p, 0, sizeof(double) * (size_t)nCoarse, st));
        PRB_LAUNCH(c, k_count_items, grid_for(c, nCoarse, 256), 256, 0, c.dnum.p, nCoarse, items.p);
        i64 nItems = 0;
        PRB_TRY(exclusive_scan(c, items.p, itemBase.p, nCoarse, &nItems));
        if (nItems > 0)
            PRB_LAUNCH(c, k_divergence_scatter, (int)nItems, 256, 0, c.V.p, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, c.dDfT.p, c.dDfOffset.p, itemBase.p,
                       nCoarse, c.base[D], D, accum.p);
        PRB_LAUNCH(c, k_divergence_finish, grid_for(c, nCoarse, 256), 256, 0, accum.p, nCoarse, c.divg.p);
        items.release(); itemBase.release(); accum.release();
    }
    for (int d = (dc >= 0 ? dc + 1 : 0); d <= D; d++) {
        int k = 1 << (D - d);
        const float* row = c.dDfT.p + c.tab.dfOffset[d];
        int n = c.cnt[d];
        if (d >= D - 1)
            PRB_LAUNCH(c, k_divergence<1>, grid_for(c, n, 256), 256, 0, c.V.p, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, row, c.base[d], n, c.base[D], k, c.divg.p);
        else
            PRB_LAUNCH(c, k_divergence<32>, grid_for(c, (i64)n * 32, 256), 256, 0, c.V.p, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, row, c.base[d], n, c.base[D], k, c.divg.p);
    }
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

}  // namespace prb
