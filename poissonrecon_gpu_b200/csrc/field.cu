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

// per-sample weights: W[q][axis][t] = B-hat evaluated for the slot whose offset along `axis` is
// (offset of q's leaf) + t - 1.  A sample is seen by up to 27 slots; the shifted-polynomial
// evaluation (divisions included) is done once per sample and axis instead of once per pair.
__global__ void __launch_bounds__(128) k_splat_weights(const float* __restrict__ fn, const float* __restrict__ P, const int* __restrict__ p2n,
                                                       const ushort4* __restrict__ offsD, i64 N, float width, float* __restrict__ W) {
    __shared__ float sfn[16];
    if (threadIdx.x < 16) sfn[threadIdx.x] = fn[threadIdx.x];
    __syncthreads();
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (i64)gridDim.x * blockDim.x) {
        const ushort4 o = offsD[p2n[q]];
        const int oo[3] = {(int)o.x, (int)o.y, (int)o.z};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float pa = P[3 * q + a];
#pragma unroll
            for (int t = 0; t < 3; t++) {
                float oc = (float)((0.5 + (double)(oo[a] + t - 1)) * (double)width);
                W[9 * q + 3 * a + t] = shifted_bspline(sfn, pa, oc);
            }
        }
    }
}
// One lane per depth-D slot, four sibling groups per warp.  The 27 neighbours of the 8 siblings of a
// group all lie in the 4x4x4 node cube around it, so the sample ranges (pidx, pnum) of those 64
// nodes are fetched once per group into shared memory (192 gathers instead of 8 x 27 x 3); every
// slot then adds its terms in the reference's order: neighbour j = 0..26, samples ascending.
__global__ void __launch_bounds__(128) k_splat(const float* __restrict__ W, const float* __restrict__ Nr,
                                               const int* __restrict__ neighs, const int* __restrict__ pidx, const int* __restrict__ pnum,
                                               int baseD, int countD, float* __restrict__ V) {
    __shared__ int2 sInfo[4][4][64];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int nGroups = countD >> 3, nSteps = (nGroups + 3) >> 2;
    for (int st = blockIdx.x * 4 + wp; st < nSteps; st += gridDim.x * 4) {
        const int g0 = 4 * st;
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
            const int q = rr >> 1, e = lane + 32 * (rr & 1), ux = e >> 4, uy = (e >> 2) & 3, uz = e & 3;
            const int sx = ux >> 1, sy = uy >> 1, sz = uz >> 1;
            const int j = 9 * (ux - sx) + 3 * (uy - sy) + (uz - sz);          // 9(dx+1)+3(dy+1)+(dz+1) with d = u - 1 - s
            int2 info = make_int2(0, 0);
            if (g0 + q < nGroups) {
                const int n = neighs[27 * (i64)(baseD + 8 * (g0 + q) + ((sx << 2) | (sy << 1) | sz)) + j];
                if (n >= 0) info = make_int2(pidx[n], pnum[n]);
            }
            sInfo[wp][q][e] = info;
        }
        __syncwarp();
        const int q = lane >> 3, k = lane & 7;
        if (g0 + q < nGroups) {
            const int l = 8 * (g0 + q) + k;
            const int sx = (k >> 2) & 1, sy = (k >> 1) & 1, sz = k & 1;
            float val[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 27; j++) {
                const int2 info = sInfo[wp][q][(sx + j / 9) * 16 + (sy + (j / 3) % 3) * 4 + (sz + j % 3)];
                // the slot lies at direction -d_j of the sample's leaf: weight index t = 2 - (digit of j)
                const int tx = 2 - j / 9, ty = 2 - (j / 3) % 3, tz = 2 - j % 3;
                for (int t = 0; t < info.y; t++) {
                    const i64 sq = info.x + t;
                    const float* wq = W + 9 * sq;
                    const float w = __fmul_rn(__fmul_rn(wq[tx], wq[3 + ty]), wq[6 + tz]);
                    val[0] = __fmaf_rn(w, Nr[3 * sq], val[0]);
                    val[1] = __fmaf_rn(w, Nr[3 * sq + 1], val[1]);
                    val[2] = __fmaf_rn(w, Nr[3 * sq + 2], val[2]);
                }
            }
            V[3 * (i64)l] = val[0];
            V[3 * (i64)l + 1] = val[1];
            V[3 * (i64)l + 2] = val[2];
        }
    }
}

// Depths D-2 / D-3: one warp per node over the CONCATENATED slot ranges of its 27 neighbours.
// A neighbour of a surface octree often owns only 8 or 16 depth-D slots, so a lane-strided loop
// per neighbour leaves most lanes idle; here element e of the concatenation
// goes to lane e mod 32 and every lane advances its own segment cursor (the cursor only moves
// forward: amortised O(1) shared-memory look-ups per element).  Terms as in k_divergence_leaf;
// the double-precision summation order differs (tolerance stated in tests/test_parity_gpu.py).
__global__ void __launch_bounds__(256) k_divergence_flat(const float* __restrict__ V, const ushort4* __restrict__ offs, const int* __restrict__ neighs,
                                                         const int* __restrict__ didx, const int* __restrict__ dnum, const float* __restrict__ dfRow,
                                                         int base, int count, int baseD, int k /* 2^(D-d) */, float* __restrict__ divg) {
    __shared__ int sEnd[8][28], sOff[8][28];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int l = blockIdx.x * 8 + wp; l < count; l += gridDim.x * 8) {
        const int i = base + l;
        const ushort4 o = offs[i];
        const int bx = k * ((int)o.x - 1), by = k * ((int)o.y - 1), bz = k * ((int)o.z - 1);
        int s0 = 0, cnt = 0;
        if (lane < 27) {
            const int n = neighs[27 * (i64)i + lane];
            if (n >= 0) { s0 = didx[n]; cnt = dnum[n]; }
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 26);
        __syncwarp();
        if (lane < 27) { sEnd[wp][lane] = incl; sOff[wp][lane] = s0 - (incl - cnt); }
        if (lane == 27) { sEnd[wp][27] = 0x7fffffff; sOff[wp][27] = 0; }
        __syncwarp();
        double val = 0.0;
        int seg = 0, segEnd = sEnd[wp][0];
        for (int e = lane; e < total; e += 32) {
            while (e >= segEnd) segEnd = sEnd[wp][++seg];
            const int s = sOff[wp][seg] + e;
            const ushort4 so = offs[baseD + s];
            const float u0 = dfRow[(int)so.x - bx], u1 = dfRow[(int)so.y - by], u2 = dfRow[(int)so.z - bz];
            float dp = __fmul_rn(V[3 * (i64)s], u0);                  // DotProduct (main.cu:966-972)
            dp = __fmaf_rn(V[3 * (i64)s + 1], u1, dp);
            dp = __fmaf_rn(V[3 * (i64)s + 2], u2, dp);
            val += (double)dp;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
        if (lane == 0) divg[i] = (float)val;
    }
}

// The two finest depths without the slot indirections: at depth D a neighbour IS its slot and the
// table index is the neighbour direction; at depth D-1 the slots of a neighbour are its 8
// children (one aligned 96-byte record of V) and the index is 2*(direction+1) + child bit.
// Same terms, same order, same arithmetic as the reference (computeEncodedFinerNodesDivergence, main.cu:1007-1056).
__global__ void __launch_bounds__(256) k_divergence_leaf(const float* __restrict__ V, const int* __restrict__ neighs, const float* __restrict__ dfRow,
                                                         int baseD, int first, int count, float* __restrict__ divg) {
    const float r0 = dfRow[0], r1 = dfRow[1], r2 = dfRow[2];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < count; l += gridDim.x * blockDim.x) {
        const int* nb = neighs + 27 * (i64)(first + l);
        double val = 0.0;
#pragma unroll
        for (int j = 0; j < 27; j++) {
            int n = nb[j];
            if (n < 0) continue;
            const float* v = V + 3 * (i64)(n - baseD);
            const float u0 = (j / 9) == 0 ? r0 : ((j / 9) == 1 ? r1 : r2);
            const float u1 = ((j / 3) % 3) == 0 ? r0 : (((j / 3) % 3) == 1 ? r1 : r2);
            const float u2 = (j % 3) == 0 ? r0 : ((j % 3) == 1 ? r1 : r2);
            float dp = __fmul_rn(v[0], u0);
            dp = __fmaf_rn(v[1], u1, dp);
            dp = __fmaf_rn(v[2], u2, dp);
            val += (double)dp;
        }
        divg[first + l] = (float)val;
    }
}
__global__ void __launch_bounds__(256) k_divergence_dm1(const float* __restrict__ V, const int* __restrict__ neighs, const int* __restrict__ child0,
                                                        const float* __restrict__ dfRow, int base, int count, int baseD, float* __restrict__ divg) {
    float row[6];
#pragma unroll
    for (int t = 0; t < 6; t++) row[t] = dfRow[t];
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < count; l += gridDim.x * blockDim.x) {
        const int* nb = neighs + 27 * (i64)(base + l);
        double val = 0.0;
#pragma unroll
        for (int j = 0; j < 27; j++) {
            int n = nb[j];
            if (n < 0) continue;
            int c0 = child0[n];
            if (c0 < 0) continue;
            const float4* v4 = reinterpret_cast<const float4*>(V + 3 * (i64)(c0 - baseD));
            float v[24];
#pragma unroll
            for (int t = 0; t < 6; t++) { float4 a = v4[t]; v[4 * t] = a.x; v[4 * t + 1] = a.y; v[4 * t + 2] = a.z; v[4 * t + 3] = a.w; }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float u0 = row[2 * (j / 9) + (q >> 2)], u1 = row[2 * ((j / 3) % 3) + ((q >> 1) & 1)], u2 = row[2 * (j % 3) + (q & 1)];
                float dp = __fmul_rn(v[3 * q], u0);
                dp = __fmaf_rn(v[3 * q + 1], u1, dp);
                dp = __fmaf_rn(v[3 * q + 2], u2, dp);
                val += (double)dp;
            }
        }
        divg[base + l] = (float)val;
    }
}

// ---- coarse depths: scatter form.  A coarse node has up to 27 * 8^(D-d) terms, far too many for
// one thread block (the reference runs depths 0-4 as a host loop with one launch set per node,
// main.cu:3419-3458).  Work item = (node n, chunk of <= kChunk of the depth-D slots under n);
// every slot contributes to the <= 27 neighbours o of n, so a block accumulates 27 partial sums
// in double, reduces them and issues at most 27 atomicAdd(double).
constexpr int kChunk = 8192;
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
    {
        DBuf<float> W;
        PRB_TRY(W.alloc(9 * (size_t)c.N, st));
        PRB_LAUNCH(c, k_splat_weights, grid_for(c, c.N, 128, 16), 128, 0, c.dMaxDepthFn.p, c.P.p, c.p2n.p, c.offs.p + c.base[D], c.N, width, W.p);
        PRB_LAUNCH(c, k_splat, grid_for(c, (i64)c.cnt[D], 128, 16), 128, 0, W.p, c.Nr.p, c.neighs.p, c.pidx.p, c.pnum.p, c.base[D], c.cnt[D], c.V.p);
        W.release();
    }
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
        // multi-GPU: at the sharded depths a rank only needs the right-hand side of its own rows
        const bool sh = c.mg.active() && d >= c.shardFrom;
        const int first = sh ? c.rowLo[d][c.mg.rank] : c.base[d];
        const int n = sh ? c.rowLo[d][c.mg.rank + 1] - first : c.cnt[d];
        if (n <= 0) continue;
        if (d == D)
            PRB_LAUNCH(c, k_divergence_leaf, grid_for(c, n, 256), 256, 0, c.V.p, c.neighs.p, row, c.base[D], first, n, c.divg.p);
        else if (d == D - 1)
            PRB_LAUNCH(c, k_divergence_dm1, grid_for(c, n, 256), 256, 0, c.V.p, c.neighs.p, c.child0.p, row, first, n, c.base[D], c.divg.p);
        else
            PRB_LAUNCH(c, k_divergence_flat, grid_for(c, (i64)n * 32, 256), 256, 0, c.V.p, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, row, first, n, c.base[D], k, c.divg.p);
    }
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

}  // namespace prb
