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
#include <algorithm>

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
                                               int baseD, int gFirst, int gEnd, float* __restrict__ V) {
    __shared__ int2 sInfo[4][4][64];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int nGroups = gEnd, nSteps = (gEnd - gFirst + 3) >> 2;       // sibling groups [gFirst, gEnd) of depth D (multi-GPU: this rank's share)
    for (int st = blockIdx.x * 4 + wp; st < nSteps; st += gridDim.x * 4) {
        const int g0 = gFirst + 4 * st;
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


// =================================================================================================
// Divergence, block-table form (default).  The work is split by how many depth-D slots a node sees:
//   * depths D and D-1 (94 % of the nodes): k_div_fine.  One depth-D SUPER-GROUP (solver.cu / k_sg_table:
//     the 4x4x4 cube of 8-slot blocks around the children of a sibling group of depth D-1) holds every
//     slot that the group's 64 depth-D rows AND its 8 depth-(D-1) rows read, so its 512 slots x 12 B
//     of V are staged once in shared memory (cp.async, zero fill for absent blocks) instead of being
//     gathered 27 (x 8) times per row through the 108-byte neighbour rows.  A quarter-warp owns a
//     super-group; lane li owns block li (its 8 leaf rows) and the depth-(D-1) node above that block.
//     Every row still adds its terms in the reference's order (neighbour j = 0..26, then slot 0..7;
//     float product, double sum: main.cu:1020-1058), so both depths stay bit-identical: an absent
//     neighbour contributes an exact +0.0 instead of being skipped.
//   * depths <= D-2: the terms of a row are V_s . (T[x], T[y], T[z]) with T indexed by the per-axis
//     offset difference ONLY (no cross-axis factor), so a node n can be summarised by
//     three "profiles" P_n[axis][t] = sum of V_axis over the depth-D slots under n whose offset along
//     `axis` is t (t < 2^(D-d)), built bottom-up in double (k_profile_d2, k_profile_up), and
//     b_o = sum_j sum_axis sum_t T[k (d_axis(j) + 1) + t] P_{n_j}[axis][t]   (k_div_coarse).
//     27 * 3 * 2^(D-d) multiply-adds per node instead of 27 * 8^(D-d) -- this replaces both the
//     warp-per-node gather (D-2, D-3) and the atomic scatter (<= D-4) of the first version, and the
//     reference's per-node host loop for depths 0-4 (main.cu:3419-3458).  Summation order differs from
//     the reference there (double throughout): <= 1e-7 rel-L2 per depth, tolerance 1e-6 in the tests.
// Shared-memory layout of a staged cube, in 16-byte groups: block u = (ux, uy, uz) starts at group
// ux * 113 + uy * 28 + uz * 6 (6 groups = 8 slots x 12 B; the pads make the strides 1, 4, 6 mod 8), so the eight
// lanes of a quarter-warp -- blocks {1,2}^3 plus a common offset -- always read eight different 16-byte bank groups:
// every LDS.128 below is conflict free.
constexpr int kDfWarps = 7;
constexpr int kDfGx = 113, kDfGy = 28, kDfGz = 6;
constexpr int kDfCubeFloats = 4 * (4 * kDfGx);         // 452 groups
__host__ __device__ constexpr int df_block_group(int u) { return (u >> 4) * kDfGx + ((u >> 2) & 3) * kDfGy + (u & 3) * kDfGz; }
__global__ void __launch_bounds__(kDfWarps * 32, 1) k_div_fine(const float* __restrict__ V, const int* __restrict__ sgTab, const float* __restrict__ rowD, const float* __restrict__ rowDm1,
                                                               int baseD, int sgFirst, int nSgRange, int leafSg0, int leafSg1, int dm1Sg0, int dm1Sg1, float* __restrict__ divg) {
    extern __shared__ __align__(16) float sDf[];        // [warp][4 cubes][kDfCubeFloats] then [warp][4][64] table
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, q = lane >> 3, li = lane & 7;
    float* cube = sDf + (size_t)(wp * 4 + q) * kDfCubeFloats;
    int* tab = reinterpret_cast<int*>(sDf + (size_t)kDfWarps * 4 * kDfCubeFloats) + (wp * 4 + q) * 64;
    const unsigned cubeS = (unsigned)__cvta_generic_to_shared(cube);
    float r0[3], r1[6];
#pragma unroll
    for (int t = 0; t < 3; t++) r0[t] = rowD[t];
#pragma unroll
    for (int t = 0; t < 6; t++) r1[t] = rowDm1[t];
    const int myBlock = (1 + (li >> 2)) * 16 + (1 + ((li >> 1) & 1)) * 4 + (1 + (li & 1));
    const float4* mine = reinterpret_cast<const float4*>(cube) + df_block_group(myBlock);     // group pointer of the lane's own block
    const int nSteps = (nSgRange + 3) >> 2;
    for (int st = blockIdx.x * kDfWarps + wp; st < nSteps; st += gridDim.x * kDfWarps) {
        const int sg = sgFirst + 4 * st + q;
        const bool have = 4 * st + q < nSgRange;
        __syncwarp();
        // ---- table row, then the 384 16-byte chunks of the cube: consecutive lanes copy consecutive chunks
        if (have) {
            const int4* src = reinterpret_cast<const int4*>(sgTab + (size_t)sg * 64) + 2 * li;
            const int4 a = src[0], b = src[1];
            reinterpret_cast<int4*>(tab)[2 * li] = a;
            reinterpret_cast<int4*>(tab)[2 * li + 1] = b;
        }
        __syncwarp();
        if (have) {
#pragma unroll 8
            for (int it = 0; it < 48; it++) {
                const int ch = it * 8 + li, blk = ch / 6, part = ch - blk * 6;
                const int id = tab[blk];
                const float* src = V + (id >= 0 ? 3 * (size_t)(id - baseD) + 4 * part : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(cubeS + 16u * (unsigned)(df_block_group(blk) + part)), "l"(src), "r"(id >= 0 ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        if (!have) continue;
        // ---- depth D-1: the node above block li; 27 neighbour blocks x 8 slots, reference order
        if (sg >= dm1Sg0 && sg < dm1Sg1) {
            double val = 0.0;
#pragma unroll
            for (int j = 0; j < 27; j++) {
                const int dx = j / 9 - 1, dy = (j / 3) % 3 - 1, dz = j % 3 - 1;
                const float4* b4 = mine + (dx * kDfGx + dy * kDfGy + dz * kDfGz);
                float v[24];
#pragma unroll
                for (int t = 0; t < 6; t++) { const float4 a = b4[t]; v[4 * t] = a.x; v[4 * t + 1] = a.y; v[4 * t + 2] = a.z; v[4 * t + 3] = a.w; }
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    float dp = __fmul_rn(v[3 * s], r1[2 * (dx + 1) + (s >> 2)]);
                    dp = __fmaf_rn(v[3 * s + 1], r1[2 * (dy + 1) + ((s >> 1) & 1)], dp);
                    dp = __fmaf_rn(v[3 * s + 2], r1[2 * (dz + 1) + (s & 1)], dp);
                    val += (double)dp;
                }
            }
            divg[1 + 8 * (sg - 1) + li] = (float)val;
        }
        // ---- depth D: the 8 rows of block li.  The 4x4x4 window of cells around the block is visited in
        // lexicographic order (x layer by x layer), which is neighbour order j for each of the rows it feeds.  A layer is
        // the half with x bit sx of the 9 blocks (bx; by, bz): 12 contiguous floats each
        const int own = tab[myBlock];
        if (own >= 0 && sg >= leafSg0 && sg < leafSg1) {
            double val[8];
#pragma unroll
            for (int c = 0; c < 8; c++) val[c] = 0.0;
#pragma unroll
            for (int wx = -1; wx <= 2; wx++) {
                const int bx = (wx + 2) / 2 - 1, sx = wx & 1;
                float L[9][12];
#pragma unroll
                for (int b = 0; b < 9; b++) {
                    const float4* b4 = mine + (bx * kDfGx + (b / 3 - 1) * kDfGy + (b % 3 - 1) * kDfGz) + 3 * sx;
#pragma unroll
                    for (int t = 0; t < 3; t++) { const float4 a = b4[t]; L[b][4 * t] = a.x; L[b][4 * t + 1] = a.y; L[b][4 * t + 2] = a.z; L[b][4 * t + 3] = a.w; }
                }
#pragma unroll
                for (int wy = -1; wy <= 2; wy++)
#pragma unroll
                    for (int wz = -1; wz <= 2; wz++) {
                        const int b = ((wy + 2) / 2) * 3 + ((wz + 2) / 2), sl = ((wy & 1) << 1) | (wz & 1);
                        const float vx = L[b][3 * sl], vy = L[b][3 * sl + 1], vz = L[b][3 * sl + 2];
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const int dx = wx - ((c >> 2) & 1), dy = wy - ((c >> 1) & 1), dz = wz - (c & 1);
                            if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) continue;
                            float dp = __fmul_rn(vx, r0[dx + 1]);
                            dp = __fmaf_rn(vy, r0[dy + 1], dp);
                            dp = __fmaf_rn(vz, r0[dz + 1], dp);
                            val[c] += (double)dp;
                        }
                    }
            }
            float4* o = reinterpret_cast<float4*>(divg + own);
            o[0] = make_float4((float)val[0], (float)val[1], (float)val[2], (float)val[3]);
            o[1] = make_float4((float)val[4], (float)val[5], (float)val[6], (float)val[7]);
        }
    }
}
// profiles of depth D-2: 8 lanes per node Q, lane li = child li of Q (depth D-1), whose children are one
// block of 8 depth-D slots.  P[Q][axis][t], t = 2 (bit of the child) + (bit of the slot)
__global__ void __launch_bounds__(256) k_profile_d2(const float* __restrict__ V, const int* __restrict__ child0, int base2, int count2, int baseD, double* __restrict__ prof) {
    const int lane = threadIdx.x & 31, li = lane & 7;
    const int nWork = (count2 + 3) >> 2;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nWork; w += (gridDim.x * blockDim.x) >> 5) {
        const int l = 4 * w + (lane >> 3);
        const bool have = l < count2;
        double sx[2] = {0.0, 0.0}, sy[2] = {0.0, 0.0}, sz[2] = {0.0, 0.0};
        if (have) {
            const int c1 = child0[base2 + l];
            const int c2 = c1 >= 0 ? child0[c1 + li] : -1;
            if (c2 >= 0) {
                const float4* b4 = reinterpret_cast<const float4*>(V + 3 * (size_t)(c2 - baseD));
                float v[24];
#pragma unroll
                for (int t = 0; t < 6; t++) { const float4 a = b4[t]; v[4 * t] = a.x; v[4 * t + 1] = a.y; v[4 * t + 2] = a.z; v[4 * t + 3] = a.w; }
#pragma unroll
                for (int s = 0; s < 8; s++) { sx[(s >> 2) & 1] += (double)v[3 * s]; sy[(s >> 1) & 1] += (double)v[3 * s + 1]; sz[s & 1] += (double)v[3 * s + 2]; }
            }
        }
        // x: lanes with the same bit 2 of li; y: same bit 1; z: same bit 0 (fixed butterfly order)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            sx[b] += __shfl_xor_sync(0xffffffffu, sx[b], 1); sx[b] += __shfl_xor_sync(0xffffffffu, sx[b], 2);
            sy[b] += __shfl_xor_sync(0xffffffffu, sy[b], 1); sy[b] += __shfl_xor_sync(0xffffffffu, sy[b], 4);
            sz[b] += __shfl_xor_sync(0xffffffffu, sz[b], 2); sz[b] += __shfl_xor_sync(0xffffffffu, sz[b], 4);
        }
        if (have) {
            double* P = prof + 12 * (size_t)l;
            if ((li & 3) == 0) { P[2 * (li >> 2)] = sx[0]; P[2 * (li >> 2) + 1] = sx[1]; }
            if ((li & 5) == 0) { P[4 + 2 * ((li >> 1) & 1)] = sy[0]; P[4 + 2 * ((li >> 1) & 1) + 1] = sy[1]; }
            if ((li & 6) == 0) { P[8 + 2 * (li & 1)] = sz[0]; P[8 + 2 * (li & 1) + 1] = sz[1]; }
        }
    }
}
// profiles of depth d from those of depth d + 1: entry t of a node = the sum over its four children on side
// t / (k/2) of the axis of their entries t mod (k/2), ascending child order
__global__ void __launch_bounds__(256) k_profile_up(const int* __restrict__ child0, int baseHere, int countHere, int baseBelow, int lk /* log2 k */,
                                                    const double* __restrict__ below, double* __restrict__ here) {
    const int k = 1 << lk, kh = k >> 1;
    const i64 total = (i64)countHere * 3 * k;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const int t = (int)(e & (k - 1)), q = (int)(e >> lk), a = q % 3, l = q / 3;
        const int c0 = child0[baseHere + l];
        double s = 0.0;
        if (c0 >= 0) {
            const int side = t >> (lk - 1), tt = t & (kh - 1), bit = 2 - a;
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (((c >> bit) & 1) == side) s += below[((size_t)(c0 + c - baseBelow) * 3 + a) * kh + tt];
        }
        here[e] = s;
    }
}
// b_o for the nodes of ONE depth d <= D-2 from the profiles of their 27 neighbours.
// k = 2^(D-d) <= 8: one warp per node, lane j = neighbour j: its 3k profile entries are one contiguous run of doubles.
template <int LK>
__global__ void __launch_bounds__(256) k_div_coarse_small(const int* __restrict__ neighs, const double* __restrict__ prof, const float* __restrict__ row, int base, int first, int count,
                                                          float* __restrict__ divg) {
    constexpr int K = 1 << LK;
    const int lane = threadIdx.x & 31;
    double T[3 * K];                                      // the 3k table values this lane's direction digits select, as doubles
    {
        const int j = lane < 27 ? lane : 0;
        const int dj[3] = {j / 9, (j / 3) % 3, j % 3};
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int t = 0; t < K; t++) T[a * K + t] = (double)row[(dj[a] << LK) + t];
    }
    for (int l = first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < first + count; l += (gridDim.x * blockDim.x) >> 5) {
        double acc = 0.0;
        if (lane < 27) {
            const int n = neighs[27 * (i64)(base + l) + lane];
            if (n >= 0) {
                const double2* P = reinterpret_cast<const double2*>(prof + (size_t)(n - base) * 3 * K);
#pragma unroll
                for (int e = 0; e < 3 * K / 2; e++) { const double2 v = P[e]; acc += T[2 * e] * v.x; acc += T[2 * e + 1] * v.y; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) divg[base + l] = (float)acc;
    }
}
// k >= 16: a group of G threads (a warp, or a whole block for the coarse depths with few nodes and thousands of entries per
// neighbour) per node; the 81 (neighbour, axis) runs of k entries are flattened over the group, entries fastest (coalesced)
template <int G>
__global__ void __launch_bounds__(256) k_div_coarse_wide(const int* __restrict__ neighs, const double* __restrict__ prof, const float* __restrict__ row, int base, int first, int nRows, int lk,
                                                         float* __restrict__ divg) {
    __shared__ double sRed[8];
    __shared__ int sNb[8][27];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int k = 1 << lk, total = 81 << lk;
    const int groupsPerBlock = 256 / G, gInBlock = threadIdx.x / G, tInGroup = threadIdx.x % G;
    const int count = first + nRows;
    for (int l0 = first + blockIdx.x * groupsPerBlock; l0 < count; l0 += gridDim.x * groupsPerBlock) {
        const int l = l0 + gInBlock;
        __syncthreads();
        if (tInGroup < 27 && l < count) sNb[gInBlock][tInGroup] = neighs[27 * (i64)(base + l) + tInGroup];
        __syncthreads();
        double acc = 0.0;
        if (l < count)
            for (int e = tInGroup; e < total; e += G) {
                const int t = e & (k - 1), q = e >> lk, j = q / 3, a = q - 3 * j;
                const int n = sNb[gInBlock][j];
                if (n < 0) continue;
                const int dj = a == 0 ? j / 9 : (a == 1 ? (j / 3) % 3 : j % 3);          // d_axis(j) + 1
                acc += (double)row[(dj << lk) + t] * prof[((size_t)(n - base) * 3 + a) * k + t];
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (G == 32) {
            if (lane == 0 && l < count) divg[base + l] = (float)acc;
        } else {
            if (lane == 0) sRed[wp] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) v += sRed[w];
                divg[base + l] = (float)v;
            }
        }
    }
}

int stage_splat(Context& c) {
    const int D = c.D;
    cudaStream_t st = c.stream;
    const bool shard = c.mg.active() && D >= c.shardFrom;      // multi-GPU: a rank splats the slots of its Morton range, the field is then gathered
    if (shard) {
        c.Vp = c.mg.alloc<float>(3 * (size_t)c.cnt[D], &c.mgVOff);
        if (!c.Vp) { set_error("multi-GPU arena too small for the vector field (12 bytes per depth-D slot)"); return PRB_ERR_NOMEM; }
    } else {
        PRB_TRY(c.V.alloc(3 * (size_t)c.cnt[D], st));
        c.Vp = c.V.p;
    }
    PRB_TRY(c.divg.alloc((size_t)c.M + 16, st));       // padded like x: node 1 (every sibling block) on a 32-byte boundary
    c.divgv = c.divg.p + 7;
    float width = (float)(1.0 / (1 << D));
    {
        DBuf<float> W;
        PRB_TRY(W.alloc(9 * (size_t)c.N, st));
        PRB_LAUNCH(c, k_splat_weights, grid_for(c, c.N, 128, 16), 128, 0, c.dMaxDepthFn.p, c.P.p, c.p2n.p, c.offs.p + c.base[D], c.N, width, W.p);
        const int g0 = shard ? (c.rowLo[D][c.mg.rank] - c.base[D]) / 8 : 0, g1 = shard ? (c.rowLo[D][c.mg.rank + 1] - c.base[D]) / 8 : c.cnt[D] / 8;
        if (g1 > g0) PRB_LAUNCH(c, k_splat, grid_for(c, (i64)(g1 - g0) * 8, 128, 16), 128, 0, W.p, c.Nr.p, c.neighs.p, c.pidx.p, c.pnum.p, c.base[D], g0, g1, c.Vp);
        W.release();
    }
    mark(c, "splat:kernels");
    if (shard) {
        long long lo[kMaxRanks + 1];
        for (int r = 0; r <= c.mg.world; r++) lo[r] = c.rowLo[D][r] - c.base[D];
        PRB_TRY(mg_allgather(c, c.mgVOff, 12, lo));
    }
    mark(c, "splat:done");
    PRB_CUDA(cudaEventRecord(c.ev[3], st));
    PRB_TRY(stage_divergence(c));
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

static int stage_divergence_blocks(Context& c) {
    const int D = c.D;
    cudaStream_t st = c.stream;
    const bool mg = c.mg.active();
    auto sg_start = [&](int d) { return d <= 1 ? 0 : 1 + (c.base[d - 1] - 1) / 8; };      // first super-group of depth d
    // ---- depths D and D-1 (multi-GPU: the super-groups that hold this rank's rows of either depth)
    {
        int leaf0 = sg_start(D), leaf1 = sg_start(D + 1), dm0 = leaf0, dm1 = leaf1;
        if (mg && D >= c.shardFrom) { leaf0 = c.sgLo[D][c.mg.rank]; leaf1 = c.sgLo[D][c.mg.rank + 1]; }
        if (mg && D - 1 >= c.shardFrom) {
            const int ra = c.rowLo[D - 1][c.mg.rank], rb = c.rowLo[D - 1][c.mg.rank + 1];
            dm0 = rb > ra ? 1 + (ra - 1) / 8 : leaf0;
            dm1 = rb > ra ? 1 + (rb - 1 + 7) / 8 : leaf0;
        }
        const int first = std::min(leaf0, dm0), last = std::max(leaf1, dm1);
        if (last > first) {
            const size_t smem = (size_t)kDfWarps * 4 * kDfCubeFloats * sizeof(float) + (size_t)kDfWarps * 4 * 64 * sizeof(int);
            PRB_CUDA(cudaFuncSetAttribute(k_div_fine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int steps = (last - first + 3) / 4;
            const int grid = std::max(1, std::min(c.smCount, (steps + kDfWarps - 1) / kDfWarps));
            PRB_LAUNCH(c, k_div_fine, grid, kDfWarps * 32, smem, c.Vp, c.sgTab.p, c.dDfT.p + c.tab.dfOffset[D], c.dDfT.p + c.tab.dfOffset[D - 1], c.base[D], first, last - first,
                       leaf0, leaf1, dm0, dm1, c.divgv);
        }
    }
    mark(c, "div:fine");
    if (D < 2) return PRB_OK;
    // ---- depths <= D-2 through the per-axis profiles (every rank computes all of them: 6 % of the nodes)
    std::vector<size_t> off(D + 1, 0);
    size_t total = 0;
    for (int d = D - 2; d >= 0; --d) { off[d] = total; total += (size_t)c.cnt[d] * 3 * ((size_t)1 << (D - d)); }
    DBuf<double> prof;
    PRB_TRY(prof.alloc(total, st));
    PRB_LAUNCH(c, k_profile_d2, grid_for(c, (i64)c.cnt[D - 2] * 8, 256), 256, 0, c.Vp, c.child0.p, c.base[D - 2], c.cnt[D - 2], c.base[D], prof.p + off[D - 2]);
    for (int d = D - 3; d >= 0; --d)
        PRB_LAUNCH(c, k_profile_up, grid_for(c, (i64)c.cnt[d] * 3 * ((i64)1 << (D - d)), 256), 256, 0, c.child0.p, c.base[d], c.cnt[d], c.base[d + 1], D - d, prof.p + off[d + 1], prof.p + off[d]);
    for (int d = D - 2; d >= 0; --d) {
        const float* row = c.dDfT.p + c.tab.dfOffset[d];
        const int lk = D - d;
        // multi-GPU: at a sharded depth a rank only needs the right-hand side of its own rows (the profiles are complete everywhere)
        const bool sh = mg && d >= c.shardFrom;
        const int first = sh ? c.rowLo[d][c.mg.rank] - c.base[d] : 0, n = sh ? c.rowLo[d][c.mg.rank + 1] - c.rowLo[d][c.mg.rank] : c.cnt[d];
        if (n <= 0) continue;
        if (lk == 2) PRB_LAUNCH(c, k_div_coarse_small<2>, grid_for(c, (i64)n * 32, 256), 256, 0, c.neighs.p, prof.p + off[d], row, c.base[d], first, n, c.divgv);
        else if (lk == 3) PRB_LAUNCH(c, k_div_coarse_small<3>, grid_for(c, (i64)n * 32, 256), 256, 0, c.neighs.p, prof.p + off[d], row, c.base[d], first, n, c.divgv);
        else if (lk == 4) PRB_LAUNCH(c, k_div_coarse_wide<32>, grid_for(c, (i64)n * 32, 256), 256, 0, c.neighs.p, prof.p + off[d], row, c.base[d], first, n, lk, c.divgv);
        else PRB_LAUNCH(c, k_div_coarse_wide<256>, std::max(1, std::min(n, c.smCount * 8)), 256, 0, c.neighs.p, prof.p + off[d], row, c.base[d], first, n, lk, c.divgv);
    }
    mark(c, "div:coarse");
    prof.release();
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

int stage_divergence(Context& c) {
    if (c.divMode == 1) return stage_divergence_blocks(c);
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
            PRB_LAUNCH(c, k_divergence_scatter, (int)nItems, 256, 0, c.Vp, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, c.dDfT.p, c.dDfOffset.p, itemBase.p,
                       nCoarse, c.base[D], D, accum.p);
        PRB_LAUNCH(c, k_divergence_finish, grid_for(c, nCoarse, 256), 256, 0, accum.p, nCoarse, c.divgv);
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
            PRB_LAUNCH(c, k_divergence_leaf, grid_for(c, n, 256), 256, 0, c.Vp, c.neighs.p, row, c.base[D], first, n, c.divgv);
        else if (d == D - 1)
            PRB_LAUNCH(c, k_divergence_dm1, grid_for(c, n, 256), 256, 0, c.Vp, c.neighs.p, c.child0.p, row, first, n, c.base[D], c.divgv);
        else
            PRB_LAUNCH(c, k_divergence_flat, grid_for(c, (i64)n * 32, 256), 256, 0, c.Vp, c.offs.p, c.neighs.p, c.didx.p, c.dnum.p, row, first, n, c.base[D], k, c.divgv);
    }
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

}  // namespace prb
