// Iso-value (A11) and iso-surface extraction (A12): corner values, marching cubes on the
// depth-D slots, and the refinement passes over empty coarser leaves crossed by the surface.
//
// Replaces calculatePointsImplicitFunctionValue (main.cu:1334-1381), the vertex / edge / face
// array construction (main.cu:1423-2082, 2827-2955), the depth-D marching-cubes pass
// (main.cu:2259-2757, 3652-3795) and the subdivision passes (main.cu:2957-3217, 3799-4564).
// The reference materialises 56-byte VertexNode / 24-byte EdgeNode / 20-byte FaceNode arrays with
// copy_if over 8/12/6 candidates per node and writes back-pointers into its 276-byte nodes.
// Here ownership ("min-key incident cell", which equals min node index inside a depth) is decided
// on the fly from the neighbour table, corner values live in one float[8] record per cell at the
// owner's slot, crossed edges are a 12-bit mask per cell, and output addresses come from two
// scans.  Vertex and triangle ORDER equal the reference's (edges by owner cell then edge kind;
// triangles by cell then case-table order), so meshes compare index by index.
#include "common.cuh"
#include "mg_device.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include "mc_case_table.h"
#include "scan.cuh"

namespace prb {

__constant__ signed char cMcTri[256][16];
__constant__ unsigned char cMcCount[256];
__constant__ unsigned char cEdgeVertex[12][2];   // ring indices, ascending (MarchingCubes.cuh:693-706)
__constant__ unsigned char cFaceEdgeMask[6][2];  // 12-bit mask of the 4 edges of each face (MarchingCubes.cuh:734-741), lo/hi byte

// __constant__ symbols exist once per DEVICE: readiness is tracked per device (a process may hold contexts on several GPUs)
static std::mutex g_tablesMu;
static bool g_tablesReady[64] = {false};
static int upload_mc_tables(int device) {
    std::lock_guard<std::mutex> lk(g_tablesMu);
    if (device < 0 || device >= 64) { set_error("device index out of range"); return PRB_ERR_ARG; }
    if (g_tablesReady[device]) return PRB_OK;
    signed char tri[256][16];
    unsigned char cnt[256], ev[12][2], fe[6][2];
    for (int c = 0; c < 256; c++) {
        int n = 0;
        for (int k = 0; k < 16; k++) {
            char h = kMcCaseHex[16 * c + k];
            int v = (h == 'f') ? -1 : (h <= '9' ? h - '0' : h - 'a' + 10);
            tri[c][k] = (signed char)v;
            if (v >= 0) n++;
        }
        cnt[c] = (unsigned char)(n / 3);
    }
    for (int e = 0; e < 12; e++) {
        int o = e >> 2, r[2];
        for (int s = 0; s < 2; s++) {
            int xyz[3];
            for (int a = 0; a < 3; a++) xyz[a] = (a == o) ? s : edge_off(e, a);
            r[s] = ring_index(xyz[0], xyz[1], xyz[2]);
        }
        ev[e][0] = (unsigned char)(r[0] < r[1] ? r[0] : r[1]);
        ev[e][1] = (unsigned char)(r[0] < r[1] ? r[1] : r[0]);
    }
    for (int f = 0; f < 6; f++) {
        int m = 0;
        for (int e = 0; e < 12; e++) if (edge_off(e, f >> 1) == (f & 1)) m |= 1 << e;
        fe[f][0] = (unsigned char)(m & 255);
        fe[f][1] = (unsigned char)(m >> 8);
    }
    PRB_CUDA(cudaMemcpyToSymbol(cMcTri, tri, sizeof(tri)));
    PRB_CUDA(cudaMemcpyToSymbol(cMcCount, cnt, sizeof(cnt)));
    PRB_CUDA(cudaMemcpyToSymbol(cEdgeVertex, ev, sizeof(ev)));
    PRB_CUDA(cudaMemcpyToSymbol(cFaceEdgeMask, fe, sizeof(fe)));
    g_tablesReady[device] = true;
    return PRB_OK;
}

// value at t of base function #fi (4 cumulative pieces x (c0..c3,start); ConfirmedPPolynomial.cuh:79-91)
__device__ __forceinline__ float base_value(const float* __restrict__ baseFn, int fi, float t) {
    const float* f = baseFn + 20 * (i64)fi;
    float res = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (!(t > f[5 * i + 4])) break;
        float v = f[5 * i];
        float pw = t;
        v = __fmaf_rn(pw, f[5 * i + 1], v);
        pw = __fmul_rn(pw, t);
        v = __fmaf_rn(pw, f[5 * i + 2], v);
        res = __fadd_rn(res, v);
    }
    return res;
}
// sum over the 27 neighbours of `node` (depth d, offsets o) of x[n] * F_n(pos); j order, float
__device__ __forceinline__ void accumulate_level(float& val, const int* __restrict__ nb, ushort4 o, const float* __restrict__ x,
                                                 const float* __restrict__ baseFn, const float pos[3]) {
    int d = o.w, n = 1 << d, f0 = n - 1;
    float vx[3], vy[3], vz[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int ax = (int)o.x + k - 1, ay = (int)o.y + k - 1, az = (int)o.z + k - 1;
        vx[k] = (ax >= 0 && ax < n) ? base_value(baseFn, f0 + ax, pos[0]) : 0.f;
        vy[k] = (ay >= 0 && ay < n) ? base_value(baseFn, f0 + ay, pos[1]) : 0.f;
        vz[k] = (az >= 0 && az < n) ? base_value(baseFn, f0 + az, pos[2]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 27; j++) {
        int q = nb[j];
        if (q >= 0) val = __fmaf_rn(__fmul_rn(__fmul_rn(x[q], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], val);
    }
}

#define RV_ACC27(val, X, vx, vy, vz)                                                                                      \
    _Pragma("unroll") for (int j_ = 0; j_ < 27; j_++)                                                                     \
        val = __fmaf_rn(__fmul_rn(__fmul_rn((X)[j_], (vx)[j_ / 9]), (vy)[(j_ / 3) % 3]), (vz)[j_ % 3], val)

// ------------------------------------------------------------------ A11 iso value
// sum[0] = sum of chi over the samples (the reference's iso value is their plain mean, main.cu:3494-3496); sum[1], sum[2] = the
// density-weighted sums of the opt-in mode (not in the reference): weight 1 / (samples in the sample's ancestor cell
// at depth dk), so that the mean runs over the SURFACE rather than over the samples of an unevenly dense scan.
template <bool WEIGHTED>
__global__ void __launch_bounds__(128) k_point_values(const float* __restrict__ P, const int* __restrict__ p2n, i64 N, int baseD,
                                                      const int* __restrict__ neighs, const int* __restrict__ parent, const ushort4* __restrict__ offs,
                                                      const int* __restrict__ pnum, int dk,
                                                      const float* __restrict__ x, const float* __restrict__ baseFn, float* __restrict__ pv, double* __restrict__ sum) {
    double acc = 0.0, accW = 0.0, accWX = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (i64)gridDim.x * blockDim.x) {
        float pos[3] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        int now = baseD + p2n[i];
        float val = 0.f;
        int cellSamples = 1;
        while (now != -1) {
            const ushort4 o = offs[now];
            if (WEIGHTED && (int)o.w == dk) cellSamples = pnum[now];
            accumulate_level(val, neighs + 27 * (i64)now, o, x, baseFn, pos);
            now = parent[now];
        }
        pv[i] = val;
        acc += (double)val;
        if (WEIGHTED) {
            const double w = 1.0 / (double)(cellSamples > 0 ? cellSamples : 1);
            accW += w;
            accWX += w * (double)val;
        }
    }
    __shared__ double red[3][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        accW += __shfl_down_sync(0xffffffffu, accW, o);
        accWX += __shfl_down_sync(0xffffffffu, accWX, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = accWX; red[2][threadIdx.x >> 5] = accW; }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(sum + threadIdx.x, red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3]);
}

// multi-GPU: every rank writes its partial sums into slots 29..31 of every rank's header
__global__ void k_mg_publish_sum(MgDev mg, const double* __restrict__ v) {
    if (threadIdx.x < 3 && blockIdx.x == 0)
        for (int r = 0; r < mg.world; r++) mg.peerHdr[r]->slots[0][mg.rank][31 - threadIdx.x] = v[threadIdx.x];
}

int stage_iso(Context& c) {
    cudaStream_t st = c.stream;
    mark(c, "iso:begin");
    PRB_TRY(c.pointValue.alloc((size_t)c.N, st));
    DBuf<double> sum;
    PRB_TRY(sum.alloc(3, st));
    PRB_CUDA(cudaMemsetAsync(sum.p, 0, 3 * sizeof(double), st));
    const int dk = c.D >= 3 ? c.D - 3 : 0;                 // depth of the density estimate of the weighted mode
    // multi-GPU: the samples are split evenly; the partial sums meet in the arena header
    const i64 p0 = c.mg.active() ? (c.N * c.mg.rank) / c.mg.world : 0, p1 = c.mg.active() ? (c.N * (c.mg.rank + 1)) / c.mg.world : c.N;
    // (the weighted sums cost 0.2 ms on 5 M samples: only taken when the opt-in mode asks for them)
    if (p1 > p0) {
        if (c.isoDensityWeighted)
            PRB_LAUNCH(c, k_point_values<true>, grid_for(c, p1 - p0, 128, 16), 128, 0, c.P.p + 3 * p0, c.p2n.p + p0, p1 - p0, c.base[c.D], c.neighs.p, c.parent.p, c.offs.p, c.pnum.p, dk,
                       c.xv, c.dBaseFn.p, c.pointValue.p + p0, sum.p);
        else
            PRB_LAUNCH(c, k_point_values<false>, grid_for(c, p1 - p0, 128, 16), 128, 0, c.P.p + 3 * p0, c.p2n.p + p0, p1 - p0, c.base[c.D], c.neighs.p, c.parent.p, c.offs.p, c.pnum.p, dk,
                       c.xv, c.dBaseFn.p, c.pointValue.p + p0, sum.p);
    }
    double h[3] = {0, 0, 0};
    if (c.mg.active()) {
        PRB_LAUNCH(c, k_mg_publish_sum, 1, 32, 0, c.mg.dev(), sum.p);
        PRB_TRY(mg_barrier(c));
        double parts[kMaxRanks][32];
        PRB_CUDA(cudaMemcpyAsync(parts, &((MgHeader*)c.mg.arena)->slots[0][0][0], sizeof(parts), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        for (int r = 0; r < c.mg.world; r++) { h[0] += parts[r][31]; h[1] += parts[r][30]; h[2] += parts[r][29]; }
    } else {
        PRB_CUDA(cudaMemcpyAsync(h, sum.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
    }
    if (c.mg.active()) {
        int err = 0;
        PRB_CUDA(cudaMemcpyAsync(&err, &((MgHeader*)c.mg.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        if (err) { set_error("multi-GPU iso value: timed out waiting for a peer"); return PRB_ERR_CUDA; }
    }
    // thrust::reduce(float) + "isoValue /= count" (main.cu:3494-3496)
    float iso = (float)h[0];
    iso /= (float)c.N;
    c.isoPlain = iso;
    c.isoWeighted = h[2] > 0 ? (float)(h[1] / h[2]) : iso;
    c.iso = c.isoDensityWeighted ? c.isoWeighted : iso;
    sum.release();
    return PRB_OK;
}

// ------------------------------------------------------------------ topology helpers
// A pass works on a set of depth-D cells addressed by ids.  Real pass: id = node index, the
// neighbour table is `neighs`, candidates are ids >= 0.  Refinement pass: id = M + virtual
// index, table = vneigh (row 0 = id M), candidates are ids >= M (virtual cells only,
// main.cu:1609,2000).
struct Topo {
    const int* nbr;     // 27 ids per row
    int rowBase;        // id of row 0
    int minId;          // candidates: id >= minId
    int cellBase;       // id of the first depth-D cell of the pass
    int nCells;
};
__device__ __forceinline__ int topo_nb(const Topo& T, int id, int j) { return T.nbr[27 * (i64)(id - T.rowBase) + j]; }

// owner of corner j (bits x|y<<1|z<<2) of cell id: smallest candidate among the <= 8 incident
// cells (= min key, main.cu:1474-1484); m = which axes were stepped to reach it
__device__ __forceinline__ int corner_owner(const Topo& T, int id, int j, int& m) {
    int sx = (j & 1) ? 1 : -1, sy = (j & 2) ? 1 : -1, sz = (j & 4) ? 1 : -1;
    int best = 0x7fffffff;
    m = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int dx = (q & 1) ? sx : 0, dy = (q & 2) ? sy : 0, dz = (q & 4) ? sz : 0;
        int nb = topo_nb(T, id, 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1));
        if (nb >= T.minId && nb < best) { best = nb; m = q; }
    }
    return best;
}
// owner of edge e of cell id among the <= 4 incident cells; e2 = the edge's kind in the owner's frame
__device__ __forceinline__ int edge_owner(const Topo& T, int id, int e, int& e2) {
    int o = e >> 2, a0, a1;
    other_axes(o, a0, a1);
    int s0 = (e & 1) ? 1 : -1, s1 = (e & 2) ? 1 : -1;
    int best = 0x7fffffff, bm = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int d[3] = {0, 0, 0};
        if (q & 1) d[a0] = s0;
        if (q & 2) d[a1] = s1;
        int nb = topo_nb(T, id, 9 * (d[0] + 1) + 3 * (d[1] + 1) + (d[2] + 1));
        if (nb >= T.minId && nb < best) { best = nb; bm = q; }
    }
    e2 = (o << 2) | ((e & 1) ^ (bm & 1)) | ((((e >> 1) & 1) ^ ((bm >> 1) & 1)) << 1);
    return best;
}
// Corner values live at the owner: 8 floats per id (bit order).  Multi-GPU: the ids [lo[r], lo[r+1]) were evaluated by rank r and are
// read from its peer-mapped arena (plain NVLink loads; only cells at a shard boundary ever touch another rank's values).
struct ValView {
    const float* p[kMaxRanks];     // p[r][8 * (id - valBase) + slot]; world == 1: everything in p[0]
    int valBase, world;
    int lo[kMaxRanks + 1];
};
static ValView local_view(const float* vals, int valBase) {
    ValView W;
    for (int r = 0; r < kMaxRanks; r++) W.p[r] = vals;
    W.valBase = valBase; W.world = 1;
    for (int r = 0; r <= kMaxRanks; r++) W.lo[r] = 0;
    return W;
}
__device__ __forceinline__ int view_rank(const int* lo, int world, int id) {
    int r = 0;
    while (r + 1 < world && id >= lo[r + 1]) r++;
    return r;
}
__device__ __forceinline__ float val_at(const ValView& W, int ow, int slot) {
    const int r = W.world > 1 ? view_rank(W.lo, W.world, ow) : 0;
    return W.p[r][8 * (i64)(ow - W.valBase) + slot];
}
// the 8 corner values of a cell in ring order
__device__ __forceinline__ void cell_corner_values(const Topo& T, int id, const ValView& W, float v[8]) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int j = ring_to_bits(r), m;
        int ow = corner_owner(T, id, j, m);
        v[r] = val_at(W, ow, j ^ m);
    }
}

// ------------------------------------------------------------------ corner values, real tree (main.cu:2259-2326)
// One WARP per contiguous chunk of sibling groups.  The 8 siblings of a group share every
// ancestor, so the 27 neighbour values of the ancestor levels are staged once per group in shared
// memory and reused by all the corners the group owns (the 27 points of its 3x3x3 corner grid; a
// point is evaluated by the group that holds its owner cell).  Each point still adds its terms in
// the reference's order: own level, parents up to the root, then the finer nodes at that corner
// (main.cu:2277-2322).  Sibling groups are
// numbered in Morton order inside a depth, so consecutive groups share almost all ancestors: a
// warp walks a CONTIGUOUS chunk of groups and keeps the 27 neighbour values of every ancestor
// level in shared memory, re-gathering only the levels whose ancestor changed (1.3 levels per
// group on a surface instead of all d0 of them, and no dependent parent -> neighbour -> x chain per
// level).  The base-function values B_{l, a+k-1}((o + pc) w) depend only on (depth, level, o/2, pc, k)
// per axis; they come from a table filled once per context by k_build_bv with the same
// base_value() arithmetic (bit-identical to evaluating them in place).
struct BvTables {
    const float4* anc;            // [d0][l < d0][g < 2^(d0-1)][pc < 3] -> (k = 0, 1, 2, unused)
    const float4* own;            // [d0][g][pc] -> cube coordinate 0..3
    const float4* cellD;          // [l <= D][gc < 2^D] -> (k = 0, 1, 2, unused): level-l functions around depth-D cell gc at its UPPER corner (gc+1) w
    const float4* gridLo;         // [l <= D][P <= 2^D] -> level-l functions around node P >> (D-l), at the depth-D grid point P w
    int ancOff[kMaxDepth + 1];    // first float4 of depth d0
    int ownOff[kMaxDepth + 1];
};
__global__ void __launch_bounds__(256) k_build_bv(int D, const float* __restrict__ baseFn, BvTables B, float4* __restrict__ anc, float4* __restrict__ own, float4* __restrict__ cellD,
                                                  float4* __restrict__ gridLo) {
    {
        const float w = 1.0f / (float)(1 << D);
        const int np = (1 << D) + 1, n = (D + 1) * np;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            const int l = t / np, P = t - l * np, nn = 1 << l;
            const float pos = (float)P * w;
            float v[3];
            for (int k = 0; k < 3; k++) {
                const int ao = (P >> (D - l)) + k - 1;
                v[k] = (ao >= 0 && ao < nn) ? base_value(baseFn, nn - 1 + ao, pos) : 0.f;
            }
            gridLo[t] = make_float4(v[0], v[1], v[2], 0.f);
        }
    }
    {
        const float w = 1.0f / (float)(1 << D);
        const int n = (D + 1) << D;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            const int l = t >> D, gc = t & ((1 << D) - 1), nn = 1 << l;
            const float pos = (float)(gc + 1) * w;
            float v[3];
            for (int k = 0; k < 3; k++) {
                const int ao = (gc >> (D - l)) + k - 1;
                v[k] = (ao >= 0 && ao < nn) ? base_value(baseFn, nn - 1 + ao, pos) : 0.f;
            }
            cellD[t] = make_float4(v[0], v[1], v[2], 0.f);
        }
    }
    for (int d0 = 1; d0 <= D; d0++) {
        const int ng = 1 << (d0 - 1);
        const float w = 1.0f / (float)(1 << d0);
        const int nAnc = d0 * ng * 3;
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nAnc + ng * 3; t += gridDim.x * blockDim.x) {
            if (t < nAnc) {
                const int l = t / (ng * 3), g = (t / 3) % ng, pc = t % 3;
                const int nn = 1 << l, oa = (2 * g) >> (d0 - l);
                const float pp = (float)(2 * g + pc) * w;
                float v[3];
                for (int k = 0; k < 3; k++) {
                    const int ao = oa + k - 1;
                    v[k] = (ao >= 0 && ao < nn) ? base_value(baseFn, nn - 1 + ao, pp) : 0.f;
                }
                anc[B.ancOff[d0] + t] = make_float4(v[0], v[1], v[2], 0.f);
            } else {
                const int u = t - nAnc, g = u / 3, pc = u % 3;
                const int nn = 1 << d0;
                const float pp = (float)(2 * g + pc) * w;
                float v[4];
                for (int cu = 0; cu < 4; cu++) {
                    const int ao = 2 * g + cu - 1;
                    v[cu] = (ao >= 0 && ao < nn) ? base_value(baseFn, nn - 1 + ao, pp) : 0.f;
                }
                own[B.ownOff[d0] + u] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

// base-function values of the three level-l functions around node A at the depth-D grid point P (per axis).
// A point in the closed cell of A comes from a table: the lower-corner one when A == P >> (D-l), the upper-corner
// one of cell P-1 when the point is the upper end of A's cell.  Any other point (the reference's down-walk
// follows childrenVertexKind, which for half of the corners is NOT the child that touches the corner) is
// evaluated in place.  Same base_value() results either way.
__device__ __forceinline__ float4 bv_grid(const float4* __restrict__ gridLo, const float4* __restrict__ cellD, const float* __restrict__ baseFn, int D, int l, int P, int A) {
    const int sh = D - l;
    if ((P >> sh) == A) return gridLo[l * ((1 << D) + 1) + P];
    if (P >= 1 && ((P - 1) >> sh) == A) return cellD[(l << D) + P - 1];
    const int nn = 1 << l;
    const float pos = (float)P * (1.0f / (float)(1 << D));
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ao = A + k - 1;
        v[k] = (ao >= 0 && ao < nn) ? base_value(baseFn, nn - 1 + ao, pos) : 0.f;
    }
    return make_float4(v[0], v[1], v[2], 0.f);
}
// accumulate_level for a point on the depth-D grid: table look-ups instead of nine polynomial evaluations
__device__ __forceinline__ void accumulate_level_grid(float& val, const int* __restrict__ nb, ushort4 o, const float* __restrict__ x,
                                                      const float4* __restrict__ gridLo, const float4* __restrict__ cellD, const float* __restrict__ baseFn, int D,
                                                      const int P[3]) {
    const int d = o.w;
    const float4 bx = bv_grid(gridLo, cellD, baseFn, D, d, P[0], (int)o.x), by = bv_grid(gridLo, cellD, baseFn, D, d, P[1], (int)o.y),
                 bz = bv_grid(gridLo, cellD, baseFn, D, d, P[2], (int)o.z);
    const float vx[3] = {bx.x, bx.y, bx.z}, vy[3] = {by.x, by.y, by.z}, vz[3] = {bz.x, bz.y, bz.z};
#pragma unroll
    for (int j = 0; j < 27; j++) {
        int q = nb[j];
        if (q >= 0) val = __fmaf_rn(__fmul_rn(__fmul_rn(x[q], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], val);
    }
}

// A group owns only ~8 of the 27 points of its corner grid (every vertex belongs to exactly one
// cell), so a warp takes FOUR consecutive groups per step: their owned points are compacted
// (ballot) into a list and handed out one per lane, each lane reading the staged neighbour
// values of its own group's slot.  The 27-term sums -- the bulk of the work -- then run with all
// lanes busy instead of one in four.
constexpr int kVsWarps = 4, kVsChunk = 32, kVsSlots = 4;
constexpr int kVsSlotStride = kMaxDepth * 28 + 8;     // floats; 8 mod 32 apart: the four slots sit in different banks
// The groups to evaluate: up to 16 contiguous ranges (multi-GPU: the replicated depths + this rank's share of every sharded depth) cut
// into chunks of `chunk` groups; the chunks of all ranges form one index space, so ONE launch covers them and a small problem is
// spread over many short chunks instead of a few long sequential ones.
struct VsRanges {
    int n, chunk;
    int first[16], count[16], chunk0[17];
};
__global__ void __launch_bounds__(kVsWarps * 32) k_vertex_values_stream(Topo T, const __grid_constant__ VsRanges RG, int D, const int* __restrict__ parent, const int* __restrict__ child0,
                                                                        const ushort4* __restrict__ offs, const float* __restrict__ x,
                                                                        const float* __restrict__ baseFn, const __grid_constant__ BvTables B, float iso, float* __restrict__ vval) {
    __shared__ __align__(16) float sX[kVsWarps][kVsSlots * kVsSlotStride];   // [slot][level][28]: neighbour values of the slot's cached ancestors
    __shared__ int sAnc[kVsWarps][kVsSlots][kMaxDepth];                       // ... and which nodes those are
    __shared__ float sXc[kVsWarps][kVsSlots][64];                             // solution on the 4x4x4 node cube around each group (own level)
    __shared__ unsigned short sPt[kVsWarps][kVsSlots * 27];                  // owned points: slot | point << 2 | (owner - gb) << 7 | jo << 10
    __shared__ ushort4 sO0[kVsWarps][kVsSlots];
    __shared__ unsigned char sPair[kVsWarps][kVsSlots * kMaxDepth];              // ancestors to refresh this step: slot << 4 | level
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int px = lane / 9, py = (lane / 3) % 3, pz = lane % 3;      // lane -> point of a group's 3x3x3 corner grid (staging)
    const int nChunks = RG.chunk0[RG.n];
    for (int ch = blockIdx.x * kVsWarps + wp; ch < nChunks; ch += gridDim.x * kVsWarps) {
        int rg = 0;
        while (rg + 1 < RG.n && ch >= RG.chunk0[rg + 1]) rg++;
        const int gFirst = RG.first[rg] + (ch - RG.chunk0[rg]) * RG.chunk;
        const int gEnd = min(RG.first[rg] + RG.count[rg], gFirst + RG.chunk);
        int curDepth = -1;
        for (int g0 = gFirst; g0 < gEnd;) {
            // ---- (1) headers of up to four groups: lane q < 4 looks at group g0 + q; the step takes the leading groups of one depth
            ushort4 oq = make_ushort4(0, 0, 0, 0xffff);
            if (lane < kVsSlots && g0 + lane < gEnd) oq = offs[1 + 8 * (g0 + lane)];      // first sibling (root vertices are dropped, main.cu:1634-1638)
            const int d0 = __shfl_sync(0xffffffffu, (int)oq.w, 0);
            const float w = 1.0f / (float)(1 << d0);
            __syncwarp();
            if (d0 != curDepth) {
                for (int t = lane; t < kVsSlots * kMaxDepth; t += 32) sAnc[wp][t / kMaxDepth][t % kMaxDepth] = -1;
                curDepth = d0;
                __syncwarp();
            }
            const unsigned same = __ballot_sync(0xffffffffu, lane < kVsSlots && (int)oq.w == d0);
            const int nq = __ffs(~same) - 1;                  // 1..4: the next depth, if it starts inside, gets its own step
            if (lane < nq) sO0[wp][lane] = oq;
            // ---- (2) ancestors whose node changed since the slot's previous group: the four parent chains are walked in parallel
            // (lane q), every (slot, level) to refresh goes on a list ...
            int nref = 0;
            if (lane < nq) {
                int a = parent[1 + 8 * (g0 + lane)];
                for (int l = d0 - 1; l >= 0 && sAnc[wp][lane][l] != a; --l) { sAnc[wp][lane][l] = a; a = parent[a]; nref++; }
            }
            int pre = nref;
#pragma unroll
            for (int o = 1; o < kVsSlots; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
            const int nPairs = __shfl_sync(0xffffffffu, pre, kVsSlots - 1);
            for (int k = 0; k < nref; k++) sPair[wp][pre - nref + k] = (unsigned char)((lane << 4) | (d0 - 1 - k));
            __syncwarp();
            // ---- (3) ... and the 27 neighbour values of all listed ancestors are gathered in one flat, four-deep batched loop: the
            // latency of the (table -> value) pairs overlaps instead of adding up group after group, level after level
            {
                const int total = nPairs * 27;
                for (int e0 = lane; e0 < total; e0 += 128) {
                    int nb[4], dst[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int e = e0 + 32 * u;
                        nb[u] = -2;
                        if (e < total) {
                            const int pr = sPair[wp][e / 27], j = e % 27, q = pr >> 4, l = pr & 15;
                            dst[u] = q * kVsSlotStride + l * 28 + j;
                            nb[u] = T.nbr[27 * (i64)sAnc[wp][q][l] + j];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (nb[u] != -2) sX[wp][dst[u]] = nb[u] >= 0 ? x[nb[u]] : 0.f;
                }
            }
            // ---- (4) owners of the 27 corner-grid points and the own-level 4x4x4 cube of every group, loads of all groups in flight together
            int np = 0;
            {
                int ownerQ[kVsSlots], joQ[kVsSlots], cubeQ[kVsSlots][2];
#pragma unroll
                for (int q = 0; q < kVsSlots; q++) {
                    ownerQ[q] = -1; joQ[q] = 0; cubeQ[q][0] = cubeQ[q][1] = -1;
                    if (q < nq) {
                        const int gb = 1 + 8 * (g0 + q);
                        if (lane < 27) {
                            const int sx = (px + 1) >> 1, sy = (py + 1) >> 1, sz = (pz + 1) >> 1;
                            const int id = gb + ((sx << 2) | (sy << 1) | sz);
                            const int j = (px - sx) | ((py - sy) << 1) | ((pz - sz) << 2);
                            int m;
                            ownerQ[q] = corner_owner(T, id, j, m);
                            joQ[q] = j ^ m;
                        }
                        // own level: every neighbour of every sibling lies in the 4x4x4 cube around the group
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int e = lane + 32 * h, ux = e >> 4, uy = (e >> 2) & 3, uz = e & 3;
                            const int sx = ux >> 1, sy = uy >> 1, sz = uz >> 1;
                            const int j = 9 * (ux - sx) + 3 * (uy - sy) + (uz - sz);          // 9(dx+1)+3(dy+1)+(dz+1) with d = u - 1 - s
                            cubeQ[q][h] = T.nbr[27 * (i64)(gb + ((sx << 2) | (sy << 1) | sz)) + j];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kVsSlots; q++) {
                    if (q < nq) {
                        sXc[wp][q][lane] = cubeQ[q][0] >= 0 ? x[cubeQ[q][0]] : 0.f;
                        sXc[wp][q][lane + 32] = cubeQ[q][1] >= 0 ? x[cubeQ[q][1]] : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < kVsSlots; q++) {
                    if (q < nq) {                                          // (warp-uniform)
                        const int gb = 1 + 8 * (g0 + q);
                        const bool mine = lane < 27 && ownerQ[q] >= gb && ownerQ[q] < gb + 8;
                        const unsigned mm = __ballot_sync(0xffffffffu, mine);
                        if (mine) sPt[wp][np + __popc(mm & ((1u << lane) - 1u))] = (unsigned short)(q | (lane << 2) | ((ownerQ[q] - gb) << 7) | (joQ[q] << 10));
                        np += __popc(mm);
                    }
                }
            }
            __syncwarp();
            // ---- one owned point per lane
            const int ng = 1 << (d0 - 1);
            const float4* ba = B.anc + B.ancOff[d0];
            for (int p0 = 0; p0 < np; p0 += 32) {
                if (p0 + lane < np) {
                    const int e = sPt[wp][p0 + lane];
                    const int q = e & 3, pt = (e >> 2) & 31, k = (e >> 7) & 7, jo = (e >> 10) & 7;
                    const int qx = pt / 9, qy = (pt / 3) % 3, qz = pt % 3;
                    const ushort4 o0 = sO0[wp][q];
                    const int gb = 1 + 8 * (g0 + q);
                    const int gx = o0.x >> 1, gy = o0.y >> 1, gz = o0.z >> 1;
                    float val = 0.f;
                    {
                        const int sox = (k >> 2) & 1, soy = (k >> 1) & 1, soz = k & 1;
                        const float4 bx = B.own[B.ownOff[d0] + gx * 3 + qx], by = B.own[B.ownOff[d0] + gy * 3 + qy], bz = B.own[B.ownOff[d0] + gz * 3 + qz];
                        const float vx[3] = {sox ? bx.y : bx.x, sox ? bx.z : bx.y, sox ? bx.w : bx.z};
                        const float vy[3] = {soy ? by.y : by.x, soy ? by.z : by.y, soy ? by.w : by.z};
                        const float vz[3] = {soz ? bz.y : bz.x, soz ? bz.z : bz.y, soz ? bz.w : bz.z};
                        const float* xc = &sXc[wp][q][sox * 16 + soy * 4 + soz];
#pragma unroll
                        for (int j = 0; j < 27; j++)
                            val = __fmaf_rn(__fmul_rn(__fmul_rn(xc[(j / 9) * 16 + ((j / 3) % 3) * 4 + (j % 3)], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], val);
                    }
                    // shared ancestor levels d0-1 .. 0
                    const float4* sx4 = reinterpret_cast<const float4*>(&sX[wp][q * kVsSlotStride]) + 7 * (d0 - 1);      // 28 floats per level
                    const float4 *pbx = ba + ((d0 - 1) * ng + gx) * 3 + qx, *pby = ba + ((d0 - 1) * ng + gy) * 3 + qy, *pbz = ba + ((d0 - 1) * ng + gz) * 3 + qz;
                    for (int l = d0 - 1; l >= 0; --l, sx4 -= 7, pbx -= 3 * ng, pby -= 3 * ng, pbz -= 3 * ng) {
                        const float4 bx = *pbx, by = *pby, bz = *pbz;
                        const float vx[3] = {bx.x, bx.y, bx.z}, vy[3] = {by.x, by.y, by.z}, vz[3] = {bz.x, bz.y, bz.z};
                        float X[28];
#pragma unroll
                        for (int t = 0; t < 7; t++) { const float4 v4 = sx4[t]; X[4 * t] = v4.x; X[4 * t + 1] = v4.y; X[4 * t + 2] = v4.z; X[4 * t + 3] = v4.w; }
                        RV_ACC27(val, X, vx, vy, vz);
                    }
                    // finer nodes at this corner (vertices owned above depth D)
                    const int owner = gb + k;
                    if (d0 < D) {
                        const float pos[3] = {(float)((int)o0.x + qx) * w, (float)((int)o0.y + qy) * w, (float)((int)o0.z + qz) * w};
                        int now = owner, depth = d0;
                        const int ex = jo ^ ((jo >> 1) & 1);      // childrenVertexKind {0,1,3,2,4,5,7,6}, MarchingCubes.cuh:721-723 (applied as the reference does)
                        while (depth < D) {
                            ++depth;
                            int c0 = child0[now];
                            if (c0 < 0) break;
                            now = c0 + ex;
                            accumulate_level(val, T.nbr + 27 * (i64)now, offs[now], x, baseFn, pos);      // (the down-walk leaves the corner for half of the corners: no table)
                        }
                    }
                    vval[8 * (i64)owner + jo] = __fsub_rn(val, iso);
                }
            }
            g0 += nq;
        }
    }
}

// ------------------------------------------------------------------ classification of depth-D cells
// per cell: MC case, triangle count, mask of OWNED crossed edges (v1*v2 <= 0, main.cu:2475)
__global__ void __launch_bounds__(128) k_classify(Topo T, const __grid_constant__ ValView W, unsigned char* __restrict__ cat,
                                                  int* __restrict__ ntri, unsigned short* __restrict__ emask, int* __restrict__ nvtx) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < T.nCells; l += gridDim.x * blockDim.x) {
        int id = T.cellBase + l;
        float v[8];
        cell_corner_values(T, id, W, v);
        int c = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) if (v[r] < 0.f) c |= 1 << r;
        unsigned m = 0;
#pragma unroll
        for (int e = 0; e < 12; e++) {
            if (__fmul_rn(v[cEdgeVertex[e][0]], v[cEdgeVertex[e][1]]) <= 0.f) {
                int e2;
                if (edge_owner(T, id, e, e2) == id) m |= 1u << e;
            }
        }
        cat[l] = (unsigned char)c;
        ntri[l] = cMcCount[c];
        emask[l] = (unsigned short)m;
        nvtx[l] = __popc(m);
    }
}
// interpolated vertices (main.cu:2584-2627), written at vbase[cell] + rank of the edge inside the mask
__global__ void __launch_bounds__(128) k_emit_vertices(Topo T, const __grid_constant__ ValView W, const ushort4* __restrict__ cellOffs, int D,
                                                       const unsigned short* __restrict__ emask, const int* __restrict__ vbase, float* __restrict__ outV) {
    const float w = 1.0f / (float)(1 << D);
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < T.nCells; l += gridDim.x * blockDim.x) {
        unsigned m = emask[l];
        if (!m) continue;
        int id = T.cellBase + l;
        float v[8];
        cell_corner_values(T, id, W, v);
        ushort4 o = cellOffs[l];
        int k = 0;
        for (int e = 0; e < 12; e++) {
            if (!(m & (1u << e))) continue;
            int r1 = cEdgeVertex[e][0], r2 = cEdgeVertex[e][1], dim = e >> 2;
            int b1 = ring_to_bits(r1), b2 = ring_to_bits(r2);
            float p1[3] = {(float)((int)o.x + (b1 & 1)) * w, (float)((int)o.y + ((b1 >> 1) & 1)) * w, (float)((int)o.z + ((b1 >> 2) & 1)) * w};
            float p2d = (float)((dim == 0 ? (int)o.x + (b2 & 1) : (dim == 1 ? (int)o.y + ((b2 >> 1) & 1) : (int)o.z + ((b2 >> 2) & 1)))) * w;
            float f1 = v[r1], f2 = v[r2];
            float pivot = __fdiv_rn(f1, __fsub_rn(f1, f2));
            float another = __fsub_rn(1.0f, pivot);
            float out[3] = {p1[0], p1[1], p1[2]};
            out[dim] = __fmaf_rn(p2d, pivot, __fmul_rn(p1[dim], another));
            i64 a = 3 * (i64)(vbase[l] + k);
            outV[a] = out[0]; outV[a + 1] = out[1]; outV[a + 2] = out[2];
            k++;
        }
    }
}
// triangles (main.cu:2699-2757) + marking of faces touched by the surface and of their parent faces.  A vertex id is
// (first vertex of the owner cell's rank) + (the owner's offset inside that rank) + (rank of the edge inside the owner's mask);
// multi-GPU: the owner of an edge on a shard boundary belongs to the previous rank, whose offset / mask arrays are read in place.
struct TriView {
    const int* vbase[kMaxRanks];             // [id - idxBase] per rank
    const unsigned short* emask[kMaxRanks];
    int vtxBase[kMaxRanks];                  // first vertex id of every rank's cells
    int idxBase, world;
    int lo[kMaxRanks + 1];
};
__global__ void __launch_bounds__(128) k_emit_triangles(Topo T, const unsigned char* __restrict__ cat, const int* __restrict__ ntri, const int* __restrict__ tbase,
                                                        const __grid_constant__ TriView V, int* __restrict__ outT,
                                                        int markFaces, int nUpper, const int* __restrict__ parent, const ushort4* __restrict__ offs,
                                                        unsigned* __restrict__ fmark) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < T.nCells; l += gridDim.x * blockDim.x) {
        int nt = ntri[l];
        if (!nt) continue;
        int id = T.cellBase + l, c = cat[l];
        unsigned used = 0;
        for (int j = 0; j < 3 * nt; j++) {
            int e = cMcTri[c][j], e2;
            used |= 1u << e;
            int ow = edge_owner(T, id, e, e2);
            const int r = V.world > 1 ? view_rank(V.lo, V.world, ow) : 0;
            const int ol = ow - V.idxBase;
            outT[3 * (i64)tbase[l] + j] = V.vtxBase[r] + V.vbase[r][ol] + __popc((unsigned)V.emask[r][ol] & ((1u << e2) - 1u));
        }
        if (!markFaces) continue;
        for (int f = 0; f < 6; f++) {
            unsigned fm = (unsigned)cFaceEdgeMask[f][0] | ((unsigned)cFaceEdgeMask[f][1] << 8);
            if (!(used & fm)) continue;
            // face (node, f) and the chain of parent faces (main.cu:2738-2755).  A face is shared
            // by the two cells across it; its hasParentFace flag was set by its OWNER (min index)
            // with the reference's parentFaceKind table (MarchingCubes.cuh:708-717: child code
            // read as x=bit0, literal row 7).  Only nodes above depth D are ever tested (k_find_subdivide).
            int node = id;
            int axis = f >> 1, dd[3] = {0, 0, 0};
            dd[axis] = (f & 1) ? 1 : -1;
            int jn = 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1);
            while (true) {
                int across = T.nbr[27 * (i64)node + jn];
                if (node < nUpper) atomicOr(&fmark[node], 1u << f);
                if (across >= 0 && across < nUpper) atomicOr(&fmark[across], 1u << (f ^ 1));
                int owner = (across >= 0 && across < node) ? across : node;
                int fo = (owner == node) ? f : (f ^ 1);
                int pa = parent[owner];
                if (pa < 0) break;
                ushort4 oo = offs[owner];
                int son = (((int)oo.x & 1) << 2) | (((int)oo.y & 1) << 1) | ((int)oo.z & 1);
                int pk = (((son >> (fo >> 1)) & 1) == (fo & 1)) ? fo : -1;
                if (son == 7 && fo == 0) pk = 0;
                if (pk == -1) break;
                node = parent[node];
                if (node < 0) break;
            }
        }
    }
}
// empty leaves below depth D that must be refined (main.cu:2957-2992)
struct MarkView { const unsigned* p[kMaxRanks]; int world; };      // face marks of every rank's own cells (multi-GPU: OR of the peers' arrays)
// corner values of ALL depths above D: multi-GPU, the node ranges rowLo[d][.] of the sharded depths live on their ranks (read in place)
struct UpperView {
    const float* p[kMaxRanks];
    int world, shardFrom;
    int rowLo[kMaxDepth + 1][kMaxRanks + 1];
};
__global__ void __launch_bounds__(128) k_find_subdivide(Topo T, int nNodes, const int* __restrict__ child0, const ushort4* __restrict__ offs, const __grid_constant__ UpperView U,
                                                        const __grid_constant__ MarkView F, int* __restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += gridDim.x * blockDim.x) {
        int f = 0;
        if (i > 0 && child0[i] < 0) {
            float v[8];
            const int d = offs[i].w;
            const bool remote = U.world > 1 && d >= U.shardFrom;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int j = ring_to_bits(r), m;
                const int ow = corner_owner(T, i, j, m);              // (same depth as i)
                const int rk = remote ? view_rank(U.rowLo[d], U.world, ow) : 0;
                v[r] = U.p[rk][8 * (i64)ow + (j ^ m)];
            }
            int sign = (v[0] < 0.f) ? -1 : 1;
            int ht = 0;
#pragma unroll
            for (int r = 1; r < 8; r++) if ((float)sign * v[r] < 0.f) ht = 1;
            unsigned fm = 0u;
            for (int r = 0; r < F.world; r++) fm |= F.p[r][i];
            f = (ht || fm != 0u) ? 1 : 0;
        }
        flag[i] = f;
    }
}
__global__ void __launch_bounds__(256) k_compact_ids(const int* __restrict__ flag, const int* __restrict__ excl, int n, int* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) out[excl[i]] = i;
}

// ------------------------------------------------------------------ refinement: virtual complete subtrees
struct VTree {
    int M, D, rd, nr;
    int depthAddr[kMaxDepth + 2];   // first virtual index of each level (valid for rd..D)
    const int* roots;               // real node ids of the pass, ascending
};
__device__ __forceinline__ int vt_per(const VTree& V, int d) { return 1 << (3 * (d - V.rd)); }

// neighbours of the virtual nodes of level d from their parents' (computeRebuildNeighbor, main.cu:3188-3217)
__global__ void __launch_bounds__(256) k_vneigh(VTree V, int d, const int* __restrict__ neighs, const int* __restrict__ parent, const int* __restrict__ child0,
                                                const ushort4* __restrict__ offs, const int* __restrict__ rootMap, int* __restrict__ vneigh) {
    const int per = vt_per(V, d);
    const i64 total = (i64)V.nr * per * 27;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        int vl = (int)(t / 27), j = (int)(t - (i64)vl * 27);
        int r = vl / per, l = vl - r * per;
        int c, np, pj, cc;
        if (d == V.rd) {
            int root = V.roots[r];
            ushort4 o = offs[root];
            c = (((int)o.x & 1) << 2) | (((int)o.y & 1) << 1) | ((int)o.z & 1);
            lut_parent_child(c, j, pj, cc);
            np = neighs[27 * (i64)parent[root] + pj];
        } else {
            c = l & 7;
            lut_parent_child(c, j, pj, cc);
            int pv = V.depthAddr[d - 1] + r * (per >> 3) + (l >> 3);
            np = vneigh[27 * (i64)pv + pj];
        }
        int out = -1;
        if (np >= 0) {
            if (np < V.M) {
                int c0 = child0[np];
                if (c0 >= 0) { int ch = c0 + cc; int vm = rootMap[ch]; out = vm >= 0 ? V.M + vm : ch; }
            } else {
                int pvv = np - V.M - V.depthAddr[d - 1];
                int r2 = pvv / (per >> 3), l2 = pvv - r2 * (per >> 3);
                out = V.M + V.depthAddr[d] + r2 * per + (l2 << 3) + cc;
            }
        }
        vneigh[27 * (i64)(V.depthAddr[d] + vl) + j] = out;
    }
}
__global__ void k_set_rootmap(const int* __restrict__ roots, int nr, int first, int value, int* __restrict__ rootMap) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nr; r += gridDim.x * blockDim.x) rootMap[roots[r]] = value < 0 ? -1 : first + r;
}
// offsets of the depth-D virtual cells: root offsets extended by the octal digits of l
__global__ void __launch_bounds__(256) k_vcell_offsets(VTree V, const ushort4* __restrict__ offs, ushort4* __restrict__ voffs) {
    const int per = vt_per(V, V.D), lv = V.D - V.rd;
    const int total = V.nr * per;
    for (int vl = blockIdx.x * blockDim.x + threadIdx.x; vl < total; vl += gridDim.x * blockDim.x) {
        int r = vl / per, l = vl - r * per;
        ushort4 o = offs[V.roots[r]];
        int ox = o.x, oy = o.y, oz = o.z;
        for (int s = lv - 1; s >= 0; --s) {
            int c = (l >> (3 * s)) & 7;
            ox = (ox << 1) | ((c >> 2) & 1);
            oy = (oy << 1) | ((c >> 1) & 1);
            oz = (oz << 1) | (c & 1);
        }
        voffs[vl] = make_ushort4((unsigned short)ox, (unsigned short)oy, (unsigned short)oz, (unsigned short)V.D);
    }
}
// per virtual node: which of its 27 neighbour slots carry a solution value (a real node, or, at
// the roots' level, a virtual root standing for the real leaf it replaces)
__global__ void __launch_bounds__(256) k_vmask(VTree V, i64 total, const int* __restrict__ vneigh, unsigned* __restrict__ vmask) {
    for (i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (i64)gridDim.x * blockDim.x) {
        bool rootLevel = v < V.depthAddr[V.rd] + V.nr;
        unsigned m = 0;
        const int* nb = vneigh + 27 * v;
#pragma unroll
        for (int j = 0; j < 27; j++) {
            int q = nb[j];
            if (q >= 0 && (q < V.M || rootLevel)) m |= 1u << j;
        }
        vmask[v] = m;
    }
}
// corner values of the depth-D virtual cells: only REAL nodes carry a solution; a virtual root
// stands for the real leaf it replaces (main.cu:2328-2442).  Two steps: k_vcorner_list -- one thread per (cell, corner): the
// corners a cell owns are appended to a list (warp-aggregated atomic; the order is irrelevant, a value is stored at its
// (cell, corner) slot) -- then k_vvertex_values with one thread per OWNED corner, so that no lane idles while another walks the
// levels of eight corners (one thread per cell: 66 ms for the two passes of the 20 M-point depth-11 scene).
// Virtual levels without any real neighbour are skipped via vmask.
__global__ void __launch_bounds__(256) k_vcorner_list(Topo T, int* __restrict__ list, int* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    for (int l0 = blockIdx.x * blockDim.x; l0 < T.nCells; l0 += gridDim.x * blockDim.x) {
        const int l = l0 + threadIdx.x;
        unsigned own = 0;                               // bit j: the cell owns its corner j
        if (l < T.nCells) {
            const int id = T.cellBase + l;
            int nb[27];                                 // the 27 neighbour ids once (every corner looks at 8 of them)
            const int* row = T.nbr + 27 * (i64)(id - T.rowBase);
#pragma unroll
            for (int j = 0; j < 27; j++) nb[j] = row[j];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int sx = (j & 1) ? 1 : -1, sy = (j & 2) ? 1 : -1, sz = (j & 4) ? 1 : -1;
                int best = 0x7fffffff;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int dx = (q & 1) ? sx : 0, dy = (q & 2) ? sy : 0, dz = (q & 4) ? sz : 0;
                    const int v = nb[9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
                    if (v >= T.minId && v < best) best = v;
                }
                if (best == id) own |= 1u << j;
            }
        }
        // warp-aggregated append
        const int mine = __popc(own);
        int pre = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
        const int tot = __shfl_sync(0xffffffffu, pre, 31);
        int base = 0;
        if (lane == 31 && tot) base = atomicAdd(count, tot);
        base = __shfl_sync(0xffffffffu, base, 31) + pre - mine;
        for (unsigned m = own; m; m &= m - 1) list[base++] = (l << 3) | (__ffs(m) - 1);
    }
}
__global__ void __launch_bounds__(128) k_vvertex_values(VTree V, const int* __restrict__ list, const int* __restrict__ count, const int* __restrict__ vneigh, const unsigned* __restrict__ vmask,
                                                        const ushort4* __restrict__ voffs,
                                                        const float* __restrict__ rootX, const ushort4* __restrict__ offs,
                                                        const float* __restrict__ x, const float4* __restrict__ gridLo, const float4* __restrict__ cellD,
                                                        const float* __restrict__ baseFn, float iso, float* __restrict__ sval) {
    const int perD = vt_per(V, V.D);
    const int n = *count;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int t = list[e], l = t >> 3, j = t & 7;
        const ushort4 o = voffs[l];
        const int r = l / perD;
        const int P[3] = {(int)o.x + (j & 1), (int)o.y + ((j >> 1) & 1), (int)o.z + ((j >> 2) & 1)};     // the corner on the depth-D grid
        float val = 0.f;
        int loc = l - r * perD;
        for (int d = V.D; d >= V.rd; --d) {           // virtual levels D .. rd
            int per = vt_per(V, d);
            int v = V.depthAddr[d] + r * per + loc;
            loc >>= 3;
            unsigned mk = vmask[v];
            if (!mk) continue;
            const int* nb = vneigh + 27 * (i64)v;
            const int od = V.D - d;
            const float4 bx = bv_grid(gridLo, cellD, baseFn, V.D, d, P[0], (int)o.x >> od), by = bv_grid(gridLo, cellD, baseFn, V.D, d, P[1], (int)o.y >> od),
                         bz = bv_grid(gridLo, cellD, baseFn, V.D, d, P[2], (int)o.z >> od);
            const float vx[3] = {bx.x, bx.y, bx.z}, vy[3] = {by.x, by.y, by.z}, vz[3] = {bz.x, bz.y, bz.z};
#pragma unroll
            for (int jj = 0; jj < 27; jj++) {
                if (!(mk & (1u << jj))) continue;
                int q = nb[jj];
                if (q >= V.M) q = V.roots[q - V.M - V.depthAddr[V.rd]];   // virtual root -> the real leaf it replaces
                val = __fmaf_rn(__fmul_rn(__fmul_rn(x[q], vx[jj / 9]), vy[(jj / 3) % 3]), vz[jj % 3], val);
            }
        }
        // real ancestors, levels rd-1 .. 0: the solution at their 27 neighbours comes from the per-root table (k_rv_roots), shared by all
        // the corners of the root, instead of 27 table look-ups + 27 gathers per level and corner
        const ushort4 ro = offs[V.roots[r]];
        const float* RX = rootX + (i64)r * (V.rd + 1) * 27;
        for (int lv = V.rd - 1; lv >= 0; --lv) {
            const int sh = V.rd - lv;
            const float4 bx = bv_grid(gridLo, cellD, baseFn, V.D, lv, P[0], (int)ro.x >> sh), by = bv_grid(gridLo, cellD, baseFn, V.D, lv, P[1], (int)ro.y >> sh),
                         bz = bv_grid(gridLo, cellD, baseFn, V.D, lv, P[2], (int)ro.z >> sh);
            const float vx[3] = {bx.x, bx.y, bx.z}, vy[3] = {by.x, by.y, by.z}, vz[3] = {bz.x, bz.y, bz.z};
            RV_ACC27(val, RX + lv * 27, vx, vy, vz);
        }
        sval[8 * (i64)l + j] = __fsub_rn(val, iso);
    }
}
__global__ void __launch_bounds__(256) k_offset_triangles(int* __restrict__ t, i64 n, int off) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) t[i] += off;
}

// ================================================================== refinement, implicit form
// Passes whose roots lie at least 3 levels above D (n = 2^(D-rd) >= 8 cells per side, up to
// 8^(D-rd) cells per root) never materialise the virtual subtree: a virtual depth-D cell is
// (root r, local Morton code l), its neighbours, the owners of its corners and edges and the
// real nodes next to it all follow from arithmetic on (r, l) plus one 27-entry table per ROOT.
// Values: the corner (+,+,+) of a cell is always owned by that cell (it is the lowest of the 8
// incident cells in Morton = id order), so one float per cell (val7) holds every grid value
// except those on the lower faces of a root without a virtual neighbour there (kept in `low`).
// A block evaluates one 8x8x8 brick: the real nodes next to the brick's ancestors are found once
// per brick by descending the real tree, the per-axis base-function values once per
// coordinate, and every cell then sums its levels D..0 in the reference's order
// (main.cu:2328-2442) from shared memory.
struct RGeom {
    int M, D, rd, lv, n, nr;
    unsigned per;                 // n^3
    const int* roots;             // [nr] real node ids of the pass, ascending
    const int* rootNb;            // [nr][27] depth-rd neighbours: real id, M + r' (virtual root r') or -1
    const float* rootX;           // [nr][rd+1][27] solution at the 27 neighbours of the root (level rd) and of its ancestors (0 = absent)
    const ushort4* offs;
    const int* child0;
    const float* x;
    const float* baseFn;
    float iso;
    const float4* bvCell;         // BvTables::cellD
    float* val7;                  // corner-7 values, 512 per STORED brick: slot[brick] * 512 + (cell & 511)
    float* low;                   // [nr][3][(n+1)^2]
    const int* slot;              // brick -> storage slot (its position among the bricks that are evaluated); null: every brick is stored
};
// storage index of virtual cell `cell` (= brick * 512 + position) in the per-brick arrays val7 / cat / emask / vpre: only the bricks
// that are evaluated -- a few per cent of a pass -- are stored
__device__ __forceinline__ i64 rv_store(const RGeom& G, i64 cell) {
    return G.slot ? (((i64)G.slot[cell >> 9]) << 9) | (cell & 511) : cell;
}

static int ensure_bv_tables(Context& c);

__device__ __forceinline__ unsigned spread3(unsigned v) {   // bit s -> bit 3s (10 bits)
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned compact3(unsigned v) {  // bit 3s -> bit s
    v &= 0x09249249u;
    v = (v ^ (v >> 2)) & 0x030C30C3u;
    v = (v ^ (v >> 4)) & 0x0300F00Fu;
    v = (v ^ (v >> 8)) & 0x030000FFu;
    v = (v ^ (v >> 16)) & 0x000003FFu;
    return v;
}
// local Morton code of a cell: octal digit per level, digit = x<<2|y<<1|z (child code)
__device__ __forceinline__ unsigned rv_morton(int x, int y, int z) { return (spread3((unsigned)x) << 2) | (spread3((unsigned)y) << 1) | spread3((unsigned)z); }

__global__ void __launch_bounds__(256) k_rv_roots(int nr, int rd, int M, const int* __restrict__ roots, const int* __restrict__ rootMap,
                                                  const int* __restrict__ neighs, const int* __restrict__ parent, const float* __restrict__ x,
                                                  int* __restrict__ rootNb, float* __restrict__ rootX) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nr || lane >= 27) return;
    int root = roots[w];
    int q = neighs[27 * (i64)root + lane];
    int nb = -1;
    float xv = 0.f;
    if (q >= 0) { int vm = rootMap[q]; nb = vm >= 0 ? M + vm : q; xv = x[q]; }   // a virtual root stands for the real leaf it replaces
    rootNb[27 * w + lane] = nb;
    float* X = rootX + (i64)w * (rd + 1) * 27;
    X[rd * 27 + lane] = xv;
    int now = parent[root];
    for (int l = rd - 1; l >= 0; --l) {
        int q2 = neighs[27 * (i64)now + lane];
        X[l * 27 + lane] = q2 >= 0 ? x[q2] : 0.f;
        now = parent[now];
    }
}


// Values at the upper corner (corner 7) of every virtual depth-D cell, one CTA of 64 threads per
// brick of 8x8x8 cells.  A thread owns a z COLUMN of 8 cells: for every level and neighbour slot
// j the product (x_j * Bx) * By is formed once and feeds the 8 cells' FMAs with their own Bz --
// 1.25 instead of 3 instructions per term and cell, with every value computed exactly as in the
// one-thread-per-cell formulation (same operations in the same order per cell).  Base-function
// values come from the per-context table BvTables::cellD.  The levels BELOW the brick level (real nodes next to the brick) come from
// three dense windows staged once per brick, see below.
__global__ void __launch_bounds__(64) k_rv_brick_values(RGeom G, const int* __restrict__ list, const int* __restrict__ listCount, int nAll) {
    __shared__ __align__(16) float sX[kMaxDepth + 1][28];
    __shared__ int sIds[27];
    __shared__ int sAny[kMaxDepth + 1];
    __shared__ int sNeedFine;
    // the real nodes of the three levels below the brick level that can reach its cells, as dense windows around the brick
    // (level D-2: 4^3, D-1: 6^3, D: 10^3 nodes; -1 / 0 where the real tree has nothing): ids of the first two, solution of all three
    __shared__ int sId1[64], sId2[216];
    __shared__ float sF1[64], sF2[216], sF3[1000];
    const int tid = threadIdx.x;
    // the list length stays on the device (no host round trip between the brick selection and the evaluation): persistent CTAs
    const int nWork = list ? *listCount : nAll;
    for (int work = blockIdx.x; work < nWork; work += gridDim.x) {
    __syncthreads();
    const i64 cell0 = (i64)(list ? list[work] : work) * 512;
    const int r = (int)(cell0 / G.per);
    const unsigned l0 = (unsigned)(cell0 - (i64)r * G.per);
    const int L = G.D - 3;                                   // level of the brick
    const ushort4 ro = G.offs[G.roots[r]];
    const int bx = ((int)ro.x << G.lv) + (int)compact3(l0 >> 2), by = ((int)ro.y << G.lv) + (int)compact3(l0 >> 1), bz = ((int)ro.z << G.lv) + (int)compact3(l0);
    for (int t = tid; t < (G.rd + 1) * 27; t += 64) sX[t / 27][t % 27] = G.rootX[(i64)r * (G.rd + 1) * 27 + t];
    if (tid <= G.rd) sAny[tid] = 1;   // levels rd+1..L are written by the descending warp below
    if (tid >= 32) {
        // one warp descends the REAL tree along the brick's path, levels rd+1 .. L: the 27
        // neighbours of the brick's ancestor at every level (virtual ones carry no solution)
        const int lane = tid - 32;
        int cur = -1;
        if (lane < 27) { int q = G.rootNb[27 * r + lane]; cur = (q >= 0 && q < G.M) ? q : -1; }
        for (int d = G.rd + 1; d <= L; d++) {
            int c = (int)((l0 >> (3 * (G.D - d))) & 7u), pj = 0, cc = 0;
            if (lane < 27) lut_parent_child(c, lane, pj, cc);
            int p = __shfl_sync(0xffffffffu, cur, pj);
            int nxt = -1;
            if (lane < 27 && p >= 0) { int c0 = G.child0[p]; if (c0 >= 0) nxt = c0 + cc; }
            unsigned any = __ballot_sync(0xffffffffu, nxt >= 0);
            if (lane < 27) sX[d][lane] = nxt >= 0 ? G.x[nxt] : 0.f;
            if (lane == 0) sAny[d] = any != 0u;
            cur = nxt;
        }
        bool kids = lane < 27 && cur >= 0 && G.child0[cur] >= 0;
        unsigned anyKids = __ballot_sync(0xffffffffu, kids);
        if (lane < 27) sIds[lane] = kids ? cur : -1;
        if (lane == 0) sNeedFine = anyKids != 0u;
    }
    __syncthreads();
    const int cx = tid >> 3, cy = tid & 7;
    const int gx = bx + cx, gy = by + cy;
    const unsigned lxy = l0 + (spread3((unsigned)cx) << 2) + (spread3((unsigned)cy) << 1);
    float val[8];
#pragma unroll
    for (int cz = 0; cz < 8; cz++) val[cz] = 0.f;
    if (sNeedFine) {
        // ---- levels D-2, D-1, D: descend the real tree ONCE per brick into the three windows (window coordinate u = node offset
        // from the brick origin at that level + 1; its parent sits at (u + 1) >> 1 of the coarser window, child bit (u + 1) & 1), ...
        for (int t = tid; t < 64; t += 64) {
            const int ux = t >> 4, uy = (t >> 2) & 3, uz = t & 3;
            const int p = sIds[9 * ((ux + 1) >> 1) + 3 * ((uy + 1) >> 1) + ((uz + 1) >> 1)];
            const int id = p >= 0 ? G.child0[p] + ((((ux + 1) & 1) << 2) | (((uy + 1) & 1) << 1) | ((uz + 1) & 1)) : -1;      // (sIds holds nodes WITH children only)
            sId1[t] = id;
            sF1[t] = id >= 0 ? G.x[id] : 0.f;
        }
        __syncthreads();
        for (int t = tid; t < 216; t += 64) {
            const int ux = t / 36, uy = (t / 6) % 6, uz = t % 6;
            const int p = sId1[16 * ((ux + 1) >> 1) + 4 * ((uy + 1) >> 1) + ((uz + 1) >> 1)];
            int id = -1;
            if (p >= 0) { const int c0 = G.child0[p]; if (c0 >= 0) id = c0 + ((((ux + 1) & 1) << 2) | (((uy + 1) & 1) << 1) | ((uz + 1) & 1)); }
            sId2[t] = id;
            sF2[t] = id >= 0 ? G.x[id] : 0.f;
        }
        __syncthreads();
        for (int t = tid; t < 1000; t += 64) {
            const int ux = t / 100, uy = (t / 10) % 10, uz = t % 10;
            const int p = sId2[36 * ((ux + 1) >> 1) + 6 * ((uy + 1) >> 1) + ((uz + 1) >> 1)];
            float xv = 0.f;
            if (p >= 0) { const int c0 = G.child0[p]; if (c0 >= 0) xv = G.x[c0 + ((((ux + 1) & 1) << 2) | (((uy + 1) & 1) << 1) | ((uz + 1) & 1))]; }
            sF3[t] = xv;
        }
        __syncthreads();
        // ---- ... then every cell sums its 27 neighbours of level D, D-1, D-2 (the reference's order; a missing node adds an exact 0)
        // from shared memory: 2 300 loads per brick instead of 83 000 through per-cell descents
        const float4* tD = G.bvCell + ((L + 3) << G.D);
        const float4* tD1 = G.bvCell + ((L + 2) << G.D);
        const float4* tD2 = G.bvCell + ((L + 1) << G.D);
        const float4 fx3 = tD[gx], fy3 = tD[gy], fx2 = tD1[gx], fy2 = tD1[gy], fx1 = tD2[gx], fy1 = tD2[gy];
#pragma unroll 1
        for (int cz = 0; cz < 8; cz++) {
            const float4 fz3 = tD[bz + cz], fz2 = tD1[bz + cz], fz1 = tD2[bz + cz];
            float v = 0.f;
            {
                const float vx[3] = {fx3.x, fx3.y, fx3.z}, vy[3] = {fy3.x, fy3.y, fy3.z}, vz[3] = {fz3.x, fz3.y, fz3.z};
                const float* X = sF3 + 100 * cx + 10 * cy + cz;                 // window coordinate of neighbour d = cell + d + 1
#pragma unroll
                for (int j = 0; j < 27; j++) v = __fmaf_rn(__fmul_rn(__fmul_rn(X[100 * (j / 9) + 10 * ((j / 3) % 3) + (j % 3)], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], v);
            }
            {
                const float vx[3] = {fx2.x, fx2.y, fx2.z}, vy[3] = {fy2.x, fy2.y, fy2.z}, vz[3] = {fz2.x, fz2.y, fz2.z};
                const float* X = sF2 + 36 * (cx >> 1) + 6 * (cy >> 1) + (cz >> 1);
#pragma unroll
                for (int j = 0; j < 27; j++) v = __fmaf_rn(__fmul_rn(__fmul_rn(X[36 * (j / 9) + 6 * ((j / 3) % 3) + (j % 3)], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], v);
            }
            {
                const float vx[3] = {fx1.x, fx1.y, fx1.z}, vy[3] = {fy1.x, fy1.y, fy1.z}, vz[3] = {fz1.x, fz1.y, fz1.z};
                const float* X = sF1 + 16 * (cx >> 2) + 4 * (cy >> 2) + (cz >> 2);
#pragma unroll
                for (int j = 0; j < 27; j++) v = __fmaf_rn(__fmul_rn(__fmul_rn(X[16 * (j / 9) + 4 * ((j / 3) % 3) + (j % 3)], vx[j / 9]), vy[(j / 3) % 3]), vz[j % 3], v);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) if (k == cz) val[k] = v;
        }
    }
    for (int lvl = L; lvl >= 0; --lvl) {
        if (!sAny[lvl]) continue;
        const float4* tb = G.bvCell + (lvl << G.D);
        const float4 fx = tb[gx], fy = tb[gy];
        const float vx[3] = {fx.x, fx.y, fx.z}, vy[3] = {fy.x, fy.y, fy.z};
        float vz[3][8];
#pragma unroll
        for (int cz = 0; cz < 8; cz++) {
            const float4 fz = tb[bz + cz];
            vz[0][cz] = fz.x; vz[1][cz] = fz.y; vz[2][cz] = fz.z;
        }
        float X[28];
#pragma unroll
        for (int q = 0; q < 7; q++) {
            const float4 xv = *reinterpret_cast<const float4*>(&sX[lvl][4 * q]);
            X[4 * q] = xv.x; X[4 * q + 1] = xv.y; X[4 * q + 2] = xv.z; X[4 * q + 3] = xv.w;
        }
#pragma unroll
        for (int j = 0; j < 27; j++) {
            const float t = __fmul_rn(__fmul_rn(X[j], vx[j / 9]), vy[(j / 3) % 3]);
#pragma unroll
            for (int cz = 0; cz < 8; cz++) val[cz] = __fmaf_rn(t, vz[j % 3][cz], val[cz]);
        }
    }
#pragma unroll
    for (int cz = 0; cz < 8; cz++) G.val7[rv_store(G, cell0) + (lxy - l0) + spread3((unsigned)cz)] = __fsub_rn(val[cz], G.iso);
    }
}

// virtual cell at root-local coordinates (x,y,z), each in [-1, n]: which root of the pass holds
// it (false: the cell is not virtual), coordinates wrapped into that root
__device__ __forceinline__ bool rv_locate(const RGeom& G, int r, int& x, int& y, int& z, int& r2) {
    int dx = x < 0 ? -1 : (x >= G.n ? 1 : 0), dy = y < 0 ? -1 : (y >= G.n ? 1 : 0), dz = z < 0 ? -1 : (z >= G.n ? 1 : 0);
    if ((dx | dy | dz) == 0) { r2 = r; return true; }
    int q = G.rootNb[27 * r + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
    if (q < G.M) return false;
    r2 = q - G.M;
    x -= dx * G.n; y -= dy * G.n; z -= dz * G.n;
    return true;
}
// owner of the grid point g (root-local, each in [0,n]) of root r: the incident virtual cell with
// the smallest id = (root, Morton) (main.cu:1474-1484 "min key"); jb = corner bits of g in that cell
__device__ __forceinline__ void rv_point_owner(const RGeom& G, int r, int gx, int gy, int gz, int& r2, int& ox, int& oy, int& oz, int& jb) {
    unsigned long long best = ~0ull;
    r2 = -1;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int x = gx - 1 + (q & 1), y = gy - 1 + ((q >> 1) & 1), z = gz - 1 + ((q >> 2) & 1), rr;
        if (!rv_locate(G, r, x, y, z, rr)) continue;
        unsigned long long key = ((unsigned long long)rr << 32) | rv_morton(x, y, z);
        if (key < best) { best = key; r2 = rr; ox = x; oy = y; oz = z; jb = 7 ^ q; }
    }
}
__device__ __forceinline__ i64 rv_low_index(const RGeom& G, int r, int gx, int gy, int gz) {
    int a = gx == 0 ? 0 : (gy == 0 ? 1 : 2);
    int u = a == 0 ? gy : gx, v = a == 2 ? gy : gz;
    const int n1 = G.n + 1;
    return ((i64)r * 3 + a) * n1 * n1 + (i64)u * n1 + v;
}
// value at grid point g of root r
__device__ __forceinline__ float rv_point_value(const RGeom& G, int r, int gx, int gy, int gz) {
    if (gx >= 1 && gy >= 1 && gz >= 1) return G.val7[rv_store(G, (i64)r * G.per + rv_morton(gx - 1, gy - 1, gz - 1))];
    int r2, ox, oy, oz, jb;
    rv_point_owner(G, r, gx, gy, gz, r2, ox, oy, oz, jb);
    if (jb == 7) return G.val7[rv_store(G, (i64)r2 * G.per + rv_morton(ox, oy, oz))];
    return G.low[rv_low_index(G, r2, ox + (jb & 1), oy + ((jb >> 1) & 1), oz + ((jb >> 2) & 1))];
}
// corner jb of virtual cell (r, c) evaluated by one thread (lower faces of the pass region only)
__device__ float rv_eval_slow(const RGeom& G, int r, int cx, int cy, int cz, int jb) {
    const ushort4 ro = G.offs[G.roots[r]];
    const int g[3] = {((int)ro.x << G.lv) + cx, ((int)ro.y << G.lv) + cy, ((int)ro.z << G.lv) + cz};
    const float w = 1.0f / (float)(1 << G.D);
    const float pos[3] = {(float)(g[0] + (jb & 1)) * w, (float)(g[1] + ((jb >> 1) & 1)) * w, (float)(g[2] + ((jb >> 2) & 1)) * w};
    const unsigned l = rv_morton(cx, cy, cz);
    int ids[kMaxDepth][27];
    int deepest = G.rd;
    {
        int cur[27];
        for (int j = 0; j < 27; j++) { int q = G.rootNb[27 * r + j]; cur[j] = (q >= 0 && q < G.M) ? q : -1; }
        for (int d = G.rd + 1; d <= G.D; d++) {
            int c = (int)((l >> (3 * (G.D - d))) & 7u);
            bool any = false;
            for (int j = 0; j < 27; j++) {
                int pj, cc;
                lut_parent_child(c, j, pj, cc);
                int p = cur[pj], nxt = -1;
                if (p >= 0) { int c0 = G.child0[p]; if (c0 >= 0) nxt = c0 + cc; }
                ids[d - G.rd - 1][j] = nxt;
                any |= nxt >= 0;
            }
            if (!any) break;
            deepest = d;
            for (int j = 0; j < 27; j++) cur[j] = ids[d - G.rd - 1][j];
        }
    }
    float val = 0.f;
    for (int d = deepest; d >= 0; --d) {
        const int nn = 1 << d;
        float v[3][3];
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < 3; k++) {
                int ao = (g[a] >> (G.D - d)) + k - 1;
                v[a][k] = (ao >= 0 && ao < nn) ? base_value(G.baseFn, nn - 1 + ao, pos[a]) : 0.f;
            }
        if (d > G.rd) {
            for (int j = 0; j < 27; j++) {
                int q = ids[d - G.rd - 1][j];
                if (q >= 0) val = __fmaf_rn(__fmul_rn(__fmul_rn(G.x[q], v[0][j / 9]), v[1][(j / 3) % 3]), v[2][j % 3], val);
            }
        } else {
            const float* X = G.rootX + ((i64)r * (G.rd + 1) + d) * 27;
            for (int j = 0; j < 27; j++) val = __fmaf_rn(__fmul_rn(__fmul_rn(X[j], v[0][j / 9]), v[1][(j / 3) % 3]), v[2][j % 3], val);
        }
    }
    return __fsub_rn(val, G.iso);
}
// grid points on the lower faces of every root: evaluated here when this root owns them.  Tiles of 128 points of one root, 32-bit
// index arithmetic; a point nobody reads is dropped BEFORE the owner search (8 cell look-ups across the neighbouring roots): its
// owner, if it lies in this root at all, is one of the (at most 8) cells of this root around the point, and only bricks staged
// for classification (needLow) read lower-face values.
__global__ void __launch_bounds__(128) k_rv_low_values(RGeom G, const unsigned char* __restrict__ needed /* per brick (k_rv_brick_needed needLow), or null: all */) {
    const int n1 = G.n + 1, nn = n1 * n1, per = 3 * nn;
    const unsigned tilesPerRoot = (unsigned)((per + 127) / 128);
    const i64 nTiles = (i64)G.nr * tilesPerRoot;
    const bool small = nTiles <= 0xffffffffll;
    const i64 bricksPerRoot = (i64)(G.per >> 9);
    for (i64 tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int r = small ? (int)((unsigned)tile / tilesPerRoot) : (int)(tile / tilesPerRoot);
        const int i = (int)(tile - (i64)r * tilesPerRoot) * 128 + (int)threadIdx.x;
        if (i >= per) continue;
        const int a = i / nn, rem = i - a * nn, u = rem / n1, v = rem - u * n1;
        const i64 t = (i64)r * per + i;                                    // = ((r * 3 + a) * n1 + u) * n1 + v
        int gx = a == 0 ? 0 : u, gy = a == 0 ? u : (a == 1 ? 0 : v), gz = a == 2 ? 0 : v;
        if ((a >= 1 && gx == 0) || (a == 2 && gy == 0)) continue;   // stored under the lowest zero axis
        if (needed) {
            const int bx0 = max(gx - 1, 0) >> 3, bx1 = min(gx, G.n - 1) >> 3, by0 = max(gy - 1, 0) >> 3, by1 = min(gy, G.n - 1) >> 3,
                      bz0 = max(gz - 1, 0) >> 3, bz1 = min(gz, G.n - 1) >> 3;
            bool any = false;
            for (int bx = bx0; bx <= bx1; bx++)
                for (int by = by0; by <= by1; by++)
                    for (int bz = bz0; bz <= bz1; bz++) any = any || needed[r * bricksPerRoot + rv_morton(bx, by, bz)] != 0;
            if (!any) continue;
        }
        int r2, ox, oy, oz, jb;
        rv_point_owner(G, r, gx, gy, gz, r2, ox, oy, oz, jb);
        if (r2 != r || jb == 7) continue;
        // only the bricks that are staged for classification read grid values: the point lies in the closed box of its owner's brick
        if (needed && !needed[(int)(((i64)r * G.per + rv_morton(ox, oy, oz)) >> 9)]) continue;
        G.low[t] = rv_eval_slow(G, r, ox, oy, oz, jb);
    }
}
// owner of edge e of cell (r, c): (flat owner cell index, edge kind in the owner's frame)
__device__ __forceinline__ i64 rv_edge_owner(const RGeom& G, int r, int cx, int cy, int cz, int e, int& e2) {
    const int o = e >> 2;
    int a0, a1;
    other_axes(o, a0, a1);
    const int s0 = (e & 1) ? 1 : -1, s1 = (e & 2) ? 1 : -1;
    int c[3] = {cx, cy, cz};
    int bm;
    i64 owner;
    if ((s0 > 0 || c[a0] >= 1) && (s1 > 0 || c[a1] >= 1)) {
        bm = (s0 < 0 ? 1 : 0) | (s1 < 0 ? 2 : 0);
        if (s0 < 0) c[a0] -= 1;
        if (s1 < 0) c[a1] -= 1;
        owner = (i64)r * G.per + rv_morton(c[0], c[1], c[2]);
    } else {
        unsigned long long best = ~0ull;
        bm = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int d[3] = {cx, cy, cz}, rr;
            if (q & 1) d[a0] += s0;
            if (q & 2) d[a1] += s1;
            if (!rv_locate(G, r, d[0], d[1], d[2], rr)) continue;
            unsigned long long key = ((unsigned long long)rr << 32) | rv_morton(d[0], d[1], d[2]);
            if (key < best) { best = key; bm = q; }
        }
        owner = (i64)(best >> 32) * G.per + (unsigned)(best & 0xffffffffu);
    }
    e2 = (o << 2) | ((e & 1) ^ (bm & 1)) | ((((e >> 1) & 1) ^ ((bm >> 1) & 1)) << 1);
    return owner;
}
__device__ __forceinline__ void rv_cell_values(const RGeom& G, int r, int cx, int cy, int cz, float v[8]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int j = ring_to_bits(q);
        v[q] = rv_point_value(G, r, cx + (j & 1), cy + ((j >> 1) & 1), cz + ((j >> 2) & 1));
    }
}
// ---- certified signs without evaluation.  Hundreds of thousands (depth 10) to millions (depth 11)
// of bricks lie far from the surface; evaluating their 512 cells only to find one common sign is
// most of the refinement cost.  Inside a brick no basis function of a level <= D-3 has a knot, so
// every level's contribution is ONE tensor-product quadratic there and chi - iso on the brick's closed
// box (its 9x9x9 grid points, evaluated along the chain of any of its cells) is bounded by its 27 Bernstein coefficients (convex hull property), computed in
// double from the same piece tables.  The float evaluation the kernels above would perform differs
// from that polynomial by at most E = sum |x_j| (e_x M_y M_z + M_x e_y M_z + M_x M_y e_z + 1e-4 M_x M_y M_z)
// (e: evaluation error bound of the cumulative-piece form, 8 ulp of the sum of the absolute monomial
// values; M: bound of the factor on the interval; 1e-4 covers the <= 330 product / accumulation
// roundings).  A brick is certified (1: all > 0, 2: all < 0) only when every coefficient clears iso
// by 2E; bricks with contributions from levels finer than D-3 are never certified.  Certified
// bricks are not evaluated unless a brick that has to be classified reads them.
// S = log2 of the box edge in cells: 3 = one brick; 5 = a 4x4x4 block of bricks (64 consecutive bricks in Morton
// order), certified as a whole first -- far from the surface that settles 64 bricks with one test, and the brick
// pass (S = 3, `coarse` = the S = 5 certificates) only looks at the rest.  The argument is the same with D-S in
// place of D-3: no knots of levels <= D-S inside the box, and no contribution from finer levels.
__global__ void __launch_bounds__(256) k_rv_brick_bound(RGeom G, int S, int nBricks, const unsigned char* __restrict__ coarse, unsigned char* __restrict__ cert) {
    __shared__ double sBeta[8][9][3];        // [warp][axis*3 + k][Bernstein index]
    __shared__ double sErr[8][9], sMax[8][9];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int L = G.D - S;
    const double w = 1.0 / (double)(1 << G.D);
    const int ci = lane / 9, cj = (lane / 3) % 3, ck = lane % 3;
    for (int b = blockIdx.x * 8 + wp; b < nBricks; b += gridDim.x * 8) {
        if (coarse) {
            const unsigned char cb = coarse[b >> 6];
            if (cb) { if (lane == 0) cert[b] = cb; continue; }
        }
        const i64 cell0 = (i64)b << (3 * S);
        const int r = (int)(cell0 / G.per);
        const unsigned l0 = (unsigned)(cell0 - (i64)r * G.per);
        const ushort4 ro = G.offs[G.roots[r]];
        const int g[3] = {((int)ro.x << G.lv) + (int)compact3(l0 >> 2), ((int)ro.y << G.lv) + (int)compact3(l0 >> 1), ((int)ro.z << G.lv) + (int)compact3(l0)};
        double acc = 0.0, E = 0.0;
        int cur = -1;
        if (lane < 27) { int q = G.rootNb[27 * r + lane]; cur = (q >= 0 && q < G.M) ? q : -1; }
        bool broke = false;
        for (int lvl = 0; lvl <= L; lvl++) {
            // the 27 neighbour values of the brick's ancestor at this level
            float X = 0.f;
            if (lvl <= G.rd) {
                if (lane < 27) X = G.rootX[((i64)r * (G.rd + 1) + lvl) * 27 + lane];
            } else {
                int c = (int)((l0 >> (3 * (G.D - lvl))) & 7u), pj = 0, cc = 0;
                if (lane < 27) lut_parent_child(c, lane, pj, cc);
                int p = __shfl_sync(0xffffffffu, cur, pj);
                int nxt = -1;
                if (lane < 27 && p >= 0) { int c0 = G.child0[p]; if (c0 >= 0) nxt = c0 + cc; }
                if (nxt >= 0) X = G.x[nxt];
                cur = nxt;
                if (!__any_sync(0xffffffffu, nxt >= 0)) { broke = true; break; }      // nothing real at this level: nothing below it either
            }
            // lanes 0..8: the quadratic of function k of axis a on the brick's own points [t0, t1]
            __syncwarp();
            if (lane < 9) {
                const int a = lane / 3, k = lane % 3;
                const int nn = 1 << lvl, ao = (g[a] >> (G.D - lvl)) + k - 1;
                double C0 = 0.0, C1 = 0.0, C2 = 0.0, A = 0.0;
                const double t0 = (double)g[a] * w, t1 = (double)(g[a] + (1 << S)) * w, tm = 0.5 * (t0 + t1);
                if (ao >= 0 && ao < nn) {
                    const float* f = G.baseFn + 20 * (i64)(nn - 1 + ao);
                    bool on = true;
                    for (int i = 0; i < 4; i++) {
                        on = on && tm > (double)f[5 * i + 4];
                        if (on) { C0 += (double)f[5 * i]; C1 += (double)f[5 * i + 1]; C2 += (double)f[5 * i + 2]; }
                        // (all pieces count for the error bound: a piece whose start sits on the brick's edge may
                        // switch on at the last grid point, where it is zero up to its own evaluation noise)
                        A += fabs((double)f[5 * i]) + fabs((double)f[5 * i + 1]) * t1 + fabs((double)f[5 * i + 2]) * t1 * t1;
                    }
                }
                const double b0 = C0 + C1 * t0 + C2 * t0 * t0, b2 = C0 + C1 * t1 + C2 * t1 * t1, b1 = C0 + C1 * tm + C2 * t0 * t1;
                sBeta[wp][lane][0] = b0; sBeta[wp][lane][1] = b1; sBeta[wp][lane][2] = b2;
                sErr[wp][lane] = 8.0 * 1.1920929e-7 * A;
                sMax[wp][lane] = fmax(fabs(b0), fmax(fabs(b1), fabs(b2)));
            }
            __syncwarp();
            // lanes 0..26: Bernstein coefficient (ci, cj, ck) += sum_j X_j bx[jx][ci] by[jy][cj] bz[jz][ck]; lane j also the error term of x_j
            double term = 0.0;
            if (lane < 27) {
                const double ax = fabs((double)X);
                const int jx = lane / 9, jy = (lane / 3) % 3, jz = lane % 3;
                const double mx = sMax[wp][jx], my = sMax[wp][3 + jy], mz = sMax[wp][6 + jz];
                term = ax * (sErr[wp][jx] * my * mz + mx * sErr[wp][3 + jy] * mz + mx * my * sErr[wp][6 + jz] + 1e-4 * mx * my * mz);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
            E += term;
            // separable contraction over (jz, jy, jx): lane (a, b, c) = (lane / 9, (lane / 3) % 3, lane % 3)
            {
                const int l3 = lane < 27 ? lane - ck : 0;           // 9 a + 3 b
                double u = 0.0, v = 0.0, cc = 0.0;
#pragma unroll
                for (int jz = 0; jz < 3; jz++) u += (double)__shfl_sync(0xffffffffu, X, l3 + jz) * sBeta[wp][6 + jz][ck];          // (jx, jy, ck)
                const int l9 = lane < 27 ? 9 * ci + ck : 0;
#pragma unroll
                for (int jy = 0; jy < 3; jy++) v += __shfl_sync(0xffffffffu, u, l9 + 3 * jy) * sBeta[wp][3 + jy][cj];             // (jx, cj, ck)
                const int l1 = lane < 27 ? 3 * cj + ck : 0;
#pragma unroll
                for (int jx = 0; jx < 3; jx++) cc += __shfl_sync(0xffffffffu, v, l1 + 9 * jx) * sBeta[wp][jx][ci];                // (ci, cj, ck)
                if (lane < 27) acc += cc;
            }
        }
        // levels finer than the brick level (children of the level-L neighbours): never certified
        const bool finer = !broke && __any_sync(0xffffffffu, lane < 27 && cur >= 0 && G.child0[cur] >= 0);
        double lo = lane < 27 ? acc : 1e300, hi = lane < 27 ? acc : -1e300;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (lane == 0) {
            const double iso = (double)G.iso;
            const double margin = 2.0 * E + 1e-6 * (fabs(iso) + fmax(fabs(lo), fabs(hi))) + 1e-30;
            unsigned char c = 0;
            if (!finer) { if (lo - iso > margin) c = 1; else if (hi - iso < -margin) c = 2; }
            cert[b] = c;
        }
    }
}
// brick holding the virtual cell (x, y, z) of root r, coordinates in [-n, 2n) (-1: that cell is not virtual)
__device__ __forceinline__ int rv_brick_of(const RGeom& G, int r, int x, int y, int z) {
    int r2;
    if (!rv_locate(G, r, x, y, z, r2)) return -1;
    return (int)(((i64)r2 * G.per + rv_morton(x, y, z)) >> 9);
}
// The 9x9x9 grid of brick b' takes its values from virtual cells of b' and of the 7 bricks below it (in its
// root or in a virtual neighbour root); every one of them is evaluated along the chain of a cell of the
// brick that holds it, inside that brick's closed box.  b' has to be classified unless all of those bricks
// (that exist) are certified with one common sign.
__global__ void __launch_bounds__(256) k_rv_brick_full(RGeom G, int nBricks, const unsigned char* __restrict__ cert, unsigned char* __restrict__ full) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBricks; b += gridDim.x * blockDim.x) {
        const i64 cell0 = (i64)b * 512;
        const int r = (int)(cell0 / G.per);
        const unsigned l0 = (unsigned)(cell0 - (i64)r * G.per);
        const int bx = (int)compact3(l0 >> 2), by = (int)compact3(l0 >> 1), bz = (int)compact3(l0);
        const int s0 = cert[b];
        bool same = s0 != 0;
        for (int q = 1; q < 8 && same; q++) {
            const int src = rv_brick_of(G, r, bx - ((q & 1) ? 8 : 0), by - ((q & 2) ? 8 : 0), bz - ((q & 4) ? 8 : 0));
            if (src >= 0) same = cert[src] == s0;
        }
        full[b] = same ? 0 : 1;
    }
}
// brick b has to be evaluated when a brick that is classified reads it: itself or one of the 7 bricks above it
// (needLow: the lower-face grid points owned by a cell of b are read by any classified brick that touches the
// point, which can lie on either side of b in every axis: the whole 3x3x3 neighbourhood counts)
__global__ void __launch_bounds__(256) k_rv_brick_needed(RGeom G, int nBricks, const unsigned char* __restrict__ full, int* __restrict__ needed, unsigned char* __restrict__ needLow) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBricks; b += gridDim.x * blockDim.x) {
        const i64 cell0 = (i64)b * 512;
        const int r = (int)(cell0 / G.per);
        const unsigned l0 = (unsigned)(cell0 - (i64)r * G.per);
        const int bx = (int)compact3(l0 >> 2), by = (int)compact3(l0 >> 1), bz = (int)compact3(l0);
        bool need = false;
        for (int q = 0; q < 8 && !need; q++) {
            const int dst = rv_brick_of(G, r, bx + ((q & 1) ? 8 : 0), by + ((q & 2) ? 8 : 0), bz + ((q & 4) ? 8 : 0));
            if (dst >= 0) need = full[dst] != 0;
        }
        needed[b] = need ? 1 : 0;
        bool low = need;
        if (!low && (bx == 0 || by == 0 || bz == 0))
            for (int q = 0; q < 27 && !low; q++) {
                const int dst = rv_brick_of(G, r, bx + 8 * (q / 9 - 1), by + 8 * ((q / 3) % 3 - 1), bz + 8 * (q % 3 - 1));
                if (dst >= 0) low = full[dst] != 0;
            }
        needLow[b] = low ? 1 : 0;
    }
}
__global__ void __launch_bounds__(256) k_rv_bound_check(int nBricks, const unsigned char* __restrict__ cert, const unsigned char* __restrict__ sign, int* __restrict__ bad) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBricks; b += gridDim.x * blockDim.x)
        if (cert[b] != 0 && cert[b] != sign[b]) atomicAdd(bad, 1);
}
// sign summary of a brick's own 512 corner-7 values: 1 all > 0, 2 all < 0, 0 otherwise (one warp per brick)
__global__ void __launch_bounds__(256) k_rv_brick_sign(const float* __restrict__ val7, int nBricks, unsigned char* __restrict__ sign) {
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nBricks; b += (gridDim.x * blockDim.x) >> 5) {
        const float4* v = reinterpret_cast<const float4*>(val7 + (i64)b * 512);
        bool pos = true, neg = true;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 f = v[lane + 32 * k];
            pos = pos && f.x > 0.f && f.y > 0.f && f.z > 0.f && f.w > 0.f;
            neg = neg && f.x < 0.f && f.y < 0.f && f.z < 0.f && f.w < 0.f;
        }
        const bool ap = __all_sync(0xffffffffu, pos), an = __all_sync(0xffffffffu, neg);
        if (lane == 0) sign[b] = ap ? 1 : (an ? 2 : 0);
    }
}
// brick form of the classification: the 9x9x9 grid values of a brick are staged in shared memory
// (own corner-7 values + the 217 points of its three lower faces), every cell then reads its 8
// corners from there
// Output per cell (only for bricks that produce anything): case, mask of owned crossed edges and the
// exclusive prefix of the vertex count INSIDE the brick; per brick: vertex and triangle totals.  The
// pass-wide offsets then come from a scan over bricks (hundreds of thousands) instead of cells
// (hundreds of millions), and only the active bricks are visited again for the emission.
__global__ void __launch_bounds__(512) k_rv_classify_brick(RGeom G, const unsigned char* __restrict__ bsign /* full[] flags */, unsigned char* __restrict__ cat, unsigned short* __restrict__ emask,
                                                           unsigned short* __restrict__ vpre, int* __restrict__ brickV, int* __restrict__ brickT, int nBricks,
                                                           int chunk /* 1..64 bricks whose flags a CTA looks at together */) {
    __shared__ int sScan[33];
    __shared__ float sV[9 * 9 * 9];
    __shared__ int sList[64], sWarpCnt[2];
    const int tid = threadIdx.x;
    // persistent CTAs over chunks of bricks: the flags of a chunk are fetched together (one memory round trip per chunk instead of
    // one per brick -- most bricks are certified and leave at once) and only the flagged bricks of the chunk are visited
    const int nChunks = (nBricks + chunk - 1) / chunk;
    for (int ch = blockIdx.x; ch < nChunks; ch += gridDim.x) {
    __syncthreads();                                  // the previous chunk's list is consumed
    const int mine = ch * chunk + tid;
    const bool flagged = tid < chunk && mine < nBricks && bsign[mine] != 0;
    unsigned bal = 0;
    if (tid < 64) {
        if (tid < chunk && mine < nBricks && !flagged) { brickV[mine] = 0; brickT[mine] = 0; }
        bal = __ballot_sync(0xffffffffu, flagged);
        if ((tid & 31) == 0) sWarpCnt[tid >> 5] = __popc(bal);
    }
    __syncthreads();
    if (flagged) sList[(tid >= 32 ? sWarpCnt[0] : 0) + __popc(bal & ((1u << (tid & 31)) - 1u))] = mine;
    const int nList = sWarpCnt[0] + sWarpCnt[1];
    for (int li = 0; li < nList; li++) {
    __syncthreads();                                  // list written / the previous brick's shared values are consumed
    const int brick = sList[li];
    const i64 cell0 = (i64)brick * 512;
    const int r = (int)(cell0 / G.per);
    const unsigned l0 = (unsigned)(cell0 - (i64)r * G.per);
    const int bx = (int)compact3(l0 >> 2), by = (int)compact3(l0 >> 1), bz = (int)compact3(l0);   // root-local origin of the brick
    // (bricks whose grid is certified to be of one strict sign, k_rv_brick_bound / k_rv_brick_full, produce nothing: not listed)
    const int cx = (int)compact3((unsigned)tid >> 2), cy = (int)compact3((unsigned)tid >> 1), cz = (int)compact3((unsigned)tid);
    const i64 store0 = rv_store(G, cell0);
    const float own = G.val7[store0 + tid];
    sV[(cx + 1) * 81 + (cy + 1) * 9 + (cz + 1)] = own;
    bool pos = own > 0.f, neg = own < 0.f;
    if (tid < 217) {
        // points with a zero brick-local coordinate: 81 with x = 0, 72 with y = 0 (x >= 1), 64 with z = 0 (x, y >= 1)
        int gx, gy, gz;
        if (tid < 81) { gx = 0; gy = tid / 9; gz = tid % 9; }
        else if (tid < 153) { int t = tid - 81; gx = 1 + t / 9; gy = 0; gz = t % 9; }
        else { int t = tid - 153; gx = 1 + t / 8; gy = 1 + t % 8; gz = 0; }
        const float f = rv_point_value(G, r, bx + gx, by + gy, bz + gz);
        sV[gx * 81 + gy * 9 + gz] = f;
        pos = pos && f > 0.f; neg = neg && f < 0.f;
    }
    // a brick whose 729 grid values all have the same strict sign has no crossed edge and the
    // same trivial case (0 or 255, both without triangles) in every cell: nothing else to do
    const int allPos = __syncthreads_and(pos), allNeg = __syncthreads_and(neg);
    if (allPos || allNeg) {
        if (tid == 0) { brickV[brick] = 0; brickT[brick] = 0; }
        continue;
    }
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int j = ring_to_bits(q);
        v[q] = sV[(cx + (j & 1)) * 81 + (cy + ((j >> 1) & 1)) * 9 + (cz + ((j >> 2) & 1))];
    }
    int c = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) if (v[q] < 0.f) c |= 1 << q;
    unsigned m = 0;
    const i64 t = cell0 + tid;
#pragma unroll
    for (int e = 0; e < 12; e++) {
        if (__fmul_rn(v[cEdgeVertex[e][0]], v[cEdgeVertex[e][1]]) <= 0.f) {
            int e2;
            if (rv_edge_owner(G, r, bx + cx, by + cy, bz + cz, e, e2) == t) m |= 1u << e;
        }
    }
    int totV, totT;
    const int pre = block_exclusive_scan(__popc(m), &totV, sScan);
    block_exclusive_scan((int)cMcCount[c], &totT, sScan);
    if (tid == 0) { brickV[brick] = totV; brickT[brick] = totT; }
    if (totV | totT) {
        cat[store0 + tid] = (unsigned char)c;
        emask[store0 + tid] = (unsigned short)m;
        vpre[store0 + tid] = (unsigned short)pre;
    }
    }
    }
}
__global__ void __launch_bounds__(256) k_brick_flags(const int* __restrict__ brickV, const int* __restrict__ brickT, int n, int* __restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flag[i] = (brickV[i] | brickT[i]) != 0;
}
// emission of one active brick per CTA: vertices of the owned crossed edges, then the triangles
// (vertex ids through the owner cell's brick base + in-brick prefix + rank of the edge in its mask)
__global__ void __launch_bounds__(512) k_rv_emit_brick(RGeom G, const int* __restrict__ active, const int* __restrict__ activeCount, const unsigned char* __restrict__ cat, const unsigned short* __restrict__ emask,
                                                       const unsigned short* __restrict__ vpre, const int* __restrict__ brickVBase, const int* __restrict__ brickTBase,
                                                       float* __restrict__ outV, int* __restrict__ outT) {
    __shared__ int sScan[33];
    const int tid = threadIdx.x;
    const int nWork = *activeCount;
    for (int work = blockIdx.x; work < nWork; work += gridDim.x) {
    const int b = active[work];
    const i64 t = (i64)b * 512 + tid;
    const int r = (int)(t / G.per);
    const unsigned l = (unsigned)(t - (i64)r * G.per);
    const int cx = (int)compact3(l >> 2), cy = (int)compact3(l >> 1), cz = (int)compact3(l);
    const i64 ts = rv_store(G, t);
    const unsigned m = emask[ts];
    const int c = cat[ts], nt = cMcCount[c];
    int totT;
    const int tb = brickTBase[b] + block_exclusive_scan(nt, &totT, sScan);
    if (m) {
        const float w = 1.0f / (float)(1 << G.D);
        float v[8];
        rv_cell_values(G, r, cx, cy, cz, v);
        const ushort4 ro = G.offs[G.roots[r]];
        const int ox = ((int)ro.x << G.lv) + cx, oy = ((int)ro.y << G.lv) + cy, oz = ((int)ro.z << G.lv) + cz;
        int k = brickVBase[b] + (int)vpre[ts];
        for (int e = 0; e < 12; e++) {
            if (!(m & (1u << e))) continue;
            int r1 = cEdgeVertex[e][0], r2 = cEdgeVertex[e][1], dim = e >> 2;
            int b1 = ring_to_bits(r1), b2 = ring_to_bits(r2);
            float p1[3] = {(float)(ox + (b1 & 1)) * w, (float)(oy + ((b1 >> 1) & 1)) * w, (float)(oz + ((b1 >> 2) & 1)) * w};
            float p2d = (float)((dim == 0 ? ox + (b2 & 1) : (dim == 1 ? oy + ((b2 >> 1) & 1) : oz + ((b2 >> 2) & 1)))) * w;
            float f1 = v[r1], f2 = v[r2];
            float pivot = __fdiv_rn(f1, __fsub_rn(f1, f2));
            float another = __fsub_rn(1.0f, pivot);
            float out[3] = {p1[0], p1[1], p1[2]};
            out[dim] = __fmaf_rn(p2d, pivot, __fmul_rn(p1[dim], another));
            i64 a = 3 * (i64)k;
            outV[a] = out[0]; outV[a + 1] = out[1]; outV[a + 2] = out[2];
            k++;
        }
    }
    for (int j = 0; j < 3 * nt; j++) {
        int e = cMcTri[c][j], e2;
        const i64 ow = rv_edge_owner(G, r, cx, cy, cz, e, e2);
        const i64 os = rv_store(G, ow);
        outT[3 * (i64)tb + j] = brickVBase[(int)(ow >> 9)] + (int)vpre[os] + __popc((unsigned)emask[os] & ((1u << e2) - 1u));
    }
    }
}
struct PassOut {
    DBuf<float> v;
    DBuf<int> t;
    int nv = 0, nt = 0;
};

// multi-GPU main pass: every rank marches the depth-D cells of its Morton range
struct McShard {
    ValView W;                               // corner values of depth D, per owner rank
    int* vbaseFull = nullptr;                // [cnt[D]] in the arena (same offset on every rank): vertex offset of a cell inside its rank's part
    unsigned short* emaskFull = nullptr;     // [cnt[D]]
    size_t vbaseOff = 0, emaskOff = 0;
    int rankV[kMaxRanks] = {0}, rankT[kMaxRanks] = {0};     // vertices / triangles of every rank's part (out)
};

// classification + emission shared by the main pass and the refinement passes
static int run_mc_on_cells(Context& c, const Topo& T, const ValView& W, const ushort4* cellOffs, bool markFaces, unsigned* fmark, PassOut& out, McShard* sh = nullptr) {
    cudaStream_t st = c.stream;
    const int n = T.nCells, D = c.D;
    DBuf<unsigned char> cat;
    DBuf<unsigned short> emaskL;
    DBuf<int> ntri, nvtx, tbase, vbaseL;
    PRB_TRY(cat.alloc((size_t)n, st));
    PRB_TRY(ntri.alloc((size_t)n, st));
    PRB_TRY(nvtx.alloc((size_t)n, st));
    PRB_TRY(tbase.alloc((size_t)n, st));
    unsigned short* emask;
    int* vbase;
    if (sh) {                                // this rank's window of the arena arrays
        emask = sh->emaskFull + (T.cellBase - c.base[D]);
        vbase = sh->vbaseFull + (T.cellBase - c.base[D]);
    } else {
        PRB_TRY(emaskL.alloc((size_t)n, st));
        PRB_TRY(vbaseL.alloc((size_t)n, st));
        emask = emaskL.p; vbase = vbaseL.p;
    }
    if (n) PRB_LAUNCH(c, k_classify, grid_for(c, n, 128, 16), 128, 0, T, W, cat.p, ntri.p, emask, nvtx.p);
    i64 totV = 0, totT = 0;
    PRB_TRY(exclusive_scan(c, nvtx.p, vbase, n, &totV));
    PRB_TRY(exclusive_scan(c, ntri.p, tbase.p, n, &totT));
    out.nv = (int)totV;
    out.nt = (int)totT;
    TriView V;
    for (int r = 0; r < kMaxRanks; r++) { V.vbase[r] = vbase; V.emask[r] = emask; V.vtxBase[r] = 0; }
    for (int r = 0; r <= kMaxRanks; r++) V.lo[r] = 0;
    V.idxBase = T.cellBase; V.world = 1;
    if (sh) {
        // the other ranks' counts (and, with the barrier inside, their finished offset / mask arrays)
        int mine[2] = {out.nv, out.nt}, all[kMaxRanks][2];
        PRB_TRY(mg_exchange_ints(c, mine, 2, &all[0][0]));
        int accV = 0;
        for (int r = 0; r < c.mg.world; r++) {
            sh->rankV[r] = all[r][0]; sh->rankT[r] = all[r][1];
            V.vbase[r] = (const int*)(c.mg.peer[r] + sh->vbaseOff);
            V.emask[r] = (const unsigned short*)(c.mg.peer[r] + sh->emaskOff);
            V.vtxBase[r] = accV;
            accV += all[r][0];
        }
        V.idxBase = c.base[D]; V.world = c.mg.world;
        for (int r = 0; r <= c.mg.world; r++) V.lo[r] = c.rowLo[D][r];
    }
    PRB_TRY(out.v.alloc(3 * (size_t)totV, st));
    PRB_TRY(out.t.alloc(3 * (size_t)totT, st));
    if (totV) PRB_LAUNCH(c, k_emit_vertices, grid_for(c, n, 128, 16), 128, 0, T, W, cellOffs, c.D, emask, vbase, out.v.p);
    if (n && (totT || markFaces))
        PRB_LAUNCH(c, k_emit_triangles, grid_for(c, n, 128, 16), 128, 0, T, cat.p, ntri.p, tbase.p, V, out.t.p, markFaces ? 1 : 0, c.base[D],
                   c.parent.p, c.offs.p, fmark);
    return PRB_OK;
}

// refinement pass over roots at depth rd <= D-3 (implicit virtual subtrees, see RGeom)
static int refine_pass_implicit(Context& c, const int* dRoots, int nr, int rd, bool single, DBuf<int>& rootMap, std::vector<PassOut>& outs) {
    cudaStream_t st = c.stream;
    const int D = c.D, lv = D - rd;
    if (lv > 10) { set_error("refinement pass too large (root more than 10 levels above maxDepth)"); return PRB_ERR_NOMEM; }
    const int n = 1 << lv, n1 = n + 1;
    const unsigned per = 1u << (3 * lv);
    const i64 total = (i64)nr * per;
    if (total / 512 > 0x7fffffffll) { set_error("refinement pass too large (more than 2^31 bricks)"); return PRB_ERR_NOMEM; }
    if ((double)total / 512.0 * 40.0 > (double)c.deviceMemBytes * 0.5) { set_error("refinement pass too large for device memory (per-brick tables)"); return PRB_ERR_NOMEM; }
    DBuf<int> rootNb;
    DBuf<float> rootX;
    DBuf<float>& low = c.wsLow;
    const bool mg = c.mg.active();
    DBuf<unsigned char>& cat = c.wsCat;
    DBuf<unsigned short>& emask = c.wsEmask;
    PRB_TRY(rootNb.alloc(27 * (size_t)nr, st));
    PRB_TRY(rootX.alloc((size_t)nr * (rd + 1) * 27, st));
    PRB_TRY(low.ensure((size_t)nr * 3 * n1 * n1, st));
    PRB_LAUNCH(c, k_set_rootmap, grid_for(c, nr, 256), 256, 0, dRoots, nr, 0, 0, rootMap.p);
    PRB_LAUNCH(c, k_rv_roots, div_up((i64)nr * 32, 256), 256, 0, nr, rd, c.M, dRoots, rootMap.p, c.neighs.p, c.parent.p, c.xv, rootNb.p, rootX.p);
    PRB_LAUNCH(c, k_set_rootmap, grid_for(c, nr, 256), 256, 0, dRoots, nr, 0, -1, rootMap.p);
    RGeom G;
    G.M = c.M; G.D = D; G.rd = rd; G.lv = lv; G.n = n; G.nr = nr; G.per = per;
    G.roots = dRoots; G.rootNb = rootNb.p; G.rootX = rootX.p; G.offs = c.offs.p; G.child0 = c.child0.p; G.x = c.xv; G.baseFn = c.dBaseFn.p;
    PRB_TRY(ensure_bv_tables(c));
    G.bvCell = (const float4*)c.dBvCell.p;
    G.iso = c.iso; G.val7 = nullptr; G.low = low.p; G.slot = nullptr;
    const int nBricks = (int)(total / 512);
    // certified signs -> bricks to classify -> bricks to evaluate (every rank evaluates the same short list:
    // it is a few per cent of the pass, less than a rank's share of all bricks plus the peer pulls would be)
    DBuf<unsigned char> cert, full, needLow;
    DBuf<int> needFlag, needExcl, needList;
    PRB_TRY(cert.alloc((size_t)nBricks, st)); PRB_TRY(full.alloc((size_t)nBricks, st)); PRB_TRY(needLow.alloc((size_t)nBricks, st));
    PRB_TRY(needFlag.alloc((size_t)nBricks, st)); PRB_TRY(needExcl.alloc((size_t)nBricks, st));
    if (lv >= 5) {
        DBuf<unsigned char> superCert;
        const int nSuper = nBricks >> 6;
        PRB_TRY(superCert.alloc((size_t)nSuper, st));
        PRB_LAUNCH(c, k_rv_brick_bound, grid_for(c, (i64)nSuper * 32, 256, 8), 256, 0, G, 5, nSuper, (const unsigned char*)nullptr, superCert.p);
        PRB_LAUNCH(c, k_rv_brick_bound, grid_for(c, (i64)nBricks * 32, 256, 8), 256, 0, G, 3, nBricks, (const unsigned char*)superCert.p, cert.p);
        superCert.release();
    } else {
        PRB_LAUNCH(c, k_rv_brick_bound, grid_for(c, (i64)nBricks * 32, 256, 8), 256, 0, G, 3, nBricks, (const unsigned char*)nullptr, cert.p);
    }
    PRB_LAUNCH(c, k_rv_brick_full, grid_for(c, nBricks, 256), 256, 0, G, nBricks, cert.p, full.p);
    PRB_LAUNCH(c, k_rv_brick_needed, grid_for(c, nBricks, 256), 256, 0, G, nBricks, full.p, needFlag.p, needLow.p);
    // (the counts of the selected bricks stay on the device: the kernels that consume the lists read them there)
    DBuf<int> dCounts;
    PRB_TRY(dCounts.alloc(2, st));
    i64 nNeeded = 0;
    PRB_TRY(exclusive_scan(c, needFlag.p, needExcl.p, nBricks, &nNeeded));       // (host round trip: the number of stored bricks sizes the value arrays)
    PRB_CUDA(cudaMemcpyAsync(dCounts.p, c.scanWork.ticket + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
    const unsigned persist = (unsigned)std::min<i64>(nBricks, (i64)c.smCount * 32);
    // values, cases, masks and prefixes are stored for the bricks that are evaluated only (slot = position in that list; the test mode
    // that evaluates everything stores everything)
    const i64 nStored = c.refineBoundCheck ? (i64)nBricks : nNeeded;
    if ((double)nStored * 512.0 * 9.0 > (double)c.deviceMemBytes * 0.6) { set_error("refinement pass too large for device memory (" + std::to_string(nStored) + " bricks to evaluate)"); return PRB_ERR_NOMEM; }
    PRB_TRY(c.wsVal7.ensure((size_t)std::max<i64>(nStored, 1) * 512, st));
    float* val7p = c.wsVal7.p;
    G.val7 = val7p;
    G.slot = c.refineBoundCheck ? nullptr : needExcl.p;
    if (c.refineBoundCheck) {
        // debug / test mode: evaluate everything and verify every certificate against the real values
        PRB_LAUNCH(c, k_rv_brick_values, persist, 64, 0, G, (const int*)nullptr, (const int*)nullptr, nBricks);
        DBuf<unsigned char> sign;
        DBuf<int> bad;
        PRB_TRY(sign.alloc((size_t)nBricks, st)); PRB_TRY(bad.alloc(1, st));
        PRB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
        PRB_LAUNCH(c, k_rv_brick_sign, grid_for(c, (i64)nBricks * 32, 256), 256, 0, val7p, nBricks, sign.p);
        PRB_LAUNCH(c, k_rv_bound_check, grid_for(c, nBricks, 256), 256, 0, nBricks, cert.p, sign.p, bad.p);
        int hb = 0;
        PRB_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        sign.release(); bad.release();
        c.boundChecked += nBricks; c.boundEvaluated += nNeeded;
        if (hb) { set_error("refinement: " + std::to_string(hb) + " bricks certified with the wrong sign (depth " + std::to_string(rd) + " pass)"); return PRB_ERR_STATE; }
    } else {
        PRB_TRY(needList.alloc((size_t)nBricks, st));
        PRB_LAUNCH(c, k_compact_ids, grid_for(c, nBricks, 256), 256, 0, needFlag.p, needExcl.p, nBricks, needList.p);
        PRB_LAUNCH(c, k_rv_brick_values, persist, 64, 0, G, (const int*)needList.p, (const int*)dCounts.p, 0);
    }
    PRB_LAUNCH(c, k_rv_low_values, grid_for(c, (i64)nr * 3 * n1 * n1, 128, 16), 128, 0, G, c.refineBoundCheck ? (const unsigned char*)nullptr : (const unsigned char*)needLow.p);
    cert.release(); needFlag.release(); needList.release(); needLow.release();
    PRB_TRY(cat.ensure((size_t)std::max<i64>(nStored, 1) * 512, st));
    PRB_TRY(emask.ensure((size_t)std::max<i64>(nStored, 1) * 512, st));
    PRB_TRY(c.wsVpre.ensure((size_t)std::max<i64>(nStored, 1) * 512, st));
    DBuf<int> brickV, brickT, brickVBase, brickTBase, bflag, bexcl, active;
    PRB_TRY(brickV.alloc((size_t)nBricks, st)); PRB_TRY(brickT.alloc((size_t)nBricks, st));
    PRB_TRY(brickVBase.alloc((size_t)nBricks, st)); PRB_TRY(brickTBase.alloc((size_t)nBricks, st));
    PRB_TRY(bflag.alloc((size_t)nBricks, st)); PRB_TRY(bexcl.alloc((size_t)nBricks, st));
    {
        // chunk: as many bricks per flag fetch as leaves every CTA several chunks (a small pass keeps one brick per chunk)
        const int chunk = std::max(1, std::min(64, nBricks / (c.smCount * 4 * 4)));
        const int nChunks = (nBricks + chunk - 1) / chunk;
        PRB_LAUNCH(c, k_rv_classify_brick, (unsigned)std::min(nChunks, c.smCount * 4), 512, 0, G, full.p, cat.p, emask.p, c.wsVpre.p, brickV.p, brickT.p, nBricks, chunk);
    }
    full.release();
    PRB_LAUNCH(c, k_brick_flags, grid_for(c, nBricks, 256), 256, 0, brickV.p, brickT.p, nBricks, bflag.p);
    i64 totV = 0, totT = 0;
    PRB_TRY(exclusive_scan(c, bflag.p, bexcl.p, nBricks, nullptr));
    PRB_CUDA(cudaMemcpyAsync(dCounts.p + 1, c.scanWork.ticket + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
    PRB_TRY(exclusive_scan(c, brickV.p, brickVBase.p, nBricks, nullptr, 1));  // (total -> pinned word 1)
    PRB_TRY(exclusive_scan(c, brickT.p, brickTBase.p, nBricks, &totT));       // (the one host round trip of the pass: the output sizes)
    totV = ((volatile int*)c.hScanTotal)[1];
    outs.emplace_back();
    PassOut& po = outs.back();
    po.nv = (int)totV;
    po.nt = (int)totT;
    PRB_TRY(po.v.alloc(3 * (size_t)totV, st));
    PRB_TRY(po.t.alloc(3 * (size_t)totT, st));
    if (totV || totT) {
        PRB_TRY(active.alloc((size_t)nBricks, st));
        PRB_LAUNCH(c, k_compact_ids, grid_for(c, nBricks, 256), 256, 0, bflag.p, bexcl.p, nBricks, active.p);
        PRB_LAUNCH(c, k_rv_emit_brick, (unsigned)std::min<i64>(nBricks, (i64)c.smCount * 8), 512, 0, G, active.p, (const int*)(dCounts.p + 1), cat.p, emask.p, c.wsVpre.p, brickVBase.p,
                   brickTBase.p, po.v.p, po.t.p);
    }
    brickV.release(); brickT.release(); brickVBase.release(); brickTBase.release(); bflag.release(); bexcl.release(); active.release();
    if (single && po.nv == 0) {      // main.cu:4095-4103: nothing is inserted for a coarse root without crossings
        po.v.release(); po.t.release();
        outs.pop_back();
    } else {
        c.passes.push_back({single ? 1 : 2, po.nv, po.nt});
    }
    rootNb.release(); rootX.release();
    return PRB_OK;
}

static int refine_pass(Context& c, const int* dRoots, int nr, int rd, bool single, DBuf<int>& rootMap, std::vector<PassOut>& outs) {
    cudaStream_t st = c.stream;
    const int D = c.D, M = c.M;
    if (nr == 0) {
        if (!single) { c.passes.push_back({2, 0, 0}); }
        return PRB_OK;
    }
    if (D - rd >= 3 && c.refineImplicit) return refine_pass_implicit(c, dRoots, nr, rd, single, rootMap, outs);
    VTree V;
    V.M = M; V.D = D; V.rd = rd; V.nr = nr; V.roots = dRoots;
    i64 total = 0;
    for (int d = 0; d <= kMaxDepth + 1; d++) V.depthAddr[d] = 0;
    for (int d = rd; d <= D; d++) { V.depthAddr[d] = (int)total; total += (i64)nr << (3 * (d - rd)); }
    if (total * 27 > 0x7fffffffll * 4) { set_error("refinement pass too large"); return PRB_ERR_NOMEM; }
    DBuf<int> vneigh;
    PRB_TRY(vneigh.alloc(27 * (size_t)total, st));
    PRB_LAUNCH(c, k_set_rootmap, grid_for(c, nr, 256), 256, 0, dRoots, nr, V.depthAddr[rd], 0, rootMap.p);
    // solution at the 27 neighbours of every root's ancestors (level rd of the table is not used here: the virtual levels read vneigh)
    DBuf<int> rootNbTmp;
    DBuf<float> rootX;
    PRB_TRY(rootNbTmp.alloc(27 * (size_t)nr, st));
    PRB_TRY(rootX.alloc((size_t)nr * (rd + 1) * 27, st));
    PRB_LAUNCH(c, k_rv_roots, div_up((i64)nr * 32, 256), 256, 0, nr, rd, M, dRoots, rootMap.p, c.neighs.p, c.parent.p, c.xv, rootNbTmp.p, rootX.p);
    rootNbTmp.release();
    for (int d = rd; d <= D; d++) {
        i64 cnt = ((i64)nr << (3 * (d - rd))) * 27;
        PRB_LAUNCH(c, k_vneigh, grid_for(c, cnt, 256), 256, 0, V, d, c.neighs.p, c.parent.p, c.child0.p, c.offs.p, rootMap.p, vneigh.p);
    }
    PRB_LAUNCH(c, k_set_rootmap, grid_for(c, nr, 256), 256, 0, dRoots, nr, 0, -1, rootMap.p);
    const int nD = nr << (3 * (D - rd));
    Topo T;
    T.nbr = vneigh.p; T.rowBase = M; T.minId = M; T.cellBase = M + V.depthAddr[D]; T.nCells = nD;
    DBuf<ushort4> voffs;
    DBuf<float> sval;
    PRB_TRY(voffs.alloc((size_t)nD, st));
    PRB_TRY(sval.alloc(8 * (size_t)nD, st));
    PRB_LAUNCH(c, k_vcell_offsets, grid_for(c, nD, 256), 256, 0, V, c.offs.p, voffs.p);
    DBuf<unsigned> vmask;
    PRB_TRY(vmask.alloc((size_t)total, st));
    PRB_LAUNCH(c, k_vmask, grid_for(c, total, 256), 256, 0, V, total, vneigh.p, vmask.p);
    PRB_TRY(ensure_bv_tables(c));
    {
        if ((i64)nD * 8 > 0x7fffffffll) { set_error("refinement pass too large"); return PRB_ERR_NOMEM; }
        DBuf<int> clist, ccount;
        PRB_TRY(clist.alloc(4 * (size_t)nD, st));         // a grid point has ONE owner and a root's cells have (n+1)^3 <= 3.375 n^3 grid points
        PRB_TRY(ccount.alloc(1, st));
        PRB_CUDA(cudaMemsetAsync(ccount.p, 0, sizeof(int), st));
        PRB_LAUNCH(c, k_vcorner_list, grid_for(c, (i64)nD * 8, 256, 8), 256, 0, T, clist.p, ccount.p);
        PRB_LAUNCH(c, k_vvertex_values, grid_for(c, (i64)nD * 2, 128, 16), 128, 0, V, (const int*)clist.p, (const int*)ccount.p, vneigh.p, vmask.p, voffs.p, rootX.p, c.offs.p, c.xv,
                   (const float4*)c.dBvGrid.p, (const float4*)c.dBvCell.p, c.dBaseFn.p, c.iso, sval.p);
    }
    outs.emplace_back();
    PassOut& po = outs.back();
    PRB_TRY(run_mc_on_cells(c, T, local_view(sval.p, T.cellBase), voffs.p, false, nullptr, po));
    if (single && po.nv == 0) {      // main.cu:4095-4103: nothing is inserted for a coarse root without crossings
        po.v.release(); po.t.release();
        outs.pop_back();
    } else {
        c.passes.push_back({single ? 1 : 2, po.nv, po.nt});
    }
    vneigh.release(); voffs.release(); sval.release(); vmask.release();
    return PRB_OK;
}

// base-value tables of k_vertex_values_stream (depend on the depth only: built once per context)
static int ensure_bv_tables(Context& c) {
    if (c.dBvAnc.p) return PRB_OK;
    const int D = c.D;
    size_t na = 0, no = 0;
    for (int d = 0; d <= kMaxDepth; d++) { c.bvAncOff[d] = 0; c.bvOwnOff[d] = 0; }
    for (int d0 = 1; d0 <= D; d0++) {
        const size_t ng = (size_t)1 << (d0 - 1);
        c.bvAncOff[d0] = (int)na; c.bvOwnOff[d0] = (int)no;
        na += (size_t)d0 * ng * 3;
        no += ng * 3;
    }
    PRB_TRY(c.dBvAnc.alloc(4 * na, c.stream));
    PRB_TRY(c.dBvOwn.alloc(4 * no, c.stream));
    PRB_TRY(c.dBvCell.alloc(4 * ((size_t)(D + 1) << D), c.stream));
    PRB_TRY(c.dBvGrid.alloc(4 * (size_t)(D + 1) * (((size_t)1 << D) + 1), c.stream));
    BvTables B;
    B.anc = (const float4*)c.dBvAnc.p; B.own = (const float4*)c.dBvOwn.p; B.cellD = (const float4*)c.dBvCell.p; B.gridLo = (const float4*)c.dBvGrid.p;
    for (int d = 0; d <= kMaxDepth; d++) { B.ancOff[d] = c.bvAncOff[d]; B.ownOff[d] = c.bvOwnOff[d]; }
    PRB_LAUNCH(c, k_build_bv, c.smCount * 4, 256, 0, D, c.dBaseFn.p, B, (float4*)c.dBvAnc.p, (float4*)c.dBvOwn.p, (float4*)c.dBvCell.p, (float4*)c.dBvGrid.p);
    return PRB_OK;
}

// cost model of a refinement pass (multi-GPU: whole passes are dealt out to the ranks, largest first)
static double pass_weight(int D, int rd, int nr) {
    const int lv = D - rd;
    if (lv >= 3) return (double)nr * (std::pow(8.0, lv) / 512.0 + 6.0 * std::pow(4.0, lv));     // brick certificates + the evaluated shell
    return (double)nr * std::pow(8.0, lv) * 40.0;                                                  // materialised subtrees: every virtual cell is evaluated
}

// whole passes -> ranks: largest estimated cost first, each to the least loaded rank (ties: lowest rank).  Pure host arithmetic on
// replicated inputs, so every rank computes the same deal (exported as prb_mg_deal_passes for the CPU tests).
void deal_passes(int D, int n, const int* depth, const int* count, int world, int* owner) {
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pass_weight(D, depth[a], count[a]) > pass_weight(D, depth[b], count[b]); });
    double load[kMaxRanks] = {0};
    for (int i : order) {
        int best = 0;
        for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
        owner[i] = best;
        load[best] += pass_weight(D, depth[i], count[i]);
    }
}

int stage_extract(Context& c) {
    cudaStream_t st = c.stream;
    const int D = c.D, M = c.M;
    PRB_TRY(upload_mc_tables(c.device));
    c.passes.clear();
    c.layout.clear();
    c.subdivide.clear();
    c.hMeshValid = false;
    if (c.copyStream) PRB_CUDA(cudaStreamSynchronize(c.copyStream));
    c.earlyV = c.earlyT = 0;
    Topo R;
    R.nbr = c.neighs.p; R.rowBase = 0; R.minId = 0; R.cellBase = c.base[D]; R.nCells = c.cnt[D];
    const bool mg = c.mg.active();
    const int W = c.mg.world, me = c.mg.rank;
    const bool shard = mg && D >= c.shardFrom;            // the depth-D cells (and every depth >= shardFrom) are split by Morton range
    const int nUpper = c.base[D];
    if (mg) {
        if (!c.mgVval) c.mgVval = c.mg.alloc<float>(8 * (size_t)M, &c.mgVvalOff);
        if (!c.mgVval) { set_error("multi-GPU arena too small for the corner values (32 bytes per node)"); return PRB_ERR_NOMEM; }
        c.vvalPtr = c.mgVval;
    } else {
        PRB_TRY(c.vval.alloc(8 * (size_t)M, st));
        c.vvalPtr = c.vval.p;
    }
    // ---- corner values.  Multi-GPU: a rank evaluates its node range of every sharded depth (and all of the small replicated
    // depths); the values of depth D stay where they are: the marching cubes of a rank's own cells only reads a neighbour rank's
    // values on a shard boundary
    {
        PRB_TRY(ensure_bv_tables(c));
        BvTables B;
        B.anc = (const float4*)c.dBvAnc.p; B.own = (const float4*)c.dBvOwn.p; B.cellD = (const float4*)c.dBvCell.p; B.gridLo = (const float4*)c.dBvGrid.p;
        for (int d = 0; d <= kMaxDepth; d++) { B.ancOff[d] = c.bvAncOff[d]; B.ownOff[d] = c.bvOwnOff[d]; }
        VsRanges RG;
        RG.n = 0;
        i64 totalGroups = 0;
        auto add = [&](int g0, int g1) {
            if (g1 <= g0 || RG.n >= 16) return;
            RG.first[RG.n] = g0; RG.count[RG.n] = g1 - g0; RG.n++;
            totalGroups += g1 - g0;
        };
        auto launch = [&]() {
            if (!RG.n) return;
            // chunk length: long enough to reuse the cached ancestors, short enough that every SM gets several warps' worth of chunks
            i64 ck = totalGroups / ((i64)c.smCount * 32);
            RG.chunk = (int)std::min<i64>(kVsChunk, std::max<i64>(4, ck & ~(i64)3));
            RG.chunk0[0] = 0;
            for (int k = 0; k < RG.n; k++) RG.chunk0[k + 1] = RG.chunk0[k] + (RG.count[k] + RG.chunk - 1) / RG.chunk;
            for (int k = RG.n; k < 16; k++) { RG.first[k] = 0; RG.count[k] = 0; RG.chunk0[k + 1] = RG.chunk0[RG.n]; }
            const i64 nChunks = RG.chunk0[RG.n];
            PRB_LAUNCH(c, k_vertex_values_stream, grid_for(c, nChunks * 32, kVsWarps * 32, 8), kVsWarps * 32, 0, R, RG, D, c.parent.p, c.child0.p, c.offs.p,
                       c.xv, c.dBaseFn.p, B, c.iso, c.vvalPtr);
        };
        if (!shard) {
            add(0, (M - 1) / 8);
            launch();
        } else {
            add(0, (c.base[c.shardFrom] - 1) / 8);
            for (int d = c.shardFrom; d <= D; d++) add((c.rowLo[d][me] - 1) / 8, (c.rowLo[d][me + 1] - 1) / 8);
            launch();
            PRB_TRY(mg_barrier(c));               // every rank's values are complete before anyone reads a neighbour rank's
            // the values of the depths ABOVE D are gathered (k_find_subdivide, on every rank, looks at all leaves up there; reading them in
            // place costs 8 scattered 4-byte NVLink reads per leaf: 1.2 ms on 8 GPUs against 0.3 ms for the bulk pull)
            {
                const void* src[32];
                void* dst[32];
                size_t bytes[32];
                int n = 0;
                for (int qi = 1; qi < W; qi++) {          // start with the next rank: the peers are not all pulled from in the same order
                    const int q = (me + qi) % W;
                    const float* from = (const float*)(c.mg.peer[q] + c.mgVvalOff);
                    for (int d = c.shardFrom; d < D; d++) {
                        const size_t a = (size_t)c.rowLo[d][q], b = (size_t)c.rowLo[d][q + 1];
                        if (b <= a) continue;
                        if (n == 32) { PRB_TRY(mg_pull(c, n, src, dst, bytes)); n = 0; }
                        src[n] = from + 8 * a; dst[n] = c.vvalPtr + 8 * a; bytes[n] = 32 * (b - a); n++;
                    }
                }
                PRB_TRY(mg_pull(c, n, src, dst, bytes));
            }
            PRB_TRY(mg_barrier(c));
        }
    }
    mark(c, "extract:corner_values");
    // ---- main pass
    DBuf<unsigned> fmarkL;
    unsigned* fmark = nullptr;
    size_t fmarkOff = 0;
    if (shard) {
        fmark = c.mg.alloc<unsigned>((size_t)nUpper, &fmarkOff);
        if (!fmark) { set_error("multi-GPU arena too small for the face marks"); return PRB_ERR_NOMEM; }
    } else {
        PRB_TRY(fmarkL.alloc((size_t)nUpper, st));
        fmark = fmarkL.p;
    }
    PRB_CUDA(cudaMemsetAsync(fmark, 0, sizeof(unsigned) * (size_t)nUpper, st));
    std::vector<PassOut> outs;
    outs.reserve(64);
    outs.emplace_back();
    McShard sh;
    if (shard) {
        sh.vbaseFull = c.mg.alloc<int>((size_t)c.cnt[D], &sh.vbaseOff);
        sh.emaskFull = c.mg.alloc<unsigned short>((size_t)c.cnt[D], &sh.emaskOff);
        if (!sh.vbaseFull || !sh.emaskFull) { set_error("multi-GPU arena too small for the marching-cubes offsets (6 bytes per depth-D slot)"); return PRB_ERR_NOMEM; }
        for (int r = 0; r < kMaxRanks; r++) sh.W.p[r] = r < W ? (const float*)(c.mg.peer[r] + c.mgVvalOff) : nullptr;
        sh.W.valBase = 0; sh.W.world = W;
        for (int r = 0; r <= kMaxRanks; r++) sh.W.lo[r] = r <= W ? c.rowLo[D][r] : 0;
        Topo Rm = R;
        Rm.cellBase = c.rowLo[D][me]; Rm.nCells = c.rowLo[D][me + 1] - c.rowLo[D][me];
        PRB_TRY(run_mc_on_cells(c, Rm, sh.W, c.offs.p + Rm.cellBase, true, fmark, outs.back(), &sh));
        PRB_TRY(mg_barrier(c));               // every rank's face marks are complete
    } else {
        PRB_TRY(run_mc_on_cells(c, R, local_view(c.vvalPtr, 0), c.offs.p + c.base[D], true, fmark, outs.back()));
        sh.rankV[0] = outs.back().nv; sh.rankT[0] = outs.back().nt;
    }
    // ---- the main piece is final from here on (it leads this rank's mesh and its triangles carry global vertex ids): with the option
    // "early_mesh_copy" its device -> host copy starts now, on a second stream, and runs under the refinement passes; prb_get_mesh
    // copies the rest.  (Off by default -- no net gain measured, see Context::earlyMeshCopy.)
    bool earlyIssued = false;
    if (c.earlyMeshCopy && c.doRefine && (outs[0].nv || outs[0].nt)) {
        PRB_TRY(ensure_copy_stream(c));
        // (a first run allocates the pinned buffers here, with the head room HBuf adds; if the refinement pieces do not fit later,
        // prb_get_mesh reallocates and copies everything)
        const size_t nv0 = (size_t)outs[0].nv, nt0 = (size_t)outs[0].nt;
        PRB_TRY(c.hMeshV.reserve(3 * nv0 + 3 * (nv0 / 2) + 1));
        PRB_TRY(c.hMeshT.reserve(3 * nt0 + 3 * (nt0 / 2) + 1));
        PRB_CUDA(cudaEventRecord(c.evMainPiece, st));
        PRB_CUDA(cudaStreamWaitEvent(c.copyStream, c.evMainPiece, 0));
        const size_t kPiece = (size_t)1 << 21;
        for (size_t o = 0; o < 12 * nv0; o += kPiece)
            PRB_CUDA(cudaMemcpyAsync((char*)c.hMeshV.p + o, (const char*)outs[0].v.p + o, std::min(kPiece, 12 * nv0 - o), cudaMemcpyDeviceToHost, c.copyStream));
        for (size_t o = 0; o < 12 * nt0; o += kPiece)
            PRB_CUDA(cudaMemcpyAsync((char*)c.hMeshT.p + o, (const char*)outs[0].t.p + o, std::min(kPiece, 12 * nt0 - o), cudaMemcpyDeviceToHost, c.copyStream));
        PRB_CUDA(cudaEventRecord(c.evEarlyCopy, c.copyStream));
        c.earlyV = outs[0].nv; c.earlyT = outs[0].nt;
        earlyIssued = true;
    }
    mark(c, "extract:main_pass");
    const int nMainParts = shard ? W : 1;
    // ---- leaves to refine (multi-GPU: the same list on every rank, from the OR of all ranks' face marks)
    DBuf<int> flag, excl, subIds;
    PRB_TRY(flag.alloc((size_t)nUpper, st));
    PRB_TRY(excl.alloc((size_t)nUpper, st));
    {
        MarkView F;
        F.world = shard ? W : 1;
        for (int r = 0; r < kMaxRanks; r++) F.p[r] = (shard && r < W) ? (const unsigned*)(c.mg.peer[r] + fmarkOff) : fmark;
        UpperView U;
        U.world = 1; U.shardFrom = c.shardFrom;      // (gathered above: everything is local; the view can also read the ranks' shares in place)
        for (int r = 0; r < kMaxRanks; r++) U.p[r] = c.vvalPtr;
        for (int d = 0; d <= kMaxDepth; d++)
            for (int r = 0; r <= kMaxRanks; r++) U.rowLo[d][r] = (d <= D && r <= W) ? c.rowLo[d][r] : 0;
        PRB_LAUNCH(c, k_find_subdivide, grid_for(c, nUpper, 128, 16), 128, 0, R, nUpper, c.child0.p, c.offs.p, U, F, flag.p);
    }
    i64 nSub = 0;
    PRB_TRY(exclusive_scan(c, flag.p, excl.p, nUpper, &nSub));
    PRB_TRY(subIds.alloc((size_t)nSub, st));
    if (nSub) {
        PRB_LAUNCH(c, k_compact_ids, grid_for(c, nUpper, 256), 256, 0, flag.p, excl.p, nUpper, subIds.p);
        c.subdivide.resize((size_t)nSub);
        PRB_CUDA(cudaMemcpyAsync(c.subdivide.data(), subIds.p, sizeof(int) * (size_t)nSub, cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
    }
    mark(c, "extract:subdivide_list");
    flag.release(); excl.release();
    // ---- refinement passes in the reference's order: single roots of depth 1, 2 (main.cu:3887-4202), then one batch per depth
    // (main.cu:4211-4561).  Passes are independent of each other (vertices are shared inside a pass only), so with several GPUs whole
    // passes are dealt out: largest estimated cost first, to the least loaded rank
    struct PassPlan { int kind, depth, first, count, owner, out; };
    std::vector<PassPlan> plan;
    if (c.doRefine) {
        std::vector<int> firstOfDepth(D + 2, (int)nSub);      // node ids ascend with depth, so the list is grouped by depth already
        size_t q = 0;
        for (int d = 0; d <= D; d++) {
            while (q < c.subdivide.size() && c.subdivide[q] < c.base[d]) q++;
            firstOfDepth[d] = (int)q;
        }
        firstOfDepth[D + 1] = (int)nSub;
        const int finerDepth = 3;    // main.cu:3886
        for (int d = 1; d < finerDepth && d < D; d++)
            for (int k = firstOfDepth[d]; k < firstOfDepth[d + 1]; k++) plan.push_back({1, d, k, 1, 0, -1});
        for (int d = finerDepth; d < D; d++) plan.push_back({2, d, firstOfDepth[d], firstOfDepth[d + 1] - firstOfDepth[d], 0, -1});
        if (mg) {
            std::vector<int> dep(plan.size()), cnt(plan.size()), own(plan.size());
            for (size_t i = 0; i < plan.size(); i++) { dep[i] = plan[i].depth; cnt[i] = plan[i].count; }
            deal_passes(D, (int)plan.size(), dep.data(), cnt.data(), W, own.data());
            for (size_t i = 0; i < plan.size(); i++) plan[i].owner = own[i];
        }
        DBuf<int> rootMap;
        PRB_TRY(rootMap.alloc((size_t)M, st));
        PRB_CUDA(cudaMemsetAsync(rootMap.p, 0xff, sizeof(int) * (size_t)M, st));
        for (auto& pp : plan) {
            if (mg && pp.owner != me) continue;
            const size_t before = outs.size();
            const size_t passesBefore = c.passes.size();
            PRB_TRY(refine_pass(c, subIds.p + pp.first, pp.count, pp.depth, pp.kind == 1, rootMap, outs));
            c.passes.resize(passesBefore);              // (the global list is rebuilt below)
            if (c.detail) { static const char* nm[13] = {"pass:d0", "pass:d1", "pass:d2", "pass:d3", "pass:d4", "pass:d5", "pass:d6", "pass:d7", "pass:d8", "pass:d9", "pass:d10", "pass:d11", "pass:d12"}; mark(c, nm[pp.depth]); }
            pp.out = outs.size() > before ? (int)outs.size() - 1 : -1;
        }
        rootMap.release();
    }
    mark(c, "extract:passes");
    subIds.release();
    // ---- counts of every pass on every rank -> global pass list, vertex / triangle bases of the local pieces
    const int nPlan = (int)plan.size();
    std::vector<int> pnv(nPlan, 0), pnt(nPlan, 0);
    for (int i = 0; i < nPlan; i++)
        if (plan[i].out >= 0) { pnv[i] = outs[plan[i].out].nv; pnt[i] = outs[plan[i].out].nt; }
    if (mg) {
        for (int i0 = 0; i0 < nPlan; i0 += 32) {
            const int n = std::min(32, nPlan - i0);
            int mine[64], all[kMaxRanks][64];
            for (int k = 0; k < n; k++) { mine[2 * k] = pnv[i0 + k]; mine[2 * k + 1] = pnt[i0 + k]; }
            PRB_TRY(mg_exchange_ints(c, mine, 2 * n, &all[0][0], 64));
            for (int k = 0; k < n; k++) { pnv[i0 + k] = all[plan[i0 + k].owner][2 * k]; pnt[i0 + k] = all[plan[i0 + k].owner][2 * k + 1]; }
        }
        if (nPlan == 0) PRB_TRY(mg_barrier(c));      // nobody leaves while a peer may still read this rank's arena
    }
    mark(c, "extract:counts_exchanged");
    i64 gv = 0, gt = 0;
    struct Piece { int out; i64 vBase, tBase; int nv, nt, pass; };
    std::vector<Piece> mine;
    {
        int mainV = 0, mainT = 0;
        for (int r = 0; r < nMainParts; r++) {
            if (r == (shard ? me : 0)) mine.push_back({0, gv, gt, outs[0].nv, outs[0].nt, 0});
            gv += sh.rankV[r]; gt += sh.rankT[r];
            mainV += sh.rankV[r]; mainT += sh.rankT[r];
        }
        c.passes.push_back({0, mainV, mainT});
    }
    for (int i = 0; i < nPlan; i++) {
        if (plan[i].kind == 1 && pnv[i] == 0) continue;      // main.cu:4095-4103: nothing is inserted for a coarse root without crossings
        if (plan[i].out >= 0) mine.push_back({plan[i].out, gv, gt, pnv[i], pnt[i], (int)c.passes.size()});
        c.passes.push_back({plan[i].kind, pnv[i], pnt[i]});
        gv += pnv[i]; gt += pnt[i];
    }
    if (gv > 0x7fffffffll / 3 || gt > 0x7fffffffll / 3) { set_error("mesh too large for 32-bit indices"); return PRB_ERR_NOMEM; }
    // ---- this rank's pieces, concatenated (insertTriangle, main.cu:3220-3245: indices offset by the vertices so far)
    i64 tv = 0, tt = 0;
    for (auto& pc : mine) { tv += pc.nv; tt += pc.nt; }
    PRB_TRY(c.meshV.alloc(3 * (size_t)tv, st));
    PRB_TRY(c.meshT.alloc(3 * (size_t)tt, st));
    i64 av = 0, at = 0;
    for (auto& pc : mine) {
        PassOut& o = outs[pc.out];
        if (pc.nv) PRB_CUDA(cudaMemcpyAsync(c.meshV.p + 3 * av, o.v.p, 12 * (size_t)pc.nv, cudaMemcpyDeviceToDevice, st));
        if (pc.nt) {
            PRB_CUDA(cudaMemcpyAsync(c.meshT.p + 3 * at, o.t.p, 12 * (size_t)pc.nt, cudaMemcpyDeviceToDevice, st));
            // (the triangles of the main pass already carry global vertex ids)
            if (pc.pass != 0 && pc.vBase) PRB_LAUNCH(c, k_offset_triangles, grid_for(c, 3 * (i64)pc.nt, 256), 256, 0, c.meshT.p + 3 * at, 3 * (i64)pc.nt, (int)pc.vBase);
        }
        c.layout.push_back({(i64)pc.pass, pc.vBase, (i64)pc.nv, pc.tBase, (i64)pc.nt});
        av += pc.nv; at += pc.nt;
    }
    mark(c, "extract:assembled");
    if (earlyIssued) PRB_CUDA(cudaStreamWaitEvent(st, c.evEarlyCopy, 0));      // the piece buffers go back to the arena: not before the copy has read them
    for (auto& o : outs) { o.v.release(); o.t.release(); }
    c.nMeshV = tv;
    c.nMeshT = tt;
    c.nGlobalV = gv;
    c.nGlobalT = gt;
    if (mg) {
        int err = 0;
        PRB_CUDA(cudaMemcpyAsync(&err, &((MgHeader*)c.mg.arena)->error, sizeof(int), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        if (err) { set_error("multi-GPU extraction: timed out waiting for a peer"); return PRB_ERR_CUDA; }
    }
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

}  // namespace prb
