// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked, imported or executed by the product
// (poissonrecon_gpu_b200/, include/, the CLI).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
//
// Sequential CPU restatement of the reference pipeline (DavidXu-JJ/PoissonRecon_GPU), stage by
// stage in the order of main.cu:3247-4573.  Each function cites the reference lines it
// restates.  Where the reference is racy or reads uninitialised memory the
// INTENDED semantics are restated and the deviation is named in the comment.
//
// Parity pinning: (1) B-spline tables are bit-identical to the reference's own host code
// (tests/golden/tables_d*.bin, made by oracle/_ref/ref_tables); (2) the GPU stages are pinned
// against dumps of the reference binary oracle/_ref/ref_poisson_d<D> run on the B200 box
// (tests/golden/ref_*.json digests; tools/ref_compare.py).  Keys are 64-bit here so that
// maxDepth 11/12 work (the reference itself stops at 9).
//
// Build: oracle/Makefile -> oracle/build/liborc.so (+ oracle/build/orc_cli).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>
#include "bspline_oracle.hpp"
#include "mc_case_table_oracle.h"

namespace orc {

typedef long long i64;

// ---------------------------------------------------------------- small tables
// LUTparent / LUTchild (main.cu:80-99) in closed form: per axis t = childbit + dir,
// parent dir = floor(t/2), child bit = t mod 2.  Child code c = x<<2|y<<1|z,
// neighbour slot j = 9(dx+1)+3(dy+1)+(dz+1).
static int LUTparent[8][27], LUTchild[8][27];
static void init_luts() {
    for (int c = 0; c < 8; c++)
        for (int j = 0; j < 27; j++) {
            int d[3] = {j / 9 - 1, (j / 3) % 3 - 1, j % 3 - 1};
            int b[3] = {(c >> 2) & 1, (c >> 1) & 1, c & 1};
            int pj = 0, cc = 0;
            for (int a = 0; a < 3; a++) {
                int t = b[a] + d[a];
                int pd = t < 0 ? -1 : (t > 1 ? 1 : 0);
                pj = pj * 3 + (pd + 1);
                cc = (cc << 1) | (t & 1);
            }
            LUTparent[c][j] = pj;
            LUTchild[c][j] = cc;
        }
}
// MarchingCubes.cuh tables.  Corner "ring" order (main.cu:2444-2455): 0:(0,0,0) 1:(1,0,0)
// 2:(1,1,0) 3:(0,1,0), 4-7 the same with z=1.
static int ring_index(int x, int y, int z) {
    int r = x | (z << 2);
    if (y) r += (r & 1) ? 1 : 3;
    return r;
}
static int mcTri[256][16], mcTriCount[256];
static int edgeVertex[12][2];   // MarchingCubes.cuh:693-706
static int faceEdges[6][4];     // MarchingCubes.cuh:734-741
static int parentFaceKind[8][6];  // MarchingCubes.cuh:708-717 (including its quirks, see below)
static const int childrenVertexKind[8] = {0, 1, 3, 2, 4, 5, 7, 6};  // MarchingCubes.cuh:721-723
// offset of edge e along axis a (-1 if a is the edge's own orientation). edgeKind =
// orientation<<2 | off0 | off1<<1, off0/off1 = the two other axes in increasing order
// (main.cu:1823-1840).
static int edge_off(int e, int a) {
    int o = e >> 2;
    if (a == o) return -1;
    int dim = 0;
    for (int k = 0; k < a; k++) if (k != o) dim++;
    return (e >> dim) & 1;
}
static void init_mc_tables() {
    for (int c = 0; c < 256; c++) {
        int n = 0;
        for (int k = 0; k < 16; k++) {
            char h = kMcCaseHex[16 * c + k];
            int v = (h == 'f') ? -1 : (h <= '9' ? h - '0' : h - 'a' + 10);
            mcTri[c][k] = v;
            if (v >= 0) n++;
        }
        mcTriCount[c] = n / 3;
    }
    for (int e = 0; e < 12; e++) {
        int o = e >> 2, r[2];
        for (int s = 0; s < 2; s++) {
            int xyz[3];
            for (int a = 0; a < 3; a++) xyz[a] = (a == o) ? s : edge_off(e, a);
            r[s] = ring_index(xyz[0], xyz[1], xyz[2]);
        }
        edgeVertex[e][0] = std::min(r[0], r[1]);
        edgeVertex[e][1] = std::max(r[0], r[1]);
    }
    for (int f = 0; f < 6; f++) {
        int a = f >> 1, s = f & 1, n = 0;
        for (int e = 0; e < 12; e++) if (edge_off(e, a) == s) faceEdges[f][n++] = e;
    }
    // The reference's table reads the child code as x=bit0,y=bit1,z=bit2 although node keys use
    // x=bit2 (main.cu:2881,2890) and row 7 has a stray 0 -- restated literally.
    for (int c = 0; c < 8; c++)
        for (int f = 0; f < 6; f++) parentFaceKind[c][f] = (((c >> (f >> 1)) & 1) == (f & 1)) ? f : -1;
    parentFaceKind[7][0] = 0;
}

struct PassInfo { std::string kind; int nv, nt; };

struct Oracle {
    int D = 0, N = 0;
    float center[3] = {0, 0, 0}, scale = 1;
    std::vector<float> P, Nr;          // sorted samples (normalised) and rescaled normals, 3N
    std::vector<int> sortedIdx;        // sorted -> original index
    std::vector<i64> sortedKey;        // Morton key per sorted sample
    int M = 0;
    std::vector<int> base, cnt;        // BaseAddressArray / NodeArrayCount (main.cu:772-776)
    std::vector<i64> key;
    std::vector<int> pidx, pnum, parent, didx, dnum, children, neighs, depthOf, p2n;
    std::vector<int> off;              // 3 per node: per-axis offset at its depth
    BSplineData bs;
    std::vector<float> V, divg, x, pointValue;
    std::vector<int> cgIters;
    std::vector<double> cgResidual;
    float iso = 0;
    // MC state
    std::vector<int> vOwner, vKind, vDepth;  // VertexArray
    std::vector<float> vPos, vvalue;
    std::vector<int> nodeVerts, nodeEdges, nodeFaces;   // 8/12/6 per node, 1-based (0 = unset)
    std::vector<int> eOwner, eKind;    // EdgeArray (depth D)
    std::vector<int> fOwner, fKind, fHasParent;  // FaceArray
    std::vector<int> hasSurf, hasTri, hasInter, subdivide;
    std::vector<float> meshV;
    std::vector<int> meshT;
    std::vector<PassInfo> passes;
    int finerDepth = 3;

    // -------------------------------------------------------------- A0  main.cu:530-571
    void normalise(const float* xyz, const float* nrm, int n, std::vector<float>& p, std::vector<float>& q) {
        N = n;
        float mn[3], mx[3];
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) {
                float v = xyz[3 * i + a];
                if (!i || v < mn[a]) mn[a] = v;
                if (!i || v > mx[a]) mx[a] = v;
            }
        scale = 1;
        for (int a = 0; a < 3; a++) {
            if (!a || scale < mx[a] - mn[a]) scale = float(mx[a] - mn[a]);
            center[a] = float(mx[a] + mn[a]) / 2;
        }
        scale *= 1.25f;
        for (int a = 0; a < 3; a++) center[a] -= scale / 2;
        p.resize(3 * (size_t)n);
        q.resize(3 * (size_t)n);
        for (int i = 0; i < n; i++) {
            for (int a = 0; a < 3; a++) p[3 * i + a] = (xyz[3 * i + a] - center[a]) / scale;
            float nx = nrm[3 * i], ny = nrm[3 * i + 1], nz = nrm[3 * i + 2];
            float sq = nx * nx + ny * ny + nz * nz;        // Geometry.inl:30 (float products, float sum)
            float len = float(std::sqrt((double)sq));     // Geometry.inl:33
            if (len > float(1e-6)) len = 1.0f / len;
            len *= (2 << D);
            q[3 * i] = nx * len; q[3 * i + 1] = ny * len; q[3 * i + 2] = nz * len;
        }
    }
    // -------------------------------------------------------------- A1  main.cu:132-164
    i64 encode(const float* p) const {
        float c[3] = {0.5f, 0.5f, 0.5f};
        float w = 0.25f;
        i64 k = 0;
        for (int i = D - 1; i >= 0; --i) {
            for (int a = 0; a < 3; a++) {
                if (p[a] > c[a]) { k |= 1ll << (3 * i + 2 - a); c[a] += w; }
                else c[a] -= w;
            }
            w /= 2;
        }
        return k;
    }
    // -------------------------------------------------------------- A1-A6  main.cu:583-814
    void build_octree(const std::vector<float>& p, const std::vector<float>& q) {
        int n = N;
        std::vector<i64> k(n);
        for (int i = 0; i < n; i++) k[i] = encode(&p[3 * i]);
        sortedIdx.resize(n);
        std::iota(sortedIdx.begin(), sortedIdx.end(), 0);
        // thrust::sort_by_key on (key<<32)+idx == stable sort by key  (main.cu:598-602)
        std::stable_sort(sortedIdx.begin(), sortedIdx.end(), [&](int a, int b) { return k[a] < k[b]; });
        P.resize(3 * (size_t)n); Nr.resize(3 * (size_t)n); sortedKey.resize(n);
        for (int i = 0; i < n; i++) {
            int s = sortedIdx[i];
            sortedKey[i] = k[s];
            for (int a = 0; a < 3; a++) { P[3 * i + a] = p[3 * s + a]; Nr[3 * i + a] = q[3 * s + a]; }
        }
        // unique leaves with first index / count (main.cu:608-640; hash tables replaced by the
        // sorted order, same result)
        std::vector<std::vector<i64>> nk(D + 1);          // non-empty node keys per depth
        std::vector<std::vector<int>> npidx(D + 1), npnum(D + 1);
        for (int i = 0; i < n; i++) {
            if (i == 0 || sortedKey[i] != sortedKey[i - 1]) { nk[D].push_back(sortedKey[i]); npidx[D].push_back(i); npnum[D].push_back(0); }
            npnum[D].back()++;
        }
        std::vector<std::vector<int>> prank(D + 1);       // rank of the parent among non-empty nodes of depth d-1
        for (int d = D; d >= 1; --d) {
            i64 mask = ~(7ll << (3 * (D - d)));            // clear this level's 3 bits (main.cu:293,315)
            prank[d].resize(nk[d].size());
            for (size_t r = 0; r < nk[d].size(); r++) {
                i64 fk = nk[d][r] & mask;
                if (nk[d - 1].empty() || nk[d - 1].back() != fk) { nk[d - 1].push_back(fk); npidx[d - 1].push_back(npidx[d][r]); npnum[d - 1].push_back(0); }
                npnum[d - 1].back() += npnum[d][r];
                prank[d][r] = (int)nk[d - 1].size() - 1;
            }
        }
        cnt.assign(D + 1, 0); base.assign(D + 2, 0);
        cnt[0] = 1;
        for (int d = 1; d <= D; d++) cnt[d] = 8 * (int)nk[d - 1].size();   // sibling groups of 8 (main.cu:226-249,336-359)
        for (int d = 1; d <= D + 1; d++) base[d] = base[d - 1] + cnt[d - 1];
        M = base[D + 1];
        key.assign(M, 0); pidx.assign(M, 0); pnum.assign(M, 0); parent.assign(M, -1); didx.assign(M, 0); dnum.assign(M, 0);
        children.assign(8 * (size_t)M, -1); neighs.assign(27 * (size_t)M, -1); depthOf.assign(M, 0); off.assign(3 * (size_t)M, 0);
        std::vector<std::vector<int>> slotOf(D + 1);      // global node index of non-empty node r at depth d
        slotOf[0].assign(1, 0);
        pnum[0] = n; pidx[0] = 0;
        for (int d = 1; d <= D; d++) {
            slotOf[d].resize(nk[d].size());
            for (size_t r = 0; r < nk[d].size(); r++) {
                int c = (int)((nk[d][r] >> (3 * (D - d))) & 7);
                int g = base[d] + 8 * prank[d][r] + c;
                slotOf[d][r] = g;
                pnum[g] = npnum[d][r];
                pidx[g] = npidx[d][r];
            }
        }
        for (int d = 0; d <= D; d++)
            for (int l = 0; l < cnt[d]; l++) depthOf[base[d] + l] = d;
        // keys, parents, children (intended semantics of main.cu:404-490; Q1/Q2/Q3 not reproduced)
        for (int k8 = 0; k8 < 8; k8++) children[k8] = 1 + k8;   // root: all eight depth-1 slots exist
        for (int d = 1; d <= D; d++) {
            for (size_t g = 0; g < nk[d - 1].size(); g++)
                for (int c = 0; c < 8; c++) {
                    int idx = base[d] + 8 * (int)g + c;
                    key[idx] = nk[d - 1][g] | ((i64)c << (3 * (D - d)));
                    parent[idx] = slotOf[d - 1][g];
                }
            if (d < D)
                for (size_t r = 0; r < nk[d].size(); r++)
                    for (int c = 0; c < 8; c++) children[8 * (size_t)slotOf[d][r] + c] = base[d + 1] + 8 * (int)r + c;
        }
        // didx / dnum: every depth-D slot counts 1 (main.cu:267-276); parents accumulate the
        // children whose dnum != 0 (main.cu:317-330)
        for (int l = 0; l < cnt[D]; l++) { didx[base[D] + l] = l; dnum[base[D] + l] = 1; }
        for (int d = D - 1; d >= 0; --d)
            for (int l = 0; l < cnt[d]; l++) {
                int i = base[d] + l;
                if (pnum[i] == 0) continue;
                int dn = 0, di = 0x7fffffff;
                for (int c = 0; c < 8; c++) {
                    int ch = children[8 * (size_t)i + c];
                    if (dnum[ch] != 0) { dn += dnum[ch]; di = std::min(di, didx[ch]); }
                }
                dnum[i] = dn; didx[i] = di;
            }
        // prefix pidx / didx inside every sibling group (main.cu:433-490)
        for (int d = 1; d <= D; d++)
            for (int g = 0; g < cnt[d] / 8; g++) {
                int i0 = base[d] + 8 * g, v = 0;
                while (pnum[i0 + v] == 0) v++;
                int nowP = pidx[i0 + v], nowD = didx[i0 + v];
                for (int j = 0; j < 8; j++) {
                    pidx[i0 + j] = nowP; nowP += pnum[i0 + j];
                    if (d != D) { didx[i0 + j] = nowD; nowD += dnum[i0 + j]; }
                }
            }
        // PointToNodeArrayD (main.cu:241-265)
        p2n.resize(n);
        for (size_t r = 0; r < nk[D].size(); r++)
            for (int t = 0; t < npnum[D][r]; t++) p2n[npidx[D][r] + t] = slotOf[D][r] - base[D];
        // per-axis offsets (main.cu:861-887)
        for (int i = 0; i < M; i++) {
            int d = depthOf[i];
            for (int l = 1; l <= d; l++) {
                int c = (int)((key[i] >> (3 * (D - l))) & 7);
                off[3 * (size_t)i + 0] |= ((c >> 2) & 1) << (d - l);
                off[3 * (size_t)i + 1] |= ((c >> 1) & 1) << (d - l);
                off[3 * (size_t)i + 2] |= (c & 1) << (d - l);
            }
        }
        // neighbours (main.cu:492-509, 803-814)
        neighs[13] = 0;
        for (int d = 1; d <= D; d++)
            for (int l = 0; l < cnt[d]; l++) {
                int i = base[d] + l;
                int c = (int)((key[i] >> (3 * (D - d))) & 7);
                int pa = parent[i];
                for (int j = 0; j < 27; j++) {
                    int np = neighs[27 * (size_t)pa + LUTparent[c][j]];
                    neighs[27 * (size_t)i + j] = (np != -1) ? children[8 * (size_t)np + LUTchild[c][j]] : -1;
                }
            }
    }
    int fidx(int node, int a) const { return ((1 << depthOf[node]) - 1) + off[3 * (size_t)node + a]; }

    // -------------------------------------------------------------- A7  main.cu:913-964
    void splat() {
        int MD = cnt[D], b0 = base[D];
        V.assign(3 * (size_t)MD, 0.f);
        float width = float(1.0 / (1 << D));
#pragma omp parallel for schedule(dynamic, 256)
        for (int l = 0; l < MD; l++) {
            int i = b0 + l;
            float oc[3];
            for (int a = 0; a < 3; a++) oc[a] = float((0.5 + off[3 * (size_t)i + a]) * width);   // BinaryNode.cuh:46-50
            float val[3] = {0, 0, 0};
            for (int j = 0; j < 27; j++) {
                int nb = neighs[27 * (size_t)i + j];
                if (nb == -1) continue;
                for (int k = 0; k < pnum[nb]; k++) {
                    int pi = pidx[nb] + k;
                    float w3[3];
                    for (int a = 0; a < 3; a++) {
                        PPoly f = confirmed_shift(bs.maxDepthFunction, P[3 * (size_t)pi + a]);
                        w3[a] = confirmed_value(f, oc[a]);
                    }
                    float weight = w3[0] * w3[1] * w3[2];
                    for (int a = 0; a < 3; a++) val[a] = std::fmaf(weight, Nr[3 * (size_t)pi + a], val[a]);
                }
            }
            for (int a = 0; a < 3; a++) V[3 * (size_t)l + a] += val[a];
        }
    }
    // per-depth caches of the reference's 1-D tables (literal table(a,b) lookups, memoised)
    std::vector<std::vector<double>> dfRow;   // [d][off_o*3k + t], t = off_s - k*(off_o-1), k = 2^(D-d)
    std::vector<std::vector<double>> ffSame, d2Same;   // [d][off_o*3 + (delta+1)] : table index fi_o*res+fi_n
    void build_table_caches() {
        int res = bs.res;
        (void)res;
        dfRow.assign(D + 1, {}); ffSame.assign(D + 1, {}); d2Same.assign(D + 1, {});
        for (int d = 0; d <= D; d++) {
            int nd = 1 << d, k = 1 << (D - d), nD = 1 << D;
            dfRow[d].assign((size_t)nd * 3 * k, 0.0);
            ffSame[d].assign((size_t)nd * 3, 0.0);
            d2Same[d].assign((size_t)nd * 3, 0.0);
#pragma omp parallel for schedule(dynamic, 4)
            for (int o = 0; o < nd; o++) {
                int fo = (nd - 1) + o;
                for (int t = 0; t < 3 * k; t++) {
                    int s = k * (o - 1) + t;
                    if (s < 0 || s >= nD) continue;
                    dfRow[d][(size_t)o * 3 * k + t] = bs.table(1, fo, (nD - 1) + s);   // dot_F_DF[idxO_1 + idxO_2*res], main.cu:1046
                }
                for (int dl = -1; dl <= 1; dl++) {
                    int nb = o + dl;
                    if (nb < 0 || nb >= nd) continue;
                    int fn = (nd - 1) + nb;
                    ffSame[d][o * 3 + dl + 1] = bs.table(0, fn, fo);    // dot_F_F[idxO_1*res + idxO_2], main.cu:1200
                    d2Same[d][o * 3 + dl + 1] = bs.table(2, fn, fo);
                }
            }
        }
    }
    // -------------------------------------------------------------- A8  main.cu:1007-1141, 3383-3462
    void divergence() {
        divg.assign(M, 0.f);
        int k;
#pragma omp parallel for schedule(dynamic, 64) private(k)
        for (int i = 0; i < M; i++) {
            int d = depthOf[i];
            k = 1 << (D - d);
            const std::vector<double>& row = dfRow[d];
            double val = 0;
            for (int j = 0; j < 27; j++) {
                int nb = neighs[27 * (size_t)i + j];
                if (nb == -1) continue;
                for (int q = 0; q < dnum[nb]; q++) {
                    int s = didx[nb] + q;             // depth-D slot (local)
                    int sg = base[D] + s;
                    float uo[3];
                    for (int a = 0; a < 3; a++) {
                        int o = off[3 * (size_t)i + a];
                        int t = off[3 * (size_t)sg + a] - k * (o - 1);
                        uo[a] = (float)row[(size_t)o * 3 * k + t];          // main.cu:1050-1053 (double -> float)
                    }
                    float dp = 0;                                             // DotProduct, main.cu:966-972
                    for (int a = 0; a < 3; a++) dp = std::fmaf(V[3 * (size_t)s + a], uo[a], dp);
                    val += dp;     // depth >= 5: double accumulate (main.cu:1054); depth 0-4: the reference sums the
                                   // same float terms with thrust::reduce in unspecified order (main.cu:3449)
                }
            }
            divg[i] = (float)val;
        }
    }
    // -------------------------------------------------------------- A9  main.cu:1143-1213
    float lap_entry(int o, int nb) const {
        int d = depthOf[o];
        double ff[3], d2[3];
        for (int a = 0; a < 3; a++) {
            int oo = off[3 * (size_t)o + a], dl = off[3 * (size_t)nb + a] - oo;
            ff[a] = ffSame[d][oo * 3 + dl + 1];
            d2[a] = d2Same[d][oo * 3 + dl + 1];
        }
        double e = d2[0] * ff[1] * ff[2] + d2[1] * ff[0] * ff[2] + d2[2] * ff[0] * ff[1];
        return (float)e;
    }
    // -------------------------------------------------------------- A10  CG_CUDA.cuh:186-324, 344-347
    void solve() {
        x.assign(M, 0.f);
        cgIters.assign(D + 1, 0); cgResidual.assign(D + 1, 0);
        for (int d = 0; d <= D; d++) {
            int n = cnt[d], b0 = base[d];
            // CSR in neighbour order, entries with |v| <= 1e-6 dropped (main.cu:1205)
            std::vector<int> rowptr(n + 1, 0), col;
            std::vector<float> val;
            col.reserve(27 * (size_t)n); val.reserve(27 * (size_t)n);
            for (int r = 0; r < n; r++) {
                for (int j = 0; j < 27; j++) {
                    int nb = neighs[27 * (size_t)(b0 + r) + j];
                    if (nb == -1) continue;
                    float v = lap_entry(b0 + r, nb);
                    if (std::fabs((double)v) > (double)float(1e-6)) { col.push_back(nb - b0); val.push_back(v); }
                }
                rowptr[r + 1] = (int)col.size();
            }
            std::vector<float> r(n), p(n), Ax(n);
            float* xx = &x[b0];
            auto spmv = [&](const float* in, float* out) {
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; i++) {
                    float o = 0.0f;
                    for (int j = rowptr[i]; j < rowptr[i + 1]; j++) o = std::fmaf(val[j], in[col[j]], o);   // CG_CUDA.cuh:194-199
                    out[i] = o;
                }
            };
            auto dot = [&](const float* a, const float* b) {
                double s = 0;                                                 // CG_CUDA.cuh:217-220
                for (int i = 0; i < n; i++) s += (double)(a[i] * b[i]);
                return s;
            };
            for (int i = 0; i < n; i++) { r[i] = divg[b0 + i]; xx[i] = 0; }
            const float tol = 1e-5f;
            spmv(xx, Ax.data());
            for (int i = 0; i < n; i++) r[i] = std::fmaf(-1.0f, Ax[i], r[i]);
            float r1 = (float)dot(r.data(), r.data()), r0 = 0, a, b, na;
            int k = 1;
            while (r1 > tol * tol && k <= 10000) {
                if (k > 1) {
                    b = r1 / r0;
                    for (int i = 0; i < n; i++) p[i] = r[i] + b * p[i];       // CG_CUDA.cuh:249-254 with a = 1
                } else {
                    for (int i = 0; i < n; i++) p[i] = r[i];
                }
                spmv(p.data(), Ax.data());
                double dd = dot(p.data(), Ax.data());
                a = (float)((double)r1 / dd);
                for (int i = 0; i < n; i++) xx[i] = std::fmaf(a, p[i], xx[i]);
                na = -a;
                for (int i = 0; i < n; i++) r[i] = std::fmaf(na, Ax[i], r[i]);
                r0 = r1;
                r1 = (float)dot(r.data(), r.data());
                k++;
            }
            cgIters[d] = k - 1;
            cgResidual[d] = std::sqrt((double)r1);
        }
    }
    // value of node nb's basis function at pos (ConfirmedPPolynomial.cuh:79-91 on baseFunctions)
    inline float node_value(int nb, const float* pos) const {
        float vx = confirmed_value(bs.baseFunctions[fidx(nb, 0)], pos[0]);
        float vy = confirmed_value(bs.baseFunctions[fidx(nb, 1)], pos[1]);
        float vz = confirmed_value(bs.baseFunctions[fidx(nb, 2)], pos[2]);
        return x[nb] * vx * vy * vz;
    }
    // val += d_x*vx*vy*vz with the final multiply fused into the add (nvcc -fmad default)
    inline void accum_node(float& val, int nb, const float* pos) const {
        float vx = confirmed_value(bs.baseFunctions[fidx(nb, 0)], pos[0]);
        float vy = confirmed_value(bs.baseFunctions[fidx(nb, 1)], pos[1]);
        float vz = confirmed_value(bs.baseFunctions[fidx(nb, 2)], pos[2]);
        val = std::fmaf(x[nb] * vx * vy, vz, val);
    }
    // -------------------------------------------------------------- A11  main.cu:1334-1381, 3480-3496
    void iso_value() {
        pointValue.assign(N, 0.f);
#pragma omp parallel for schedule(dynamic, 1024)
        for (int i = 0; i < N; i++) {
            int now = base[D] + p2n[i];
            float val = 0.0f;
            const float* pos = &P[3 * (size_t)i];
            while (now != -1) {
                for (int j = 0; j < 27; j++) {
                    int nb = neighs[27 * (size_t)now + j];
                    if (nb != -1) accum_node(val, nb, pos);
                }
                now = parent[now];
            }
            pointValue[i] = val;
        }
        // thrust::reduce(float) has no defined order; restated as a double sum rounded to float
        double s = 0;
        for (int i = 0; i < N; i++) s += pointValue[i];
        iso = (float)s;
        iso /= N;
    }

    // -------------------------------------------------------------- A12 topology  main.cu:1423-1681, 1795-2082, 2827-2955
    static void corner_dirs(int j, int s[3]) { s[0] = (j & 1) ? 1 : -1; s[1] = (j & 2) ? 1 : -1; s[2] = (j & 4) ? 1 : -1; }
    void build_vertices() {
        vOwner.clear(); vKind.clear(); vDepth.clear(); vPos.clear();
        nodeVerts.assign(8 * (size_t)M, 0);
        for (int i = 0; i < M; i++) {
            int d = depthOf[i];
            for (int j = 0; j < 8; j++) {
                int s[3];
                corner_dirs(j, s);
                i64 bestKey = 0x7fffffffffffffffll;
                int best = -1;
                for (int m = 0; m < 8; m++) {
                    int dx = (m & 1) ? s[0] : 0, dy = (m & 2) ? s[1] : 0, dz = (m & 4) ? s[2] : 0;
                    int nb = neighs[27 * (size_t)i + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
                    if (nb != -1 && key[nb] < bestKey) { bestKey = key[nb]; best = nb; }
                }
                if (best == i && i > 0) {            // validVertex: ownerNodeIdx > 0 (main.cu:1634-1638)
                    vOwner.push_back(i); vKind.push_back(j); vDepth.push_back(d);
                    float w = 1.0f / (1 << d);
                    for (int a = 0; a < 3; a++) vPos.push_back((off[3 * (size_t)i + a] + ((j >> a) & 1)) * w);
                }
            }
        }
        // back pointers node -> vertex (main.cu:1640-1681)
        for (size_t v = 0; v < vOwner.size(); v++) {
            int i = vOwner[v], j = vKind[v], s[3];
            corner_dirs(j, s);
            for (int m = 0; m < 8; m++) {
                int dx = (m & 1) ? s[0] : 0, dy = (m & 2) ? s[1] : 0, dz = (m & 4) ? s[2] : 0;
                int nb = neighs[27 * (size_t)i + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
                if (nb == -1) continue;
                int cx = (j & 1) ^ (m & 1), cy = ((j >> 1) & 1) ^ ((m >> 1) & 1), cz = ((j >> 2) & 1) ^ ((m >> 2) & 1);
                nodeVerts[8 * (size_t)nb + ring_index(cx, cy, cz)] = (int)v + 1;
            }
        }
    }
    void build_edges() {
        eOwner.clear(); eKind.clear();
        nodeEdges.assign(12 * (size_t)M, 0);
        for (int l = 0; l < cnt[D]; l++) {
            int i = base[D] + l;
            for (int e = 0; e < 12; e++) {
                int o = e >> 2, ax[2], n = 0;
                for (int a = 0; a < 3; a++) if (a != o) ax[n++] = a;
                int sg[2] = {(e & 1) ? 1 : -1, (e & 2) ? 1 : -1};
                i64 bestKey = 0x7fffffffffffffffll;
                int best = -1;
                for (int m = 0; m < 4; m++) {
                    int dd[3] = {0, 0, 0};
                    if (m & 1) dd[ax[0]] = sg[0];
                    if (m & 2) dd[ax[1]] = sg[1];
                    int nb = neighs[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
                    if (nb != -1 && key[nb] < bestKey) { bestKey = key[nb]; best = nb; }
                }
                if (best == i && i > 0) { eOwner.push_back(i); eKind.push_back(e); }
            }
        }
        for (size_t q = 0; q < eOwner.size(); q++) {
            int i = eOwner[q], e = eKind[q];
            int o = e >> 2, ax[2], n = 0;
            for (int a = 0; a < 3; a++) if (a != o) ax[n++] = a;
            int sg[2] = {(e & 1) ? 1 : -1, (e & 2) ? 1 : -1};
            for (int m = 0; m < 4; m++) {
                int dd[3] = {0, 0, 0};
                if (m & 1) dd[ax[0]] = sg[0];
                if (m & 2) dd[ax[1]] = sg[1];
                int nb = neighs[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
                if (nb == -1) continue;
                int b0 = (e & 1) ^ (m & 1), b1 = ((e >> 1) & 1) ^ ((m >> 1) & 1);
                nodeEdges[12 * (size_t)nb + ((o << 2) | b0 | (b1 << 1))] = (int)q + 1;
            }
        }
    }
    void build_faces() {
        fOwner.clear(); fKind.clear(); fHasParent.clear();
        nodeFaces.assign(6 * (size_t)M, 0);
        for (int i = 0; i < M; i++) {
            int d = depthOf[i];
            for (int f = 0; f < 6; f++) {
                int o = f >> 1, dd[3] = {0, 0, 0};
                dd[o] = (f & 1) ? 1 : -1;
                int nb = neighs[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
                int best = i;
                if (nb != -1 && key[nb] < key[i]) best = nb;
                if (best == i) {                         // validFace: ownerNodeIdx >= 0, root included (main.cu:2902-2906)
                    fOwner.push_back(i); fKind.push_back(f);
                    int sonKey = (int)((key[i] >> (3 * (D - d))) & 7);
                    fHasParent.push_back(parent[i] == -1 ? -1 : (parentFaceKind[sonKey][f] != -1 ? 1 : -1));
                }
            }
        }
        for (size_t q = 0; q < fOwner.size(); q++) {
            int i = fOwner[q], f = fKind[q], o = f >> 1, dd[3] = {0, 0, 0};
            dd[o] = (f & 1) ? 1 : -1;
            nodeFaces[6 * (size_t)i + f] = (int)q + 1;
            int nb = neighs[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
            if (nb != -1) nodeFaces[6 * (size_t)nb + (f ^ 1)] = (int)q + 1;
        }
    }
    // -------------------------------------------------------------- main.cu:2259-2326
    void vertex_values() {
        size_t nv = vOwner.size();
        vvalue.assign(nv, 0.f);
#pragma omp parallel for schedule(dynamic, 512)
        for (size_t v = 0; v < nv; v++) {
            int depth = vDepth[v];
            float val = 0.0f;
            const float* pos = &vPos[3 * v];
            int exceed = childrenVertexKind[vKind[v]];
            int now = vOwner[v];
            while (now != -1) {
                for (int k = 0; k < 27; k++) { int nb = neighs[27 * (size_t)now + k]; if (nb != -1) accum_node(val, nb, pos); }
                now = parent[now];
            }
            now = vOwner[v];
            while (depth < D) {
                ++depth;
                now = children[8 * (size_t)now + exceed];
                if (now == -1) break;
                for (int k = 0; k < 27; k++) { int nb = neighs[27 * (size_t)now + k]; if (nb != -1) accum_node(val, nb, pos); }
            }
            vvalue[v] = val - iso;
        }
    }
    // main.cu:2584-2597 as compiled with FMA contraction
    static void interpolate(const float* p1, const float* p2, int dim, float v1, float v2, float* out) {
        for (int a = 0; a < 3; a++) if (a != dim) out[a] = p1[a];
        float pivot = v1 / (v1 - v2);
        float another = 1 - pivot;
        out[dim] = std::fmaf(p2[dim], pivot, p1[dim] * another);
    }
    void insert_mesh(const char* kind, const std::vector<float>& vb, const std::vector<int>& tb) {   // main.cu:3220-3245
        int prev = (int)(meshV.size() / 3);
        meshV.insert(meshV.end(), vb.begin(), vb.end());
        for (int t : tb) meshT.push_back(t + prev);
        passes.push_back({kind, (int)(vb.size() / 3), (int)(tb.size() / 3)});
    }
    // -------------------------------------------------------------- depth-D pass  main.cu:3652-3795
    void mc_main_pass() {
        size_t ne = eOwner.size();
        std::vector<int> vexNums(ne, 0), vexAddr(ne, 0);
        for (size_t q = 0; q < ne; q++) {                  // generateVexNums main.cu:2457-2479
            int o = eOwner[q], e = eKind[q];
            int v1 = nodeVerts[8 * (size_t)o + edgeVertex[e][0]] - 1, v2 = nodeVerts[8 * (size_t)o + edgeVertex[e][1]] - 1;
            if (vvalue[v1] * vvalue[v2] <= 0) vexNums[q] = 1;
        }
        int allVex = 0;
        for (size_t q = 0; q < ne; q++) { vexAddr[q] = allVex; allVex += vexNums[q]; }
        int MD = cnt[D], b0 = base[D];
        std::vector<int> triNums(MD), cat(MD), triAddr(MD);
        int allTri = 0;
        for (int l = 0; l < MD; l++) {                     // generateTriNums main.cu:2540-2560
            int c = 0;
            for (int j = 0; j < 8; j++) if (vvalue[nodeVerts[8 * (size_t)(b0 + l) + j] - 1] < 0) c |= 1 << j;
            cat[l] = c; triNums[l] = mcTriCount[c]; triAddr[l] = allTri; allTri += triNums[l];
        }
        std::vector<float> vb(3 * (size_t)allVex);
        for (size_t q = 0; q < ne; q++) {                  // generateIntersectionPoint main.cu:2599-2627
            if (!vexNums[q]) continue;
            int o = eOwner[q], e = eKind[q];
            int v1 = nodeVerts[8 * (size_t)o + edgeVertex[e][0]] - 1, v2 = nodeVerts[8 * (size_t)o + edgeVertex[e][1]] - 1;
            interpolate(&vPos[3 * (size_t)v1], &vPos[3 * (size_t)v2], e >> 2, vvalue[v1], vvalue[v2], &vb[3 * (size_t)vexAddr[q]]);
        }
        std::vector<int> tb(3 * (size_t)allTri);
        hasSurf.assign(fOwner.size(), 0);
        for (int l = 0; l < MD; l++) {                     // generateTrianglePos main.cu:2699-2757
            int i = b0 + l, c = cat[l];
            int ehv[12] = {0};
            for (int j = 0; j < 3 * triNums[l]; j++) {
                int e = mcTri[c][j];
                ehv[e] = 1;
                tb[3 * (size_t)triAddr[l] + j] = vexAddr[nodeEdges[12 * (size_t)i + e] - 1];
            }
            for (int f = 0; f < 6; f++) {
                int mark = 0;
                for (int k = 0; k < 4; k++) mark |= ehv[faceEdges[f][k]];
                if (!mark) continue;
                int pn = parent[i], nf = nodeFaces[6 * (size_t)i + f] - 1;
                hasSurf[nf] = 1;
                while (fHasParent[nf] != -1) {
                    nf = nodeFaces[6 * (size_t)pn + f] - 1;
                    pn = parent[pn];
                    hasSurf[nf] = 1;
                }
            }
        }
        insert_mesh("main", vb, tb);
    }
    // -------------------------------------------------------------- main.cu:2957-2992, 3799-3836
    void find_subdivide() {
        hasTri.assign(M, 0); hasInter.assign(M, 0); subdivide.clear();
        for (int i = 0; i < base[D]; i++) {
            // Q7: node 0 has vertices[] == 0 -> the reference reads vvalue[-1]; the root always has
            // children so it is never selected; skip it.
            if (i == 0) continue;
            int ht = 0;
            int sign = (vvalue[nodeVerts[8 * (size_t)i] - 1] < 0) ? -1 : 1;
            for (int j = 1; j < 8; j++) if (sign * vvalue[nodeVerts[8 * (size_t)i + j] - 1] < 0) { ht = 1; break; }
            int hi = 0;
            for (int f = 0; f < 6; f++) if (hasSurf[nodeFaces[6 * (size_t)i + f] - 1]) { hi = 1; break; }
            hasTri[i] = ht; hasInter[i] = hi;
            if (children[8 * (size_t)i] == -1 && (ht || hi)) subdivide.push_back(i);
        }
    }
    // -------------------------------------------------------------- refinement passes  main.cu:3886-4561
    // One pass = a set of roots (empty leaves of the real tree, all at the same depth for the
    // batched passes; a single root for the coarse passes) expanded to complete subtrees down to
    // depth D.  Virtual nodes are indexed M + local like the reference (NodeArray_sz + idx).
    void refine_pass(const std::vector<int>& roots, const char* kind, bool singleRootMode) {
        if (roots.empty()) { if (!singleRootMode) passes.push_back({kind, 0, 0}); return; }
        int rd = depthOf[roots[0]];
        int nr = (int)roots.size();
        // layout per depth: depthAddr[d] + r*8^(d-rd) + local  (wholeRebuildArray main.cu:3032-3101 /
        // singleRebuildArray main.cu:3114-3155 give the same order for a single root)
        std::vector<i64> depthAddr(D + 2, 0), per(D + 1, 0);
        i64 total = 0;
        for (int d = rd; d <= D; d++) { per[d] = 1ll << (3 * (d - rd)); depthAddr[d] = total; total += per[d] * nr; }
        std::vector<i64> vkey(total);
        std::vector<int> vparent(total), vdepth(total), vreplaced(total, -1);
        std::vector<char> visroot(total, 0);
        std::vector<int> vchildren(8 * (size_t)total, -1), vneigh(27 * (size_t)total, -1);
        std::vector<int> voff(3 * (size_t)total);
        std::vector<int> savedChild(nr);
        for (int r = 0; r < nr; r++) {
            int root = roots[r];
            i64 idx = depthAddr[rd] + r;
            vkey[idx] = key[root]; vparent[idx] = parent[root]; vdepth[idx] = rd; vreplaced[idx] = root; visroot[idx] = 1;
            int sonKey = (int)((key[root] >> (3 * (D - rd))) & 7);
            savedChild[r] = children[8 * (size_t)parent[root] + sonKey];
            children[8 * (size_t)parent[root] + sonKey] = M + (int)idx;      // main.cu:3069 / 3922
            for (int d = rd + 1; d <= D; d++)
                for (i64 l = 0; l < per[d]; l++) {
                    i64 me = depthAddr[d] + r * per[d] + l, pa = depthAddr[d - 1] + r * per[d - 1] + (l >> 3);
                    vparent[me] = M + (int)pa;
                    vkey[me] = vkey[pa] | ((i64)(l & 7) << (3 * (D - d)));
                    vdepth[me] = d; vreplaced[me] = root;
                    vchildren[8 * (size_t)pa + (l & 7)] = M + (int)me;
                }
        }
        for (i64 i = 0; i < total; i++) {
            int d = vdepth[i];
            for (int a = 0; a < 3; a++) voff[3 * (size_t)i + a] = 0;
            for (int l = 1; l <= d; l++) {
                int c = (int)((vkey[i] >> (3 * (D - l))) & 7);
                voff[3 * (size_t)i + 0] |= ((c >> 2) & 1) << (d - l);
                voff[3 * (size_t)i + 1] |= ((c >> 1) & 1) << (d - l);
                voff[3 * (size_t)i + 2] |= (c & 1) << (d - l);
            }
        }
        // computeRebuildNeighbor main.cu:3188-3217, level by level from the roots' depth
        for (int d = rd; d <= D; d++)
            for (i64 l = 0; l < per[d] * nr; l++) {
                i64 i = depthAddr[d] + l;
                int c = (int)((vkey[i] >> (3 * (D - d))) & 7);
                int pa = vparent[i];
                for (int j = 0; j < 27; j++) {
                    int np = (pa < M) ? neighs[27 * (size_t)pa + LUTparent[c][j]] : vneigh[27 * (size_t)(pa - M) + LUTparent[c][j]];
                    int r;
                    if (np == -1) r = -1;
                    else if (np < M) r = children[8 * (size_t)np + LUTchild[c][j]];
                    else r = vchildren[8 * (size_t)(np - M) + LUTchild[c][j]];
                    vneigh[27 * (size_t)i + j] = r;
                }
            }
        // depth-D virtual cells: vertices / edges owned by the min-key VIRTUAL incident cell
        // (initSubdivideVertexOwner main.cu:1568-1632, initSubdivideEdgeArray 1945-2019)
        i64 nD = per[D] * nr, a0 = depthAddr[D];
        std::vector<int> sOwner, sKind;
        std::vector<float> sPos;
        std::vector<int> sVerts(8 * (size_t)nD, 0), sEdges(12 * (size_t)nD, 0);
        float w = 1.0f / (1 << D);
        for (i64 l = 0; l < nD; l++) {
            i64 i = a0 + l;
            for (int j = 0; j < 8; j++) {
                int s[3];
                corner_dirs(j, s);
                i64 bestKey = 0x7fffffffffffffffll;
                int best = -1;
                for (int m = 0; m < 8; m++) {
                    int dx = (m & 1) ? s[0] : 0, dy = (m & 2) ? s[1] : 0, dz = (m & 4) ? s[2] : 0;
                    int nb = vneigh[27 * (size_t)i + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
                    if (nb == -1 || nb < M) continue;
                    if (vkey[nb - M] < bestKey) { bestKey = vkey[nb - M]; best = nb; }
                }
                if (best == M + (int)i) {
                    sOwner.push_back(best); sKind.push_back(j);
                    for (int a = 0; a < 3; a++) sPos.push_back((voff[3 * (size_t)i + a] + ((j >> a) & 1)) * w);
                }
            }
        }
        for (size_t v = 0; v < sOwner.size(); v++) {      // maintainSubdivideVertexNodePointer main.cu:1743-1793
            i64 i = sOwner[v] - M;
            int j = sKind[v], s[3];
            corner_dirs(j, s);
            for (int m = 0; m < 8; m++) {
                int dx = (m & 1) ? s[0] : 0, dy = (m & 2) ? s[1] : 0, dz = (m & 4) ? s[2] : 0;
                int nb = vneigh[27 * (size_t)i + 9 * (dx + 1) + 3 * (dy + 1) + (dz + 1)];
                if (nb == -1 || nb < M) continue;
                int cx = (j & 1) ^ (m & 1), cy = ((j >> 1) & 1) ^ ((m >> 1) & 1), cz = ((j >> 2) & 1) ^ ((m >> 2) & 1);
                sVerts[8 * (size_t)(nb - M - a0) + ring_index(cx, cy, cz)] = (int)v + 1;
            }
        }
        std::vector<int> seOwner, seKind;
        for (i64 l = 0; l < nD; l++) {
            i64 i = a0 + l;
            for (int e = 0; e < 12; e++) {
                int o = e >> 2, ax[2], n = 0;
                for (int a = 0; a < 3; a++) if (a != o) ax[n++] = a;
                int sg[2] = {(e & 1) ? 1 : -1, (e & 2) ? 1 : -1};
                i64 bestKey = 0x7fffffffffffffffll;
                int best = -1;
                for (int m = 0; m < 4; m++) {
                    int dd[3] = {0, 0, 0};
                    if (m & 1) dd[ax[0]] = sg[0];
                    if (m & 2) dd[ax[1]] = sg[1];
                    int nb = vneigh[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
                    if (nb == -1 || nb < M) continue;
                    if (vkey[nb - M] < bestKey) { bestKey = vkey[nb - M]; best = nb; }
                }
                if (best == M + (int)i) { seOwner.push_back(best); seKind.push_back(e); }
            }
        }
        for (size_t q = 0; q < seOwner.size(); q++) {     // maintainSubdivideEdgeNodePointer main.cu:2155-2222
            i64 i = seOwner[q] - M;
            int e = seKind[q], o = e >> 2, ax[2], n = 0;
            for (int a = 0; a < 3; a++) if (a != o) ax[n++] = a;
            int sg[2] = {(e & 1) ? 1 : -1, (e & 2) ? 1 : -1};
            for (int m = 0; m < 4; m++) {
                int dd[3] = {0, 0, 0};
                if (m & 1) dd[ax[0]] = sg[0];
                if (m & 2) dd[ax[1]] = sg[1];
                int nb = vneigh[27 * (size_t)i + 9 * (dd[0] + 1) + 3 * (dd[1] + 1) + (dd[2] + 1)];
                if (nb == -1 || nb < M) continue;
                int b0 = (e & 1) ^ (m & 1), b1 = ((e >> 1) & 1) ^ ((m >> 1) & 1);
                sEdges[12 * (size_t)(nb - M - a0) + ((o << 2) | b0 | (b1 << 1))] = (int)q + 1;
            }
        }
        // corner values: only REAL nodes contribute; virtual roots stand for the real node they
        // replace (computeSubdivideVertexImplicitFunctionValue main.cu:2328-2442)
        std::vector<float> sval(sOwner.size());
#pragma omp parallel for schedule(dynamic, 256)
        for (size_t v = 0; v < sOwner.size(); v++) {
            float val = 0.0f;
            const float* pos = &sPos[3 * v];
            int now = sOwner[v];
            while (now != -1) {
                for (int k = 0; k < 27; k++) {
                    int nb = (now < M) ? neighs[27 * (size_t)now + k] : vneigh[27 * (size_t)(now - M) + k];
                    if (nb == -1) continue;
                    if (nb >= M && visroot[nb - M]) nb = vreplaced[nb - M];
                    if (nb >= M) continue;
                    accum_node(val, nb, pos);
                }
                now = (now < M) ? parent[now] : vparent[now - M];
            }
            sval[v] = val - iso;
        }
        size_t ne = seOwner.size();
        std::vector<int> vexNums(ne, 0), vexAddr(ne, 0);
        int allVex = 0;
        for (size_t q = 0; q < ne; q++) {
            i64 o = seOwner[q] - M - a0;
            int e = seKind[q];
            int v1 = sVerts[8 * (size_t)o + edgeVertex[e][0]] - 1, v2 = sVerts[8 * (size_t)o + edgeVertex[e][1]] - 1;
            if (sval[v1] * sval[v2] <= 0) vexNums[q] = 1;
            vexAddr[q] = allVex; allVex += vexNums[q];
        }
        bool emit = true;
        if (singleRootMode && allVex == 0) emit = false;     // main.cu:4095-4103
        if (emit) {
            std::vector<int> triNums(nD), cat(nD), triAddr(nD);
            int allTri = 0;
            for (i64 l = 0; l < nD; l++) {
                int c = 0;
                for (int j = 0; j < 8; j++) if (sval[sVerts[8 * (size_t)l + j] - 1] < 0) c |= 1 << j;
                cat[l] = c; triNums[l] = mcTriCount[c]; triAddr[l] = allTri; allTri += triNums[l];
            }
            std::vector<float> vb(3 * (size_t)allVex);
            for (size_t q = 0; q < ne; q++) {
                if (!vexNums[q]) continue;
                i64 o = seOwner[q] - M - a0;
                int e = seKind[q];
                int v1 = sVerts[8 * (size_t)o + edgeVertex[e][0]] - 1, v2 = sVerts[8 * (size_t)o + edgeVertex[e][1]] - 1;
                interpolate(&sPos[3 * (size_t)v1], &sPos[3 * (size_t)v2], e >> 2, sval[v1], sval[v2], &vb[3 * (size_t)vexAddr[q]]);
            }
            std::vector<int> tb(3 * (size_t)allTri);
            for (i64 l = 0; l < nD; l++)
                for (int j = 0; j < 3 * triNums[l]; j++) tb[3 * (size_t)triAddr[l] + j] = vexAddr[sEdges[12 * (size_t)l + mcTri[cat[l]][j]] - 1];
            insert_mesh(kind, vb, tb);
        }
        // The coarse loop restores the parent's child pointer (main.cu:4096,4191); the batched
        // passes leave it dangling (never read again).  Restored here in both cases.
        for (int r = 0; r < nr; r++) {
            int root = roots[r];
            int sonKey = (int)((key[root] >> (3 * (D - rd))) & 7);
            children[8 * (size_t)parent[root] + sonKey] = savedChild[r];
        }
    }
    void refine() {
        // coarse roots (depth < finerDepth), one at a time, in node order (main.cu:3887-4202)
        size_t q = 0;
        for (; q < subdivide.size(); q++) {
            int r = subdivide[q];
            if (depthOf[r] >= finerDepth) break;
            refine_pass({r}, "coarse", true);
        }
        // batched per depth (main.cu:4211-4561)
        for (int d = finerDepth; d < D; d++) {
            std::vector<int> roots;
            for (int r : subdivide) if (depthOf[r] == d) roots.push_back(r);
            refine_pass(roots, "finer", false);
        }
    }

    int run(const float* xyz, const float* nrm, int n, int depth, int stages) {
        D = depth;
        std::vector<float> p, q;
        normalise(xyz, nrm, n, p, q);
        build_octree(p, q);
        if (stages < 2) return 0;
        bs.set(D);
        build_table_caches();
        splat();
        divergence();
        if (stages < 3) return 0;
        solve();
        iso_value();
        if (stages < 4) return 0;
        build_vertices(); build_edges(); build_faces();
        vertex_values();
        meshV.clear(); meshT.clear(); passes.clear();
        mc_main_pass();
        find_subdivide();
        if (stages == 40) return 0;      // everything except the refinement passes (bench.py: bounded CPU sample at full size)
        refine();
        return 0;
    }
};

}  // namespace orc

// ------------------------------------------------------------------ C interface (ctypes)
using namespace orc;
static bool g_init = false;
extern "C" {
void* orc_create() {
    if (!g_init) { init_luts(); init_mc_tables(); g_init = true; }
    return new Oracle();
}
void orc_destroy(void* h) { delete (Oracle*)h; }
// stages: 1 = octree only, 2 = +splat/divergence, 3 = +solve/iso, 4 = everything, 40 = everything except the refinement passes
int orc_run(void* h, const float* xyz, const float* nrm, int n, int depth, int stages) { return ((Oracle*)h)->run(xyz, nrm, n, depth, stages); }
// Teacher forcing for stage-by-stage pinning: overwrite an intermediate with the reference's
// own values, then re-run a single stage with orc_stage().
int orc_set(void* h, const char* name, const void* src, long long bytes) {
    Oracle& o = *(Oracle*)h;
    std::string s(name);
    std::vector<float>* v = nullptr;
    if (s == "vectorfield") v = &o.V; else if (s == "divergence") v = &o.divg; else if (s == "x") v = &o.x;
    if (s == "iso") { if (bytes != 4) return -2; std::memcpy(&o.iso, src, 4); return 0; }
    if (!v) return -1;
    if ((long long)(v->size() * 4) != bytes) return -2;
    std::memcpy(v->data(), src, (size_t)bytes);
    return 0;
}
int orc_stage(void* h, const char* name) {
    Oracle& o = *(Oracle*)h;
    std::string s(name);
    if (s == "splat") o.splat();
    else if (s == "divergence") o.divergence();
    else if (s == "solve") o.solve();
    else if (s == "iso") o.iso_value();
    else if (s == "mc_main") {
        o.build_vertices(); o.build_edges(); o.build_faces(); o.vertex_values();
        o.meshV.clear(); o.meshT.clear(); o.passes.clear();
        o.mc_main_pass(); o.find_subdivide();
    } else if (s == "mc") {
        o.build_vertices(); o.build_edges(); o.build_faces(); o.vertex_values();
        o.meshV.clear(); o.meshT.clear(); o.passes.clear();
        o.mc_main_pass(); o.find_subdivide(); o.refine();
    } else return -1;
    return 0;
}
// Copies a named array into dst (if dst != NULL and cap is large enough); returns its size in bytes
// or -1 for an unknown name.
long long orc_get(void* h, const char* name, void* dst, long long cap) {
    Oracle& o = *(Oracle*)h;
    std::string s(name);
    const void* src = nullptr;
    long long bytes = -1;
    std::vector<int> tmpi;
    std::vector<double> tmpd;
    std::vector<float> tmpf;
    auto V = [&](const auto& v) { src = v.data(); bytes = (long long)(v.size() * sizeof(v[0])); };
    if (s == "points") V(o.P); else if (s == "normals") V(o.Nr); else if (s == "sorted_idx") V(o.sortedIdx);
    else if (s == "sorted_key") V(o.sortedKey); else if (s == "base") V(o.base); else if (s == "count") V(o.cnt);
    else if (s == "key") V(o.key); else if (s == "pidx") V(o.pidx); else if (s == "pnum") V(o.pnum);
    else if (s == "parent") V(o.parent); else if (s == "didx") V(o.didx); else if (s == "dnum") V(o.dnum);
    else if (s == "children") V(o.children); else if (s == "neighs") V(o.neighs); else if (s == "p2n") V(o.p2n);
    else if (s == "vectorfield") V(o.V); else if (s == "divergence") V(o.divg); else if (s == "x") V(o.x);
    else if (s == "pointvalue") V(o.pointValue); else if (s == "cg_iters") V(o.cgIters);
    else if (s == "vvalue") V(o.vvalue); else if (s == "vertex_owner") V(o.vOwner); else if (s == "vertex_kind") V(o.vKind);
    else if (s == "vertex_pos") V(o.vPos); else if (s == "edge_owner") V(o.eOwner); else if (s == "edge_kind") V(o.eKind);
    else if (s == "face_owner") V(o.fOwner); else if (s == "subdivide") V(o.subdivide);
    else if (s == "mesh_v") V(o.meshV); else if (s == "mesh_t") V(o.meshT);
    else if (s == "iso") { tmpf = {o.iso}; V(tmpf); }
    else if (s == "center_scale") { tmpf = {o.center[0], o.center[1], o.center[2], o.scale}; V(tmpf); }
    else if (s == "passes") { for (auto& p : o.passes) { tmpi.push_back(p.kind == "main" ? 0 : (p.kind == "coarse" ? 1 : 2)); tmpi.push_back(p.nv); tmpi.push_back(p.nt); } V(tmpi); }
    else if (s == "lap_stencil") {   // 4 stencil values per depth (0..3 off-centre axes), from a centre node
        for (int d = 0; d <= o.D; d++) {
            int nd = 1 << d, c = nd / 2;
            for (int t = 0; t < 4; t++) {
                double ff[3], d2[3];
                for (int a = 0; a < 3; a++) { int dl = (a < t && nd > 1) ? ((c + 1 < nd) ? 1 : -1) : 0; ff[a] = o.ffSame[d][c * 3 + dl + 1]; d2[a] = o.d2Same[d][c * 3 + dl + 1]; }
                tmpf.push_back((float)(d2[0] * ff[1] * ff[2] + d2[1] * ff[0] * ff[2] + d2[2] * ff[0] * ff[1]));
            }
        }
        V(tmpf);
    }
    if (bytes < 0) return -1;
    if (dst && cap >= bytes && bytes > 0) std::memcpy(dst, src, (size_t)bytes);
    return bytes;
}
}

#ifdef ORC_CLI
// orc_cli in.bnpts depth [stages] : runs the oracle on a raw float32 x6 file, prints counts + timings
#include <chrono>
int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.bnpts depth [stages]\n", argv[0]); return 2; }
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 3;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    int n = (int)(sz / 24);
    std::vector<float> raw(6 * (size_t)n), xyz(3 * (size_t)n), nr(3 * (size_t)n);
    if (fread(raw.data(), 24, n, fp) != (size_t)n) return 4;
    fclose(fp);
    for (int i = 0; i < n; i++) for (int a = 0; a < 3; a++) { xyz[3 * i + a] = raw[6 * i + a]; nr[3 * i + a] = raw[6 * i + 3 + a]; }
    Oracle* o = (Oracle*)orc_create();
    auto t0 = std::chrono::steady_clock::now();
    o->run(xyz.data(), nr.data(), n, atoi(argv[2]), argc > 3 ? atoi(argv[3]) : 4);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("N=%d D=%d M=%d iso=%g verts=%zu tris=%zu seconds=%.3f\n", n, o->D, o->M, o->iso, o->meshV.size() / 3, o->meshT.size() / 3, sec);
    for (int d = 0; d <= o->D; d++) printf("depth %d nodes %d cg_iters %d\n", d, o->cnt[d], o->cgIters.empty() ? 0 : o->cgIters[d]);
    for (auto& p : o->passes) printf("pass %s nv %d nt %d\n", p.kind.c_str(), p.nv, p.nt);
    return 0;
}
#endif
