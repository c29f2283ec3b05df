// Host-side B-spline precompute of the product: the quadratic B-spline basis, its per-index
// shifted/scaled copies and the 1-D integral tables, in the compact translation-invariant form
// the kernels consume.
//
// What it replaces in the reference: PPolynomial<2>::GaussianApproximation (PPolynomial.inl:
// 385-412), FunctionData<2,double>::set / setDotTables (FunctionData.inl:112-215) and the three
// res x res double tables uploaded at main.cu:3325-3336 (33 MB each at depth 10, 537 MB at 12).
// The reference does the polynomial algebra in float and its rounding noise is part of the
// numbers its kernels see (e.g. <F,F'>(0) = -1.335e-05), so the same float operation order is
// kept here; tests/test_tables.py checks the results bit-for-bit against
// tests/golden/tables_d*.bin, which were produced by the reference's own host code.
//
// Because every table entry depends only on the width ratio and the centre distance of the two
// functions (both dyadic, so exact), the res^2 tables collapse to
//   dfT[d][t]        float, t in [0,3k), k = 2^(D-d):  <dF_o , F_s> for a depth-d node o and a
//                    depth-D slot s with  t = off_s - k*(off_o - 1)           (divergence)
//   stencil[d][27]   float: the 27-point same-depth Laplacian row              (CG)
// plus baseFn[res][4][5] (piece coefficients c0..c3 + start) for point evaluation.
#pragma once
#include <vector>

namespace prb {

struct BSplineTables {
    int depth = 0, res = 0;
    float gauss[4][4];                 // B-hat pieces (c0,c1,c2,start), centred at 0, width 1
    float maxDepthFn[4][4];            // B-hat scaled to width 2^-D (main.cu:3355)
    std::vector<float> baseFn;         // res * 4 * 5
    std::vector<float> dfT;            // concatenated per depth
    std::vector<int> dfOffset;         // start of depth d inside dfT (size D+2)
    std::vector<float> stencil;        // (D+1) * 27
    // raw same-depth 1-D values per depth, (delta = 0, 1): for tests
    std::vector<double> ff0, ff1, d20, d21;
    // cross-depth 1-D integrals of the OPT-IN cascadic solver mode (not in the reference, which solves the depths independently):
    // for a node o of depth d and a node n of a coarser depth e < d, k = 2^(d-e), u = off_o - k * (off_n - 1) in [0, 3k):
    //   ffX[crossOffset(d, e) + u] = <F_o, F_n>,   d2X[...] = <F_o', F_n'>          (same integration code as the tables above)
    std::vector<double> ffX, d2X;
    std::vector<int> crossOff;         // (D+1) x (D+1), row d, column e (valid for e < d)
    int cross_offset(int d, int e) const { return crossOff[(size_t)d * (depth + 1) + e]; }
};

void build_bspline_tables(int depth, BSplineTables& out);

}  // namespace prb
