// See bspline_host.h.  Float piecewise-polynomial algebra with fixed-capacity storage.
// Compile WITHOUT FMA contraction (-ffp-contract=off, no -march): the reference's host code is
// plain x86-64 SSE arithmetic and the tables must match it bit for bit.
#include "bspline_host.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace prb {
namespace {

constexpr int kCoef = 5;   // degree <= 4
constexpr int kCap = 16;   // pieces: 4x4 products before merging

struct Piece {
    float c[kCoef];
    float start;
};
struct PW {                // piecewise polynomial: value(t) = sum of pieces with start < t
    int n = 0, deg = 0;
    Piece p[kCap];
};

void zero(Piece& q) { std::memset(&q, 0, sizeof(q)); }

// f(x) -> f(x - t): binomial re-expansion, same multiply/divide order as the reference
// (Polynomial.inl:235-247) so the float rounding is identical.
Piece translate(const Piece& in, int deg, float t) {
    Piece q;
    zero(q);
    q.start = in.start + t;
    for (int i = 0; i <= deg; i++) {
        float w = 1;
        for (int j = i; j >= 0; j--) {
            q.c[j] += in.c[i] * w;
            w *= -t * j;
            w /= (i - j + 1);
        }
    }
    return q;
}
// f(x) -> f(x / s)  (Polynomial.inl:225-234)
Piece dilate(const Piece& in, int deg, float s) {
    Piece q = in;
    q.start = in.start * s;
    float f = 1.0;
    for (int i = 0; i <= deg; i++) {
        q.c[i] *= f;
        f /= s;
    }
    return q;
}
float horner_free_eval(const float* c, int deg, float t) {   // power-sum, not Horner (Polynomial.inl:73-81)
    float pw = 1, v = 0;
    for (int i = 0; i <= deg; i++) {
        v += pw * c[i];
        pw *= t;
    }
    return v;
}
float definite_integral(const float* c, int deg, float lo, float hi) {   // Polynomial.inl:82-93
    float v = 0, a = lo, b = hi;
    for (int i = 0; i <= deg; i++) {
        v += c[i] * (b - a) / (i + 1);
        a *= lo;
        b *= hi;
    }
    return v;
}
// sort pieces by start (stable) and merge equal starts by adding coefficients
// (PPolynomial.inl:99-110; glibc qsort is a stable merge sort for these sizes)
void canonicalise(PW& f) {
    std::stable_sort(f.p, f.p + f.n, [](const Piece& a, const Piece& b) { return a.start < b.start; });
    int m = 0;
    for (int i = 0; i < f.n; i++) {
        if (m == 0 || f.p[i].start != f.p[m - 1].start) f.p[m++] = f.p[i];
        else for (int k = 0; k <= f.deg; k++) f.p[m - 1].c[k] += f.p[i].c[k];
    }
    f.n = m;
}
PW box_average(const PW& f, float radius) {    // PPolynomial.inl:389-412
    PW A;
    A.deg = f.deg + 1;
    A.n = 2 * f.n;
    for (int i = 0; i < f.n; i++) {
        Piece anti;
        zero(anti);
        for (int k = 0; k <= f.deg; k++) anti.c[k + 1] = f.p[i].c[k] / (k + 1);
        Piece q = anti;
        q.c[0] -= horner_free_eval(anti.c, A.deg, f.p[i].start);
        q.start = f.p[i].start;
        A.p[2 * i] = translate(q, A.deg, -radius);
        Piece r = translate(q, A.deg, radius);
        for (int k = 0; k <= A.deg; k++) r.c[k] = r.c[k] * -1;
        A.p[2 * i + 1] = r;
    }
    canonicalise(A);
    float d = 2 * radius;
    for (int i = 0; i < A.n; i++)
        for (int k = 0; k <= A.deg; k++) {
            A.p[i].c[k] *= 1.0f;
            A.p[i].c[k] /= d;
        }
    return A;
}
float pw_eval(const PW& f, float t) {          // PPolynomial.inl:152-157
    float v = 0;
    for (int i = 0; i < f.n && t > f.p[i].start; i++) v += horner_free_eval(f.p[i].c, f.deg, t);
    return v;
}
float pw_integral(const PW& f, float lo, float hi) {   // PPolynomial.inl:159-176 (lo < hi here)
    float v = 0;
    for (int i = 0; i < f.n && f.p[i].start < hi; i++) {
        float s = lo < f.p[i].start ? f.p[i].start : lo;
        v += definite_integral(f.p[i].c, f.deg, s, hi);
    }
    return v;
}
PW pw_product(const PW& a, const PW& b) {      // PPolynomial.inl:246-262, 30-41
    PW q;
    q.deg = a.deg + b.deg;
    q.n = a.n * b.n;
    for (int i = 0; i < a.n; i++)
        for (int j = 0; j < b.n; j++) {
            Piece& o = q.p[i * b.n + j];
            zero(o);
            o.start = a.p[i].start > b.p[j].start ? a.p[i].start : b.p[j].start;
            for (int u = 0; u <= a.deg; u++)
                for (int v = 0; v <= b.deg; v++) o.c[u + v] += a.p[i].c[u] * b.p[j].c[v];
        }
    canonicalise(q);
    return q;
}
PW pw_map(const PW& f, float s, float t) {     // scale(s).shift(t)
    PW q;
    q.deg = f.deg;
    q.n = f.n;
    for (int i = 0; i < f.n; i++) q.p[i] = translate(dilate(f.p[i], f.deg, s), f.deg, t);
    return q;
}
PW pw_derivative(const PW& f) {
    PW q;
    q.deg = f.deg - 1;
    q.n = f.n;
    for (int i = 0; i < f.n; i++) {
        zero(q.p[i]);
        q.p[i].start = f.p[i].start;
        for (int k = 0; k < f.deg; k++) q.p[i].c[k] = f.p[i].c[k + 1] * (k + 1);
    }
    return q;
}

struct Basis {
    PW B, dB;
    double r;     // |first start| = 1.5
    // integrals of FunctionData.inl:265-300 with normalize = 0; (ratio, shift) are what the
    // reference passes as float to scale()/shift(): w2/w1 and (c2-c1)/w1.
    double ff(double ratio, double shift, double w1) const {
        return pw_integral(pw_product(B, pw_map(B, (float)ratio, (float)shift)), (float)(-2 * r), (float)(2 * r)) * w1;
    }
    double df(double ratio, double shift) const {
        return pw_integral(pw_product(dB, pw_map(B, (float)ratio, (float)shift)), (float)(-2 * r), (float)(2 * r));
    }
    double d2(double ratio, double shift, double w2) const {
        return pw_integral(pw_product(dB, pw_map(dB, (float)ratio, (float)shift)), (float)(-2 * r), (float)(2 * r)) / w2;
    }
    // support test of setDotTables (FunctionData.inl:176-189) in function-1 units
    bool overlaps(double ratio, double shift) const {
        double t1 = B.p[0].start, t2 = B.p[B.n - 1].start;
        double lo = t1 * ratio + shift, hi = t2 * ratio + shift;
        if (lo < t1) lo = t1;
        if (hi > t2) hi = t2;
        return lo < hi;
    }
};

}  // namespace

void build_bspline_tables(int D, BSplineTables& T) {
    T.depth = D;
    T.res = (1 << (D + 1)) - 1;
    PW box;
    box.deg = 0;
    box.n = 2;
    zero(box.p[0]);
    zero(box.p[1]);
    box.p[0].start = -0.5f; box.p[0].c[0] = 1.0f;
    box.p[1].start = 0.5f;  box.p[1].c[0] = -1.0f;
    PW g = box_average(box_average(box, 0.5f), 0.5f);
    float g0 = pw_eval(g, 0.0f);
    Basis bs;
    bs.B = g;
    for (int i = 0; i < g.n; i++)
        for (int k = 0; k <= g.deg; k++) bs.B.p[i].c[k] /= g0;
    bs.dB = pw_derivative(bs.B);
    bs.r = std::fabs(bs.B.p[0].start);
    for (int i = 0; i < 4; i++) {
        T.gauss[i][0] = bs.B.p[i].c[0]; T.gauss[i][1] = bs.B.p[i].c[1]; T.gauss[i][2] = bs.B.p[i].c[2]; T.gauss[i][3] = bs.B.p[i].start;
        Piece m = dilate(bs.B.p[i], 2, (float)(1.0 / (1 << D)));
        T.maxDepthFn[i][0] = m.c[0]; T.maxDepthFn[i][1] = m.c[1]; T.maxDepthFn[i][2] = m.c[2]; T.maxDepthFn[i][3] = m.start;
    }
    // per-index base functions (FunctionData.inl:139-146; index -> centre/width BinaryNode.cuh:46-66)
    T.baseFn.assign((size_t)T.res * 20, 0.f);
    for (int d = 0; d <= D; d++) {
        double w = 1.0 / (1 << d);
        for (int o = 0; o < (1 << d); o++) {
            double c = (0.5 + o) * w;
            int idx = (1 << d) - 1 + o;
            for (int i = 0; i < 4; i++) {
                Piece q = translate(dilate(bs.B.p[i], 2, (float)w), 2, (float)c);
                float* dst = &T.baseFn[(size_t)idx * 20 + i * 5];
                dst[0] = q.c[0]; dst[1] = q.c[1]; dst[2] = q.c[2]; dst[3] = 0.f; dst[4] = q.start;
            }
        }
    }
    // divergence tables.  Entry = value the reference holds at dDotTable[fi(o) + res*fi(s)].
    T.dfOffset.assign(D + 2, 0);
    for (int d = 0; d <= D; d++) T.dfOffset[d + 1] = T.dfOffset[d] + 3 * (1 << (D - d));
    T.dfT.assign(T.dfOffset[D + 1], 0.f);
    double wD = 1.0 / (1 << D);
    for (int d = 0; d <= D; d++) {
        int k = 1 << (D - d);
        double wd = 1.0 / (1 << d);
        for (int t = 0; t < 3 * k; t++) {
            double v = 0;
            if (d == D) {
                // same depth: the entry with the larger index first is +<dF_i,F_j>, its mirror
                // (and the diagonal, written last) is the negative (FunctionData.inl:203-206)
                double shift = (t == 1) ? 0.0 : -1.0;
                if (bs.overlaps(1.0, shift) && std::fabs(bs.ff(1.0, shift, wD)) >= 1e-15) {
                    double dd = bs.df(1.0, shift);
                    v = (t == 0) ? dd : -dd;
                }
            } else {
                // function 1 = the depth-D slot (larger index), function 2 = the depth-d node
                double ratio = (double)k;
                double shift = 1.5 * k - t - 0.5;
                if (bs.overlaps(ratio, shift) && std::fabs(bs.ff(ratio, shift, wD)) >= 1e-15) v = -bs.df(ratio, shift);
            }
            T.dfT[T.dfOffset[d] + t] = (float)v;      // main.cu:1050-1053 narrows to float before use
        }
        (void)wd;
    }
    // same-depth Laplacian rows (main.cu:1143-1158, 1199-1207)
    T.stencil.assign((size_t)(D + 1) * 27, 0.f);
    T.ff0.resize(D + 1); T.ff1.resize(D + 1); T.d20.resize(D + 1); T.d21.resize(D + 1);
    for (int d = 0; d <= D; d++) {
        double w = 1.0 / (1 << d);
        double f[2], s[2];
        f[0] = bs.ff(1.0, 0.0, w);  s[0] = bs.d2(1.0, 0.0, w);
        f[1] = bs.ff(1.0, -1.0, w); s[1] = bs.d2(1.0, -1.0, w);
        T.ff0[d] = f[0]; T.ff1[d] = f[1]; T.d20[d] = s[0]; T.d21[d] = s[1];
        for (int j = 0; j < 27; j++) {
            int a[3] = {j / 9 != 1, (j / 3) % 3 != 1, j % 3 != 1};
            double e = s[a[0]] * f[a[1]] * f[a[2]] + s[a[1]] * f[a[0]] * f[a[2]] + s[a[2]] * f[a[0]] * f[a[1]];
            T.stencil[(size_t)d * 27 + j] = (float)e;
        }
    }
    // cross-depth integrals (cascadic mode): function 1 = the finer node o (depth d), function 2 = the coarser node n (depth e)
    T.crossOff.assign((size_t)(D + 1) * (D + 1), 0);
    int total = 0;
    for (int d = 0; d <= D; d++)
        for (int e = 0; e < d; e++) { T.crossOff[(size_t)d * (D + 1) + e] = total; total += 3 << (d - e); }
    T.ffX.assign((size_t)total, 0.0);
    T.d2X.assign((size_t)total, 0.0);
    for (int d = 0; d <= D; d++)
        for (int e = 0; e < d; e++) {
            const int k = 1 << (d - e);
            const double w1 = 1.0 / (1 << d), w2 = 1.0 / (1 << e);
            for (int u = 0; u < 3 * k; u++) {
                const double ratio = (double)k, shift = 1.5 * k - u - 0.5;
                double f = 0, s2 = 0;
                if (bs.overlaps(ratio, shift)) {
                    f = bs.ff(ratio, shift, w1);
                    if (std::fabs(f) < 1e-15) f = 0; else s2 = bs.d2(ratio, shift, w2);
                }
                T.ffX[(size_t)T.cross_offset(d, e) + u] = f;
                T.d2X[(size_t)T.cross_offset(d, e) + u] = s2;
            }
        }
}

}  // namespace prb
