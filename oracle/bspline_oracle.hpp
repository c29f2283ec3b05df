// TEST INFRASTRUCTURE ONLY (oracle/). Never included or linked by the product.
//
// CPU restatement of the reference's host-side B-spline precompute:
//   - float polynomial arithmetic         Polynomial.inl:66-260 (eval, integral, *, scale, shift)
//   - piecewise polynomials               PPolynomial.inl:30-75, 99-110, 152-176, 246-305, 367-412
//   - per-index base functions + tables   FunctionData.inl:112-215 (set / setDotTables), 265-300
//   - index -> centre/width               BinaryNode.cuh:39-66
// The reference does all polynomial arithmetic in `float` and stores table
// entries as `double`; the rounding noise (e.g. <F,F'>(0) = -1.335e-05, not 0)
// is part of the values it feeds to its kernels, so this restatement keeps the
// same operation order.  Compile with -ffp-contract=off and no -march flags
// (the reference host code is built for baseline x86-64: no FMA contraction).
//
// Pinned against the reference itself: oracle/_ref/ref_tables (the reference's
// own headers compiled by g++, see ref_tables_harness.cpp) -> tests/golden/
// tables_d5.bin, tables_d8.bin; tests/test_oracle_tables.py checks bit-equality.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

constexpr int kMaxDeg = 4;

struct Poly {              // Polynomial<Degree>, Degree <= 4
    float c[kMaxDeg + 1];
    Poly() { std::memset(c, 0, sizeof(c)); }
};
struct SPoly {             // StartingPolynomial
    Poly p;
    float start = 0;
};
struct PPoly {             // PPolynomial<deg>
    int deg = 0;
    std::vector<SPoly> s;
};

// Polynomial.inl:73-81
inline float poly_eval(const Poly& p, int deg, float t) {
    float temp = 1, v = 0;
    for (int i = 0; i <= deg; i++) { v += temp * p.c[i]; temp *= t; }
    return v;
}
// Polynomial.inl:82-93
inline float poly_integral(const Poly& p, int deg, float tMin, float tMax) {
    float v = 0, t1 = tMin, t2 = tMax;
    for (int i = 0; i <= deg; i++) {
        v += p.c[i] * (t2 - t1) / (i + 1);
        t1 *= tMin;
        t2 *= tMax;
    }
    return v;
}
// Polynomial.inl:66-72  (indefinite integral, deg -> deg+1)
inline Poly poly_antiderivative(const Poly& p, int deg) {
    Poly q;
    q.c[0] = 0;
    for (int i = 0; i <= deg; i++) q.c[i + 1] = p.c[i] / (i + 1);
    return q;
}
// Polynomial.inl:59-64
inline Poly poly_derivative(const Poly& p, int deg) {
    Poly q;
    for (int i = 0; i < deg; i++) q.c[i] = p.c[i + 1] * (i + 1);
    return q;
}
// Polynomial.inl:169-175
inline Poly poly_mul(const Poly& a, int da, const Poly& b, int db) {
    Poly q;
    for (int i = 0; i <= da; i++)
        for (int j = 0; j <= db; j++) q.c[i + j] += a.c[i] * b.c[j];
    return q;
}
// Polynomial.inl:225-234
inline Poly poly_scale(const Poly& p, int deg, float s) {
    Poly q = p;
    float s2 = 1.0;
    for (int i = 0; i <= deg; i++) { q.c[i] *= s2; s2 /= s; }
    return q;
}
// Polynomial.inl:235-247
inline Poly poly_shift(const Poly& p, int deg, float t) {
    Poly q;
    for (int i = 0; i <= deg; i++) {
        float temp = 1;
        for (int j = i; j >= 0; j--) {
            q.c[j] += p.c[i] * temp;
            temp *= -t * j;
            temp /= (i - j + 1);
        }
    }
    return q;
}

// PPolynomial.inl:99-110  set(sps,count): stable sort by start, merge equal starts
inline PPoly pp_from_pieces(std::vector<SPoly> sps, int deg) {
    std::stable_sort(sps.begin(), sps.end(), [](const SPoly& a, const SPoly& b) { return a.start < b.start; });
    PPoly q;
    q.deg = deg;
    for (auto& sp : sps) {
        if (q.s.empty() || sp.start != q.s.back().start) q.s.push_back(sp);
        else for (int k = 0; k <= deg; k++) q.s.back().p.c[k] += sp.p.c[k];
    }
    return q;
}
// PPolynomial.inl:152-157
inline float pp_eval(const PPoly& f, float t) {
    float v = 0;
    for (size_t i = 0; i < f.s.size() && t > f.s[i].start; i++) v += poly_eval(f.s[i].p, f.deg, t);
    return v;
}
// PPolynomial.inl:159-176
inline float pp_integral(const PPoly& f, float tMin, float tMax) {
    int m = 1;
    float start = tMin, end = tMax, s, v = 0;
    if (tMin > tMax) { m = -1; start = tMax; end = tMin; }
    for (size_t i = 0; i < f.s.size() && f.s[i].start < end; i++) {
        if (start < f.s[i].start) s = f.s[i].start; else s = start;
        v += poly_integral(f.s[i].p, f.deg, s, end);
    }
    return v * m;
}
// PPolynomial.inl:246-262 (operator*): all pairwise products, start = max
inline PPoly pp_mul(const PPoly& a, const PPoly& b) {
    std::vector<SPoly> sp(a.s.size() * b.s.size());
    for (size_t i = 0; i < a.s.size(); i++)
        for (size_t j = 0; j < b.s.size(); j++) {
            SPoly& o = sp[i * b.s.size() + j];
            o.start = a.s[i].start > b.s[j].start ? a.s[i].start : b.s[j].start;   // PPolynomial.inl:35-40
            o.p = poly_mul(a.s[i].p, a.deg, b.s[j].p, b.deg);
        }
    return pp_from_pieces(sp, a.deg + b.deg);
}
inline PPoly pp_scale(const PPoly& f, float s) {       // PPolynomial.inl:274-280, 48-54
    PPoly q = f;
    for (auto& sp : q.s) { sp.start = sp.start * s; sp.p = poly_scale(sp.p, f.deg, s); }
    return q;
}
inline PPoly pp_shift(const PPoly& f, float t) {       // PPolynomial.inl:281-287, 55-61
    PPoly q = f;
    for (auto& sp : q.s) { sp.start = sp.start + t; sp.p = poly_shift(sp.p, f.deg, t); }
    return q;
}
inline PPoly pp_derivative(const PPoly& f) {           // PPolynomial.inl:288-297
    PPoly q = f;
    q.deg = f.deg - 1;
    for (auto& sp : q.s) sp.p = poly_derivative(sp.p, f.deg);
    return q;
}
inline PPoly pp_div(const PPoly& f, float s) {         // PPolynomial.inl:322-325, Polynomial.inl:194-197
    PPoly q = f;
    for (auto& sp : q.s) for (int k = 0; k <= f.deg; k++) sp.p.c[k] /= s;
    return q;
}
// PPolynomial.inl:389-412 MovingAverage
inline PPoly pp_moving_average(const PPoly& f, float radius) {
    std::vector<SPoly> sps(f.s.size() * 2);
    for (size_t i = 0; i < f.s.size(); i++) {
        sps[2 * i].start = f.s[i].start - radius;
        sps[2 * i + 1].start = f.s[i].start + radius;
        Poly anti = poly_antiderivative(f.s[i].p, f.deg);
        Poly p = anti;
        p.c[0] -= poly_eval(anti, f.deg + 1, f.s[i].start);
        sps[2 * i].p = poly_shift(p, f.deg + 1, -radius);
        Poly neg = poly_shift(p, f.deg + 1, radius);
        for (int k = 0; k <= f.deg + 1; k++) neg.c[k] = neg.c[k] * -1;
        sps[2 * i + 1].p = neg;
    }
    PPoly A = pp_from_pieces(sps, f.deg + 1);
    // "A*1.0/(2*radius)": two coefficient-wise float ops
    for (auto& sp : A.s) for (int k = 0; k <= A.deg; k++) sp.p.c[k] *= 1.0f;
    float d = 2 * radius;
    for (auto& sp : A.s) for (int k = 0; k <= A.deg; k++) sp.p.c[k] /= d;
    return A;
}
// PPolynomial.inl:367-388 ConstantFunction / GaussianApproximation (default width 0.5 at every level)
inline PPoly pp_gaussian_approximation(int deg) {
    PPoly q;
    q.deg = 0;
    q.s.resize(2);
    q.s[0].start = -0.5f; q.s[0].p.c[0] = 1.0f;
    q.s[1].start = 0.5f;  q.s[1].p.c[0] = -1.0f;
    for (int d = 0; d < deg; d++) q = pp_moving_average(q, 0.5f);
    return q;
}

// BinaryNode.cuh:46-66
inline void center_and_width(int idx, double& center, double& width) {
    int i = idx + 1, depth = -1;
    while (i) { i >>= 1; depth++; }
    int offset = (idx + 1) - (1 << depth);
    width = 1.0 / (1 << depth);
    center = (0.5 + offset) * width;
}

struct BSplineData {
    int depth = 0, res = 0;
    PPoly F;                       // GaussianApproximation()/F(0)            main.cu:3308,3322
    PPoly baseFunction, dBaseFunction;   // FunctionData.inl:134-137
    std::vector<PPoly> baseFunctions;    // deg 3 (top coefficient 0), FunctionData.inl:139-146
    PPoly maxDepthFunction;        // F.scale(2^-D)                           main.cu:3355

    void set(int maxDepth) {
        depth = maxDepth;
        res = (1 << (depth + 1)) - 1;
        PPoly g = pp_gaussian_approximation(2);
        float f0 = pp_eval(g, 0.0f);
        baseFunction = pp_div(g, f0);
        dBaseFunction = pp_derivative(baseFunction);
        F = pp_div(g, f0);
        baseFunctions.resize(res);
        for (int i = 0; i < res; i++) {
            double c1, w1;
            center_and_width(i, c1, w1);
            PPoly b = pp_shift(pp_scale(baseFunction, (float)w1), (float)c1);
            b.deg = 3;
            baseFunctions[i] = b;
        }
        maxDepthFunction = pp_scale(F, (float)(1.0 / (1 << depth)));
    }
    // FunctionData.inl:265-300 (normalize = 0)
    double dotProduct(double c1, double w1, double c2, double w2) const {
        double r = std::fabs(baseFunction.s[0].start);
        PPoly o = pp_shift(pp_scale(baseFunction, (float)(w2 / w1)), (float)((c2 - c1) / w1));
        return pp_integral(pp_mul(baseFunction, o), (float)(-2 * r), (float)(2 * r)) * w1;
    }
    double dDotProduct(double c1, double w1, double c2, double w2) const {
        double r = std::fabs(baseFunction.s[0].start);
        PPoly o = pp_shift(pp_scale(baseFunction, (float)(w2 / w1)), (float)((c2 - c1) / w1));
        return pp_integral(pp_mul(dBaseFunction, o), (float)(-2 * r), (float)(2 * r));
    }
    double d2DotProduct(double c1, double w1, double c2, double w2) const {
        double r = std::fabs(baseFunction.s[0].start);
        PPoly o = pp_shift(pp_scale(dBaseFunction, (float)(w2 / w1)), (float)((c2 - c1) / w1));
        return pp_integral(pp_mul(dBaseFunction, o), (float)(-2 * r), (float)(2 * r)) / w2;
    }
    // value the reference's tables hold at flat index a + res*b (FunctionData.inl:159-215)
    // which: 0 = dotTable <F,F>, 1 = dDotTable, 2 = d2DotTable
    double table(int which, int a, int b) const {
        int i = a > b ? a : b, j = a > b ? b : a;
        double c1, w1, c2, w2;
        center_and_width(i, c1, w1);
        center_and_width(j, c2, w2);
        double t1 = baseFunction.s[0].start, t2 = baseFunction.s.back().start;
        double start1 = t1 * w1 + c1, end1 = t2 * w1 + c1;
        double start = t1 * w2 + c2, end = t2 * w2 + c2;
        if (start < start1) start = start1;
        if (end > end1) end = end1;
        if (start >= end) return 0;
        double dot = dotProduct(c1, w1, c2, w2);
        if (std::fabs(dot) < 1e-15) return 0;
        if (which == 0) return dot;
        if (which == 2) return d2DotProduct(c1, w1, c2, w2);
        double d = dDotProduct(c1, w1, c2, w2);
        // idx1 = i+res*j gets d, idx2 = j+res*i gets -d (written second: wins when i==j)
        return (a > b) ? d : -d;
    }
};

// ConfirmedPPolynomial.cuh:79-91 value(): cumulative pieces, strict '>' on starts.
// The reference evaluates this on the GPU, where nvcc contracts v += temp*c into an FMA.
inline float confirmed_value(const PPoly& f, float val) {
    float res = 0;
    for (size_t i = 0; i < f.s.size() && val > f.s[i].start; i++) {
        float temp = 1, v = 0;
        for (int j = 0; j <= f.deg; j++) { v = std::fmaf(temp, f.s[i].p.c[j], v); temp *= val; }
        res += v;
    }
    return res;
}
// ConfirmedPPolynomial.cuh:35-48 shift() as executed on the GPU (FMA-contracted accumulate)
inline PPoly confirmed_shift(const PPoly& f, float t) {
    PPoly r = f;
    for (size_t i = 0; i < f.s.size(); i++) {
        r.s[i].start = f.s[i].start + t;
        for (int k = 0; k <= f.deg; k++) r.s[i].p.c[k] = 0;
        for (int k = 0; k <= f.deg; k++) {
            float temp = 1;
            for (int j = k; j >= 0; j--) {
                r.s[i].p.c[j] = std::fmaf(f.s[i].p.c[k], temp, r.s[i].p.c[j]);
                temp *= -t * j;
                temp /= (k - j + 1);
            }
        }
    }
    return r;
}

}  // namespace orc
