// TEST INFRASTRUCTURE ONLY (oracle/). Never linked into the product.
//
// Harness around the REFERENCE's own host-side B-spline code
// (/root/reference/{Polynomial,PPolynomial,FunctionData,BinaryNode}.*, included
// where they lie — nothing is copied into this repo).  It reproduces the host
// precompute of main.cu:3308-3359 and dumps what the reference uploads to the
// GPU, so that oracle/ and the product's host tables can be pinned bit-for-bit.
//
// Built by oracle/Makefile into oracle/_ref/ref_tables (g++, CPU only).
// Usage: ref_tables <depth> <out.bin>
//   layout (little endian):
//     int32 depth, int32 res
//     float  gauss[4][4]        GaussianApproximation()/F(0): (c0,c1,c2,start) x4
//     float  maxdepth[4][4]     BaseFunctionMaxDepth = F.scale(2^-D): (c0,c1,c2,start) x4
//     float  base[res][4][5]    baseFunctions[i]: (c0..c3,start) x4   (80 B each)
//     double FF[res*res], DF[res*res], D2[res*res]   (only if depth <= 6, else
//            the compact probes below are the fixture)
//     -- probes (always): for every depth d<=D, same-depth FF/D2 at offset 0,1
//        and the full cross-depth DF row of node offset 2^(d-1) (centre node):
//        int32 nprobe; then nprobe x {int32 a, int32 b, double FF, double DF, double D2}
#include <cuda_runtime_api.h>
#undef __host__
#undef __device__
#define __host__
#define __device__
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "PPolynomial.cuh"
#include "FunctionData.cuh"
#include "BinaryNode.cuh"
#include "ConfirmedPPolynomial.cuh"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s depth out.bin\n", argv[0]); return 2; }
    int D = atoi(argv[1]);
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) return 3;
    PPolynomial<2> F = PPolynomial<2>::GaussianApproximation();
    FunctionData<2, double> fData;
    fData.set(D, F, 0, 0);
    fData.setDotTables(fData.DOT_FLAG | fData.D_DOT_FLAG | fData.D2_DOT_FLAG);
    F = F / F(0);
    int res = fData.res;
    fwrite(&D, 4, 1, fp);
    fwrite(&res, 4, 1, fp);
    for (int i = 0; i < 4; i++) {
        float v[4] = {F.polys[i].p.coefficients[0], F.polys[i].p.coefficients[1], F.polys[i].p.coefficients[2], F.polys[i].start};
        fwrite(v, 4, 4, fp);
    }
    ConfirmedPPolynomial<2, 4> bmax(F.scale(1.0 / (1 << D)));
    for (int i = 0; i < 4; i++) {
        float v[4] = {bmax.polys[i].p.coefficients[0], bmax.polys[i].p.coefficients[1], bmax.polys[i].p.coefficients[2], bmax.polys[i].start};
        fwrite(v, 4, 4, fp);
    }
    for (int f = 0; f < res; f++) {
        ConfirmedPPolynomial<3, 4> b;
        b = fData.baseFunctions[f];
        for (int i = 0; i < 4; i++) {
            float v[5] = {b.polys[i].p.coefficients[0], b.polys[i].p.coefficients[1], b.polys[i].p.coefficients[2], b.polys[i].p.coefficients[3], b.polys[i].start};
            fwrite(v, 4, 5, fp);
        }
    }
    if (D <= 6) {
        fwrite(fData.dotTable, 8, (size_t)res * res, fp);
        fwrite(fData.dDotTable, 8, (size_t)res * res, fp);
        fwrite(fData.d2DotTable, 8, (size_t)res * res, fp);
    }
    struct Probe { int a, b; double ff, df, d2; };
    std::vector<Probe> probes;
    for (int d = 0; d <= D; d++) {
        int n = 1 << d;
        int o = n / 2;
        int a = (n - 1) + o;
        // same depth: offsets -1, 0, +1 (if inside)
        for (int dd = -1; dd <= 1; dd++) {
            int ob = o + dd;
            if (ob < 0 || ob >= n) continue;
            int b = (n - 1) + ob;
            Probe p = {a, b, fData.dotTable[a + res * b], fData.dDotTable[a + res * b], fData.d2DotTable[a + res * b]};
            probes.push_back(p);
            Probe q = {b, a, fData.dotTable[b + res * a], fData.dDotTable[b + res * a], fData.d2DotTable[b + res * a]};
            probes.push_back(q);
        }
        // cross depth: node (d,o) against every depth-D function
        int nD = 1 << D;
        for (int s = 0; s < nD; s++) {
            int b = (nD - 1) + s;
            double df = fData.dDotTable[a + res * b];
            if (df != 0) { Probe p = {a, b, fData.dotTable[a + res * b], df, fData.d2DotTable[a + res * b]}; probes.push_back(p); }
        }
    }
    int np = (int)probes.size();
    fwrite(&np, 4, 1, fp);
    for (auto& p : probes) { fwrite(&p.a, 4, 1, fp); fwrite(&p.b, 4, 1, fp); fwrite(&p.ff, 8, 1, fp); fwrite(&p.df, 8, 1, fp); fwrite(&p.d2, 8, 1, fp); }
    fclose(fp);
    printf("depth %d res %d probes %d\n", D, res, np);
    return 0;
}
