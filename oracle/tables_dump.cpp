// TEST INFRASTRUCTURE ONLY (oracle/).  Dumps the oracle's restated B-spline tables in the
// byte layout of oracle/_ref/ref_tables (see ref_tables_harness.cpp) so that the two can be
// compared with cmp / numpy.  Usage: tables_dump <depth> <out.bin>
#include <cstdio>
#include <cstdlib>
#include "bspline_oracle.hpp"
using namespace orc;
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    int D = atoi(argv[1]);
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) return 3;
    BSplineData bs;
    bs.set(D);
    int res = bs.res;
    fwrite(&D, 4, 1, fp);
    fwrite(&res, 4, 1, fp);
    for (int i = 0; i < 4; i++) { float v[4] = {bs.F.s[i].p.c[0], bs.F.s[i].p.c[1], bs.F.s[i].p.c[2], bs.F.s[i].start}; fwrite(v, 4, 4, fp); }
    for (int i = 0; i < 4; i++) { const SPoly& s = bs.maxDepthFunction.s[i]; float v[4] = {s.p.c[0], s.p.c[1], s.p.c[2], s.start}; fwrite(v, 4, 4, fp); }
    for (int f = 0; f < res; f++)
        for (int i = 0; i < 4; i++) { const SPoly& s = bs.baseFunctions[f].s[i]; float v[5] = {s.p.c[0], s.p.c[1], s.p.c[2], s.p.c[3], s.start}; fwrite(v, 4, 5, fp); }
    if (D <= 6)
        for (int which = 0; which < 3; which++)
            for (int b = 0; b < res; b++)
                for (int a = 0; a < res; a++) { double v = bs.table(which, a, b); fwrite(&v, 8, 1, fp); }
    struct Probe { int a, b; double ff, df, d2; };
    std::vector<Probe> probes;
    for (int d = 0; d <= D; d++) {
        int n = 1 << d, o = n / 2, a = (n - 1) + o;
        for (int dd = -1; dd <= 1; dd++) {
            int ob = o + dd;
            if (ob < 0 || ob >= n) continue;
            int b = (n - 1) + ob;
            probes.push_back({a, b, bs.table(0, a, b), bs.table(1, a, b), bs.table(2, a, b)});
            probes.push_back({b, a, bs.table(0, b, a), bs.table(1, b, a), bs.table(2, b, a)});
        }
        int nD = 1 << D;
        for (int s = 0; s < nD; s++) {
            int b = (nD - 1) + s;
            double df = bs.table(1, a, b);
            if (df != 0) probes.push_back({a, b, bs.table(0, a, b), df, bs.table(2, a, b)});
        }
    }
    int np = (int)probes.size();
    fwrite(&np, 4, 1, fp);
    for (auto& p : probes) { fwrite(&p.a, 4, 1, fp); fwrite(&p.b, 4, 1, fp); fwrite(&p.ff, 8, 1, fp); fwrite(&p.df, 8, 1, fp); fwrite(&p.d2, 8, 1, fp); }
    fclose(fp);
    printf("depth %d res %d probes %d\n", D, res, np);
    return 0;
}
