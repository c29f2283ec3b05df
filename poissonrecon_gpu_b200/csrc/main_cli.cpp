// poisson_recon -- command-line drop-in for the reference's main() (main.cu:3247-4573).
//
//   poisson_recon --in points.{ply,bnpts,txt} --out mesh.ply --depth D [--binary] [--device k]
//                 [--no-refine] [--json]
//
// The reference hard-codes its paths (main.cu:3251-3252) and compiles the depth in
// (main.cu:69); the `--name value` convention is the one its own (unused) parser implements
// (CmdLineParser.cu:205-242).  Stage timings go to stdout like the reference's printf trail
// (Debug.cuh:53 cpuSecond deltas); `--json` adds one machine-readable line.  All compute goes
// through the C ABI (include/prb.h); there is no CPU path -- without a B200 the tool exits
// with the library's error.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "prb.h"
#include "prb_io.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void usage(const char* a0) {
    std::fprintf(stderr, "usage: %s --in <points.ply|.bnpts|ascii> --out <mesh.ply> [--depth D=8] [--binary] [--device k] [--no-refine] [--json]\n", a0);
}

int main(int argc, char** argv) {
    std::string in, out;
    int depth = 8, device = 0, binary = 0, refine = 1, json = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--in") in = val("--in");
        else if (a == "--out") out = val("--out");
        else if (a == "--depth") depth = std::atoi(val("--depth"));
        else if (a == "--device") device = std::atoi(val("--device"));
        else if (a == "--binary") binary = 1;
        else if (a == "--no-refine") refine = 0;
        else if (a == "--json") json = 1;
        else if (a == "--help" || a == "-h") { usage(argv[0]); return 0; }
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); usage(argv[0]); return 2; }
    }
    if (in.empty() || out.empty()) { usage(argv[0]); return 2; }
    const double t0 = now_s();
    float *xyz = nullptr, *nrm = nullptr;
    int64_t n = 0;
    if (prbio_read_points(in.c_str(), &xyz, &nrm, &n) != 0) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const double t1 = now_s();
    std::printf("Total points number:%lld ,Read takes:%lfs\n", (long long)n, t1 - t0);
    if (n <= 0) { std::fprintf(stderr, "no points in %s\n", in.c_str()); return 1; }
    prb_context* ctx = nullptr;
    if (prb_create(device, depth, &ctx) != PRB_OK) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    prb_set_option(ctx, "refine", refine);
    const double t2 = now_s();
    if (prb_set_points(ctx, xyz, nrm, n) != PRB_OK || prb_run(ctx) != PRB_OK) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const float* v = nullptr;
    const int32_t* t = nullptr;
    int64_t nv = 0, nt = 0;
    if (prb_get_mesh(ctx, &v, &nv, &t, &nt) != PRB_OK) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const double t3 = now_s();
    prb_stats s;
    prb_get_stats(ctx, &s);
    std::printf("NodeArray_sz:%d\n", s.n_nodes);
    std::printf("GPU NodeArray build takes:%lfs\n", (s.ms_h2d + s.ms_octree) * 1e-3);
    std::printf("Compute Vector Field takes:%lfs\n", s.ms_splat * 1e-3);
    std::printf("Compute nodes' divergence takes:%lfs\n", s.ms_divergence * 1e-3);
    std::printf("GPU Laplacian Iteration takes:%lfs\n", s.ms_solve * 1e-3);
    std::printf("isoValue:%f\nGPU calculate isoValue takes:%lfs\n", s.iso_value, s.ms_iso * 1e-3);
    std::printf("SubdivideNum:%d\n", s.n_subdivide);
    std::printf("GPU marching cubes + subdivide passes takes:%lfs (%d passes)\n", s.ms_extract * 1e-3, s.n_passes);
    std::printf("Vertices:%lld Triangles:%lld\n", (long long)nv, (long long)nt);
    if (prbio_write_mesh(out.c_str(), v, nv, t, nt, s.center, s.scale, binary) != 0) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const double t4 = now_s();
    std::printf("Output ply files takes %lfs\n", t4 - t3);
    std::printf("The whole project takes %lfs (including I/O)\n", t4 - t0);
    if (json) {
        std::printf("{\"n_points\": %lld, \"depth\": %d, \"n_nodes\": %d, \"n_vertices\": %lld, \"n_triangles\": %lld, \"iso\": %.9g, "
                    "\"read_s\": %.6f, \"create_s\": %.6f, \"compute_s\": %.6f, \"device_ms\": %.3f, \"write_s\": %.6f, \"total_s\": %.6f}\n",
                    (long long)n, depth, s.n_nodes, (long long)nv, (long long)nt, (double)s.iso_value, t1 - t0, t2 - t1, t3 - t2, (double)s.ms_total, t4 - t3, t4 - t0);
    }
    prb_destroy(ctx);
    prbio_free(xyz);
    prbio_free(nrm);
    return 0;
}
