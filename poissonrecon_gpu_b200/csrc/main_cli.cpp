// poisson_recon -- command-line drop-in for the reference's main() (main.cu:3247-4573).
//
//   poisson_recon --in points.{ply,bnpts,txt} --out mesh.ply --depth D [--binary] [--device k] [--gpus N]
//                 [--dump DIR] [--weld] [--no-refine] [--json] [--arena-gb G]
//                 [--cascadic] [--density-weighted-iso]      (opt-in modes OUTSIDE reference parity, include/prb.h options)
//
// The reference hard-codes its paths (main.cu:3251-3252) and compiles the depth in
// (main.cu:69); the `--name value` convention is the one its own (unused) parser implements
// (CmdLineParser.cu:205-242).  Stage timings go to stdout like the reference's printf trail
// (Debug.cuh:53 cpuSecond deltas); `--json` adds one machine-readable line.  All compute goes
// through the C ABI (include/prb.h); there is no CPU path -- without a B200 the tool exits
// with the library's error.
//
// --gpus N (2..8, one NVLink box): the process forks N-1 workers BEFORE touching CUDA, one rank per GPU
// (devices --device .. --device+N-1).  The ranks exchange the 64-byte CUDA-IPC handles of their arenas
// through files in a private temporary directory, every rank uploads its slice of the samples
// (prb_set_points_sharded), and rank 0 assembles the distributed mesh from the workers' pieces
// ("mesh_layout") -- no Python, no MPI, no NCCL involved.
// --dump DIR: the parity arrays of prb_get_array as raw little-endian files DIR/<name>.bin (rank 0's view).
// --weld: merge vertices with identical positions (the reference duplicates the vertices on the seams between
// passes, main.cu:3220-3245; off by default to keep count parity).
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "prb.h"
#include "prb_io.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void usage(const char* a0) {
    std::fprintf(stderr, "usage: %s --in <points.ply|.bnpts|ascii> --out <mesh.ply> [--depth D=8] [--binary] [--device k] [--gpus N] [--dump DIR] [--weld] [--no-refine] [--json] [--arena-gb G] [--cascadic] [--density-weighted-iso]\n", a0);
}

struct Args {
    std::string in, out, dump;
    int depth = 8, device = 0, binary = 0, refine = 1, json = 0, gpus = 1, weld = 0, cascadic = 0, densityIso = 0;
    double arenaGb = 0;
};

static bool write_file(const std::string& path, const void* p, size_t bytes) {
    const std::string tmp = path + ".tmp";
    FILE* fp = std::fopen(tmp.c_str(), "wb");
    if (!fp) return false;
    bool ok = bytes == 0 || std::fwrite(p, 1, bytes, fp) == bytes;
    ok &= std::fclose(fp) == 0;
    return ok && std::rename(tmp.c_str(), path.c_str()) == 0;      // readers never see a partial file
}
static bool read_file_wait(const std::string& path, std::vector<char>& out, double timeoutS) {
    const double t0 = now_s();
    for (;;) {
        FILE* fp = std::fopen(path.c_str(), "rb");
        if (fp) {
            std::fseek(fp, 0, SEEK_END);
            long sz = std::ftell(fp);
            std::fseek(fp, 0, SEEK_SET);
            out.resize(sz > 0 ? (size_t)sz : 0);
            bool ok = sz <= 0 || std::fread(out.data(), 1, (size_t)sz, fp) == (size_t)sz;
            std::fclose(fp);
            return ok;
        }
        if (now_s() - t0 > timeoutS) return false;
        usleep(2000);
    }
}

static const char* kDumpArrays[] = {"points", "normals", "sorted_idx", "sorted_key", "base", "count", "key", "pidx", "pnum", "parent", "didx", "dnum", "child0", "neighs", "p2n",
                                    "vectorfield", "divergence", "x", "pointvalue", "iso", "center_scale", "cg_iters", "passes", "subdivide", "mesh_layout"};

static int dump_arrays(prb_context* ctx, const std::string& dir) {
    mkdir(dir.c_str(), 0777);
    for (const char* name : kDumpArrays) {
        const int64_t nb = prb_get_array(ctx, name, nullptr, 0);
        if (nb < 0) continue;
        std::vector<char> buf((size_t)nb);
        if (nb && prb_get_array(ctx, name, buf.data(), nb) < 0) { std::fprintf(stderr, "dump %s: %s\n", name, prb_last_error()); return 1; }
        if (!write_file(dir + "/" + name + ".bin", buf.data(), buf.size())) { std::fprintf(stderr, "dump: cannot write %s/%s.bin\n", dir.c_str(), name); return 1; }
    }
    return 0;
}

// One rank of the reconstruction.  world == 1: the plain single-GPU path.  Returns the pieces of the mesh this rank holds.
struct RankResult {
    std::vector<int64_t> layout;            // [pieces][5]
    std::vector<float> v;
    std::vector<int32_t> t;
    prb_stats s;
    double createS = 0, computeS = 0;
};
static int run_rank(const Args& a, int rank, int world, const std::string& dir, const float* xyz, const float* nrm, int64_t n, RankResult& R) {
    const double t1 = now_s();
    prb_context* ctx = nullptr;
    if (prb_create(a.device + rank, a.depth, &ctx) != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
    prb_set_option(ctx, "refine", a.refine);
    prb_set_option(ctx, "cascadic", a.cascadic);
    prb_set_option(ctx, "iso_density_weighted", a.densityIso);
    if (world > 1) {
        const double gb = a.arenaGb > 0 ? a.arenaGb : (330.0 * (double)n + (double)(64 << 20)) / (double)(1 << 30);
        char mine[64];
        if (prb_mg_init(ctx, rank, world, (int64_t)(gb * (double)(1 << 30)), mine) != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
        if (!write_file(dir + "/h." + std::to_string(rank), mine, 64)) { std::fprintf(stderr, "[rank %d] cannot publish the arena handle\n", rank); return 1; }
        for (int q = 0; q < world; q++) {
            if (q == rank) continue;
            std::vector<char> h;
            if (!read_file_wait(dir + "/h." + std::to_string(q), h, 120.0) || h.size() != 64) { std::fprintf(stderr, "[rank %d] no arena handle from rank %d\n", rank, q); return 1; }
            if (prb_mg_set_peer(ctx, q, h.data()) != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
        }
        if (prb_mg_barrier(ctx) != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
    }
    const double t2 = now_s();
    const int64_t s0 = (n * rank) / world;
    int rc = world > 1 ? prb_set_points_sharded(ctx, xyz + 3 * s0, nrm + 3 * s0, n) : prb_set_points(ctx, xyz, nrm, n);
    if (rc == PRB_OK) rc = prb_run(ctx);
    if (rc != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
    const float* v = nullptr;
    const int32_t* t = nullptr;
    int64_t nv = 0, nt = 0;
    if (prb_get_mesh(ctx, &v, &nv, &t, &nt) != PRB_OK) { std::fprintf(stderr, "[rank %d] %s\n", rank, prb_last_error()); return 1; }
    const double t3 = now_s();
    prb_get_stats(ctx, &R.s);
    const int64_t lb = prb_get_array(ctx, "mesh_layout", nullptr, 0);
    R.layout.resize(lb > 0 ? (size_t)lb / 8 : 0);
    if (lb > 0) prb_get_array(ctx, "mesh_layout", R.layout.data(), lb);
    R.v.assign(v, v + 3 * nv);
    R.t.assign(t, t + 3 * nt);
    R.createS = t2 - t1;
    R.computeS = t3 - t2;
    if (rank == 0 && !a.dump.empty() && dump_arrays(ctx, a.dump) != 0) return 1;
    if (world > 1) prb_mg_barrier(ctx);        // nobody tears its arena down while a peer may still be reading it
    prb_destroy(ctx);
    return 0;
}

static void place(const RankResult& R, std::vector<float>& V, std::vector<int32_t>& T) {
    int64_t av = 0, at = 0;
    for (size_t k = 0; k + 5 <= R.layout.size(); k += 5) {
        const int64_t vb = R.layout[k + 1], nv = R.layout[k + 2], tb = R.layout[k + 3], nt = R.layout[k + 4];
        if (nv) std::memcpy(&V[3 * (size_t)vb], &R.v[3 * (size_t)av], 12 * (size_t)nv);
        if (nt) std::memcpy(&T[3 * (size_t)tb], &R.t[3 * (size_t)at], 12 * (size_t)nt);
        av += nv; at += nt;
    }
}

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (s == "--in") a.in = val("--in");
        else if (s == "--out") a.out = val("--out");
        else if (s == "--depth") a.depth = std::atoi(val("--depth"));
        else if (s == "--device") a.device = std::atoi(val("--device"));
        else if (s == "--gpus") a.gpus = std::atoi(val("--gpus"));
        else if (s == "--dump") a.dump = val("--dump");
        else if (s == "--arena-gb") a.arenaGb = std::atof(val("--arena-gb"));
        else if (s == "--binary") a.binary = 1;
        else if (s == "--weld") a.weld = 1;
        else if (s == "--no-refine") a.refine = 0;
        else if (s == "--cascadic") a.cascadic = 1;
        else if (s == "--density-weighted-iso") a.densityIso = 1;
        else if (s == "--json") a.json = 1;
        else if (s == "--help" || s == "-h") { usage(argv[0]); return 0; }
        else { std::fprintf(stderr, "unknown argument %s\n", s.c_str()); usage(argv[0]); return 2; }
    }
    if (a.in.empty() || a.out.empty() || a.gpus < 1 || a.gpus > 8) { usage(argv[0]); return 2; }
    if (a.cascadic && a.gpus > 1) { std::fprintf(stderr, "--cascadic is a single-GPU mode\n"); return 2; }
    const double t0 = now_s();
    float *xyz = nullptr, *nrm = nullptr;
    int64_t n = 0;
    if (prbio_read_points(a.in.c_str(), &xyz, &nrm, &n) != 0) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const double t1 = now_s();
    std::printf("Total points number:%lld ,Read takes:%lfs\n", (long long)n, t1 - t0);
    if (n <= 0) { std::fprintf(stderr, "no points in %s\n", a.in.c_str()); return 1; }
    std::fflush(stdout);
    // ---- ranks (no CUDA call has been made yet: fork is safe)
    std::string dir;
    std::vector<pid_t> kids;
    int rank = 0;
    if (a.gpus > 1) {
        char tmpl[] = "/tmp/prb_mg_XXXXXX";
        if (!mkdtemp(tmpl)) { std::perror("mkdtemp"); return 1; }
        dir = tmpl;
        for (int r = 1; r < a.gpus; r++) {
            pid_t pid = fork();
            if (pid < 0) { std::perror("fork"); return 1; }
            if (pid == 0) { rank = r; kids.clear(); break; }
            kids.push_back(pid);
        }
    }
    RankResult R;
    int rc = run_rank(a, rank, a.gpus, dir, xyz, nrm, n, R);
    if (rank != 0) {
        // worker: hand the pieces to rank 0 and leave
        if (rc == 0) {
            std::vector<char> blob;
            auto put = [&](const void* p, size_t b) { blob.insert(blob.end(), (const char*)p, (const char*)p + b); };
            int64_t hdr[3] = {(int64_t)R.layout.size(), (int64_t)R.v.size(), (int64_t)R.t.size()};
            put(hdr, sizeof(hdr)); put(R.layout.data(), 8 * R.layout.size()); put(R.v.data(), 4 * R.v.size()); put(R.t.data(), 4 * R.t.size());
            if (!write_file(dir + "/mesh." + std::to_string(rank), blob.data(), blob.size())) rc = 1;
        }
        _exit(rc);
    }
    bool ok = rc == 0;
    for (pid_t pid : kids) {
        int st = 0;
        waitpid(pid, &st, 0);
        ok &= WIFEXITED(st) && WEXITSTATUS(st) == 0;
    }
    if (!ok) { std::fprintf(stderr, "reconstruction failed\n"); return 1; }
    const prb_stats& s = R.s;
    std::vector<float> V;
    std::vector<int32_t> T;
    int64_t nv = s.n_vertices, nt = s.n_triangles;
    if (a.gpus > 1) {
        V.assign(3 * (size_t)nv, 0.f);
        T.assign(3 * (size_t)nt, 0);
        place(R, V, T);
        for (int r = 1; r < a.gpus; r++) {
            std::vector<char> blob;
            if (!read_file_wait(dir + "/mesh." + std::to_string(r), blob, 5.0) || blob.size() < 24) { std::fprintf(stderr, "no mesh pieces from rank %d\n", r); return 1; }
            int64_t hdr[3];
            std::memcpy(hdr, blob.data(), 24);
            RankResult Q;
            Q.layout.resize((size_t)hdr[0]); Q.v.resize((size_t)hdr[1]); Q.t.resize((size_t)hdr[2]);
            size_t off = 24;
            std::memcpy(Q.layout.data(), blob.data() + off, 8 * Q.layout.size()); off += 8 * Q.layout.size();
            std::memcpy(Q.v.data(), blob.data() + off, 4 * Q.v.size()); off += 4 * Q.v.size();
            std::memcpy(Q.t.data(), blob.data() + off, 4 * Q.t.size());
            place(Q, V, T);
            std::remove((dir + "/mesh." + std::to_string(r)).c_str());
        }
        for (int r = 0; r < a.gpus; r++) std::remove((dir + "/h." + std::to_string(r)).c_str());
        rmdir(dir.c_str());
    } else {
        V.swap(R.v);
        T.swap(R.t);
    }
    const double t3 = now_s();
    std::printf("NodeArray_sz:%d\n", s.n_nodes);
    std::printf("GPU NodeArray build takes:%lfs\n", (s.ms_h2d + s.ms_octree) * 1e-3);
    std::printf("Compute Vector Field takes:%lfs\n", s.ms_splat * 1e-3);
    std::printf("Compute nodes' divergence takes:%lfs\n", s.ms_divergence * 1e-3);
    std::printf("GPU Laplacian Iteration takes:%lfs\n", s.ms_solve * 1e-3);
    std::printf("isoValue:%f\nGPU calculate isoValue takes:%lfs\n", s.iso_value, s.ms_iso * 1e-3);
    std::printf("SubdivideNum:%d\n", s.n_subdivide);
    std::printf("GPU marching cubes + subdivide passes takes:%lfs (%d passes)\n", s.ms_extract * 1e-3, s.n_passes);
    if (a.weld) {
        int64_t nv2 = nv;
        if (prbio_weld_mesh(V.data(), nv, T.data(), nt, &nv2) != 0) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
        std::printf("Weld: %lld -> %lld vertices\n", (long long)nv, (long long)nv2);
        nv = nv2;
    }
    std::printf("Vertices:%lld Triangles:%lld\n", (long long)nv, (long long)nt);
    if (prbio_write_mesh(a.out.c_str(), V.data(), nv, T.data(), nt, s.center, s.scale, a.binary) != 0) { std::fprintf(stderr, "%s\n", prb_last_error()); return 1; }
    const double t4 = now_s();
    std::printf("Output ply files takes %lfs\n", t4 - t3);
    std::printf("The whole project takes %lfs (including I/O)\n", t4 - t0);
    if (a.json) {
        std::printf("{\"n_points\": %lld, \"depth\": %d, \"gpus\": %d, \"n_nodes\": %d, \"n_vertices\": %lld, \"n_triangles\": %lld, \"iso\": %.9g, "
                    "\"read_s\": %.6f, \"create_s\": %.6f, \"compute_s\": %.6f, \"device_ms\": %.3f, \"write_s\": %.6f, \"total_s\": %.6f}\n",
                    (long long)n, a.depth, a.gpus, s.n_nodes, (long long)nv, (long long)nt, (double)s.iso_value, t1 - t0, R.createS, R.computeS, (double)s.ms_total, t4 - t3, t4 - t0);
    }
    prbio_free(xyz);
    prbio_free(nrm);
    return 0;
}
