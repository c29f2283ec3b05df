// C ABI (include/prb.h): context management, staging, getters.  No torch types, no exceptions
// across the boundary.  Everything runs on the context's own non-default stream.
#include "common.cuh"
#include <cstring>
#include <mutex>

namespace prb {

static thread_local std::string g_lastError;
void set_error(const std::string& msg) { g_lastError = msg; }

static void release_all(Context& c) {
    c.rawP.release(); c.rawN.release(); c.P.release(); c.Nr.release(); c.sortedKey.release(); c.sortedIdx.release(); c.p2n.release();
    c.dBase.release(); c.key.release(); c.parent.release(); c.child0.release(); c.pidx.release(); c.pnum.release(); c.didx.release(); c.dnum.release();
    c.neighs.release(); c.offs.release(); c.sgTab.release(); c.sgTab4.release();
    c.V.release(); c.divg.release(); c.x.release(); c.pointValue.release();
    c.meshV.release(); c.meshT.release(); c.vval.release();
    c.nMeshV = c.nMeshT = 0;
    c.nGlobalV = c.nGlobalT = 0;
    c.xv = nullptr;
    c.divgv = nullptr;
    c.hMeshValid = false;
    if (c.copyStream) cudaStreamSynchronize(c.copyStream);      // (an early mesh copy of the previous run that nobody fetched)
    c.earlyV = c.earlyT = 0;
}

int ensure_copy_stream(Context& c) {
    if (c.copyStream) return PRB_OK;
    PRB_CUDA(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
    PRB_CUDA(cudaEventCreateWithFlags(&c.evMainPiece, cudaEventDisableTiming));
    PRB_CUDA(cudaEventCreateWithFlags(&c.evEarlyCopy, cudaEventDisableTiming));
    PRB_CUDA(cudaEventCreateWithFlags(&c.evPositions, cudaEventDisableTiming));
    PRB_CUDA(cudaEventCreateWithFlags(&c.evNormals, cudaEventDisableTiming));
    return PRB_OK;
}

int upload_tables(Context& c) {
    cudaStream_t st = c.stream;
    build_bspline_tables(c.D, c.tab);
    PRB_TRY(c.dMaxDepthFn.alloc(16, st));
    PRB_TRY(c.dBaseFn.alloc(c.tab.baseFn.size(), st));
    PRB_TRY(c.dDfT.alloc(c.tab.dfT.size(), st));
    PRB_TRY(c.dDfOffset.alloc(c.tab.dfOffset.size(), st));
    PRB_TRY(c.dStencil.alloc(c.tab.stencil.size(), st));
    PRB_CUDA(cudaMemcpyAsync(c.dMaxDepthFn.p, c.tab.maxDepthFn, 64, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(c.dBaseFn.p, c.tab.baseFn.data(), c.dBaseFn.bytes(), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(c.dDfT.p, c.tab.dfT.data(), c.dDfT.bytes(), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(c.dDfOffset.p, c.tab.dfOffset.data(), c.dDfOffset.bytes(), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(c.dStencil.p, c.tab.stencil.data(), c.dStencil.bytes(), cudaMemcpyHostToDevice, st));
    PRB_TRY(c.dFfX.alloc(c.tab.ffX.size(), st));
    PRB_TRY(c.dD2X.alloc(c.tab.d2X.size(), st));
    PRB_TRY(c.dCrossOff.alloc(c.tab.crossOff.size(), st));
    if (!c.tab.ffX.empty()) {
        PRB_CUDA(cudaMemcpyAsync(c.dFfX.p, c.tab.ffX.data(), c.dFfX.bytes(), cudaMemcpyHostToDevice, st));
        PRB_CUDA(cudaMemcpyAsync(c.dD2X.p, c.tab.d2X.data(), c.dD2X.bytes(), cudaMemcpyHostToDevice, st));
    }
    PRB_CUDA(cudaMemcpyAsync(c.dCrossOff.p, c.tab.crossOff.data(), c.dCrossOff.bytes(), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    return PRB_OK;
}

}  // namespace prb

using namespace prb;

extern "C" {

const char* prb_last_error(void) { return g_lastError.c_str(); }

int prb_create(int device, int depth, prb_context** out) {
    if (!out) { set_error("out is null"); return PRB_ERR_ARG; }
    *out = nullptr;
    if (depth < 2 || depth > kMaxDepth) { set_error("depth must be in [2,12]"); return PRB_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return PRB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device index out of range"); return PRB_ERR_ARG; }
    int prevDev = -1;
    if (cudaGetDevice(&prevDev) != cudaSuccess) { cudaGetLastError(); prevDev = -1; }
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prevDev};
    PRB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PRB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error(std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only"); return PRB_ERR_CUDA; }
    prb_context* h = new prb_context();
    Context& c = h->c;
    c.device = device;
    c.D = depth;
    c.smCount = prop.multiProcessorCount;
    c.deviceMemBytes = prop.totalGlobalMem;
    for (auto& e : c.ev) e = nullptr;
    auto fail = [&](int code) {            // nothing of a half-built context survives
        for (auto& e : c.ev) if (e) cudaEventDestroy(e);
        if (c.stream) { arena_unregister(c.stream); cudaStreamDestroy(c.stream); }
        c.stream = nullptr;
        delete h;
        return code;
    };
    if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); c.stream = nullptr; set_error("cudaStreamCreate failed"); return fail(PRB_ERR_CUDA); }
    arena_register(c.stream);
    for (auto& e : c.ev)
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); e = nullptr; set_error("cudaEventCreate failed"); return fail(PRB_ERR_CUDA); }
    std::memset(&c.stats, 0, sizeof(c.stats));
    int r = upload_tables(c);
    if (r != PRB_OK) {
        c.dMaxDepthFn.release(); c.dBaseFn.release(); c.dDfT.release(); c.dDfOffset.release(); c.dStencil.release();
        c.dFfX.release(); c.dD2X.release(); c.dCrossOff.release();
        return fail(r);
    }
    *out = h;
    return PRB_OK;
}

void prb_destroy(prb_context* h) {
    if (!h) return;
    Context& c = h->c;
    DeviceGuard guard__(c.device);
    release_all(c);
    c.wsVal7.release(); c.wsLow.release(); c.wsCat.release(); c.wsNtri.release(); c.wsEmask.release(); c.wsVpre.release(); c.wsVbase.release(); c.wsTbase.release();
    c.dFfX.release(); c.dD2X.release(); c.dCrossOff.release(); c.bCas.release();
    c.dBvAnc.release(); c.dBvOwn.release(); c.dBvCell.release(); c.dBvGrid.release(); c.dMaxDepthFn.release(); c.dBaseFn.release(); c.dDfT.release(); c.dDfOffset.release(); c.dStencil.release();
    c.scanDesc.release(); c.scanTicket.release();
    if (c.hScanTotal) cudaFreeHost(c.hScanTotal);
    c.hScanTotal = nullptr;
    cudaStreamSynchronize(c.stream);
    arena_unregister(c.stream);
    for (int r = 0; r < kMaxRanks; r++)
        if (c.mg.peerOpen[r] && c.mg.peer[r]) cudaIpcCloseMemHandle(c.mg.peer[r]);
    if (c.mg.arena) cudaFree(c.mg.arena);
    if (c.copyStream) { cudaStreamSynchronize(c.copyStream); cudaStreamDestroy(c.copyStream); }
    for (cudaEvent_t e : {c.evMainPiece, c.evEarlyCopy, c.evPositions, c.evNormals})
        if (e) cudaEventDestroy(e);
    c.hMeshV.release(); c.hMeshT.release();
    for (auto& e : c.ev) cudaEventDestroy(e);
    for (auto& e : c.detailEv) cudaEventDestroy(e);
    cudaStreamDestroy(c.stream);
    delete h;
}

int prb_set_option(prb_context* h, const char* key, double value) {
    if (!h || !key) { set_error("null argument"); return PRB_ERR_ARG; }
    std::string k(key);
    if (k == "cg_tol") h->c.cgTol = value;
    else if (k == "cg_max_iter") h->c.cgMaxIter = (int)value;
    else if (k == "cg_zigzag") h->c.cgZigzag = (int)value;
    else if (k == "cg_bulk") h->c.cgBulk = (int)value;
    else if (k == "cg_timing") h->c.cgTiming = (int)value;
    else if (k == "iso_density_weighted") h->c.isoDensityWeighted = (int)value;
    else if (k == "cascadic") h->c.cascadic = (int)value;
    else if (k == "detail") h->c.detail = (int)value;
    else if (k == "early_mesh_copy") h->c.earlyMeshCopy = (int)value;
    else if (k == "mg_timeout_ms") h->c.mg.spinCycles = (long long)((value < 1.0 ? 1.0 : value) * 2.0e6);
    else if (k == "refine_bound_check") h->c.refineBoundCheck = (int)value;
    else if (k == "div_mode") h->c.divMode = (int)value;
    else if (k == "refine") h->c.doRefine = (int)value;
    else if (k == "refine_implicit") h->c.refineImplicit = (int)value;
    else { set_error("unknown option " + k); return PRB_ERR_ARG; }
    return PRB_OK;
}

static int begin_run(Context& c, int64_t n) {
    release_all(c);
    c.mg.reset_allocs();
    c.mgP = c.mgX = c.mgVval = nullptr;
    c.vvalPtr = nullptr;
    c.rawPp = c.rawNp = c.Vp = nullptr;
    c.rawSharded = false;
    c.normalsPending = false;
    c.N = n;
    c.launches = 0;
    c.detailUsed = 0;
    mark(c, "set_points");
    // a peer wait that timed out in an earlier run (first-run skew: lazy module loading, large allocations) must not poison this one
    if (c.mg.active() && c.mg.arena) PRB_CUDA(cudaMemsetAsync(&((MgHeader*)c.mg.arena)->error, 0, sizeof(int), c.stream));
    return PRB_OK;
}

int prb_set_points(prb_context* h, const float* xyz, const float* normals, int64_t n) {
    if (!h || !xyz || !normals || n <= 0) { set_error("prb_set_points: bad argument"); return PRB_ERR_ARG; }
    Context& c = h->c;
    PRB_DEVICE(c);
    PRB_TRY(begin_run(c, n));
    PRB_CUDA(cudaEventRecord(c.ev[0], c.stream));
    PRB_TRY(c.rawP.alloc(3 * (size_t)n, c.stream));
    PRB_TRY(c.rawN.alloc(3 * (size_t)n, c.stream));
    c.rawPp = c.rawP.p; c.rawNp = c.rawN.p;
    // positions first, on the context stream; the normals follow on the copy stream (ordered behind the positions, so the two uploads
    // do not share the link) while the bounding box, the keys and the first sort passes already run: the gather of the last sort pass
    // is the first reader of the normals
    PRB_TRY(ensure_copy_stream(c));
    PRB_CUDA(cudaMemcpyAsync(c.rawPp, xyz, 12 * (size_t)n, cudaMemcpyDefault, c.stream));
    PRB_CUDA(cudaEventRecord(c.evPositions, c.stream));
    PRB_CUDA(cudaStreamWaitEvent(c.copyStream, c.evPositions, 0));
    PRB_CUDA(cudaMemcpyAsync(c.rawNp, normals, 12 * (size_t)n, cudaMemcpyDefault, c.copyStream));
    PRB_CUDA(cudaEventRecord(c.evNormals, c.copyStream));
    c.normalsPending = true;
    PRB_CUDA(cudaEventRecord(c.ev[1], c.stream));
    c.stage = 1;
    return PRB_OK;
}

// Multi-GPU: rank r passes only ITS slice of the cloud, the samples [plan[r], plan[r+1]) of prb_mg_plan(n_total, world); the slices
// are gathered over NVLink at the start of prb_build_octree (every rank builds the same octree from the same n_total samples).
int prb_set_points_sharded(prb_context* h, const float* xyz_slice, const float* normals_slice, int64_t n_total) {
    if (!h || !xyz_slice || !normals_slice || n_total <= 0) { set_error("prb_set_points_sharded: bad argument"); return PRB_ERR_ARG; }
    Context& c = h->c;
    if (!c.mg.active()) return prb_set_points(h, xyz_slice, normals_slice, n_total);
    PRB_DEVICE(c);
    PRB_TRY(begin_run(c, n_total));
    PRB_CUDA(cudaEventRecord(c.ev[0], c.stream));
    c.rawPp = c.mg.alloc<float>(3 * (size_t)n_total, &c.mgRawPOff);
    c.rawNp = c.mg.alloc<float>(3 * (size_t)n_total, &c.mgRawNOff);
    if (!c.rawPp || !c.rawNp) { set_error("multi-GPU arena too small for the samples (24 bytes per sample; prb_mg_init arena_bytes)"); return PRB_ERR_NOMEM; }
    const int64_t a = (n_total * c.mg.rank) / c.mg.world, b = (n_total * (c.mg.rank + 1)) / c.mg.world;
    if (b > a) {
        PRB_CUDA(cudaMemcpyAsync(c.rawPp + 3 * a, xyz_slice, 12 * (size_t)(b - a), cudaMemcpyDefault, c.stream));
        PRB_CUDA(cudaMemcpyAsync(c.rawNp + 3 * a, normals_slice, 12 * (size_t)(b - a), cudaMemcpyDefault, c.stream));
    }
    PRB_CUDA(cudaEventRecord(c.ev[1], c.stream));
    c.rawSharded = true;
    c.stage = 1;
    return PRB_OK;
}

int prb_build_octree(prb_context* h) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 1) { set_error("prb_build_octree: no points set"); return PRB_ERR_STATE; }
    PRB_DEVICE(c);
    PRB_CUDA(cudaEventRecord(c.ev[1], c.stream));
    PRB_TRY(stage_octree(c));
    PRB_CUDA(cudaEventRecord(c.ev[2], c.stream));
    c.stage = 2;
    return PRB_OK;
}
int prb_splat(prb_context* h) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 2) { set_error("prb_splat: octree not built"); return PRB_ERR_STATE; }
    PRB_DEVICE(c);
    PRB_CUDA(cudaEventRecord(c.ev[2], c.stream));
    PRB_TRY(stage_splat(c));
    PRB_CUDA(cudaEventRecord(c.ev[4], c.stream));
    c.stage = 3;
    return PRB_OK;
}
int prb_solve(prb_context* h) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 3) { set_error("prb_solve: splat not done"); return PRB_ERR_STATE; }
    PRB_DEVICE(c);
    PRB_CUDA(cudaEventRecord(c.ev[4], c.stream));
    PRB_TRY(stage_solve(c));
    PRB_CUDA(cudaEventRecord(c.ev[5], c.stream));
    PRB_TRY(stage_iso(c));
    PRB_CUDA(cudaEventRecord(c.ev[6], c.stream));
    c.stage = 4;
    return PRB_OK;
}
int prb_extract(prb_context* h) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 4) { set_error("prb_extract: solve not done"); return PRB_ERR_STATE; }
    PRB_DEVICE(c);
    PRB_CUDA(cudaEventRecord(c.ev[6], c.stream));
    PRB_TRY(stage_extract(c));
    PRB_CUDA(cudaEventRecord(c.ev[7], c.stream));
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    c.stage = 5;
    return PRB_OK;
}
int prb_run(prb_context* h) {
    PRB_TRY(prb_build_octree(h));
    PRB_TRY(prb_splat(h));
    PRB_TRY(prb_solve(h));
    PRB_TRY(prb_extract(h));
    return PRB_OK;
}

// Debug / parity: re-run one stage on the current (possibly overwritten) intermediates.
int prb_run_stage(prb_context* h, const char* name) {
    if (!h || !name) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    std::string s(name);
    int need = (s == "divergence") ? 3 : (s == "solve") ? 3 : (s == "iso") ? 4 : (s == "extract") ? 4 : 99;
    if (c.stage < need) { set_error("prb_run_stage: stage not reachable yet"); return PRB_ERR_STATE; }
    if (s == "divergence") PRB_TRY(stage_divergence(c));
    else if (s == "solve") PRB_TRY(stage_solve(c));
    else if (s == "iso") PRB_TRY(stage_iso(c));
    else if (s == "extract") { PRB_TRY(stage_extract(c)); if (c.stage < 5) c.stage = 5; }
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    return PRB_OK;
}

int prb_get_mesh_device(prb_context* h, const float** v, int64_t* nv, const int32_t** t, int64_t* nt) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 5) { set_error("prb_get_mesh: extract not done"); return PRB_ERR_STATE; }
    if (v) *v = c.meshV.p;
    if (nv) *nv = c.nMeshV;
    if (t) *t = c.meshT.p;
    if (nt) *nt = c.nMeshT;
    return PRB_OK;
}
int prb_get_mesh(prb_context* h, const float** v, int64_t* nv, const int32_t** t, int64_t* nt) {
    if (!h) return PRB_ERR_ARG;
    Context& c = h->c;
    if (c.stage < 5) { set_error("prb_get_mesh: extract not done"); return PRB_ERR_STATE; }
    PRB_DEVICE(c);
    if (!c.hMeshValid) {
        // the leading piece may already be in the pinned buffers (stage_extract started its copy under the refinement passes); it is
        // kept when the buffers did not have to grow
        if (c.copyStream) PRB_CUDA(cudaStreamSynchronize(c.copyStream));
        const unsigned gv0 = c.hMeshV.generation, gt0 = c.hMeshT.generation;
        PRB_TRY(c.hMeshV.reserve(3 * (size_t)c.nMeshV + 1));
        PRB_TRY(c.hMeshT.reserve(3 * (size_t)c.nMeshT + 1));
        const bool kept = c.hMeshV.generation == gv0 && c.hMeshT.generation == gt0 && c.earlyV <= c.nMeshV && c.earlyT <= c.nMeshT;
        const size_t ev = kept ? (size_t)c.earlyV : 0, et = kept ? (size_t)c.earlyT : 0;
        if ((size_t)c.nMeshV > ev) PRB_CUDA(cudaMemcpyAsync(c.hMeshV.p + 3 * ev, c.meshV.p + 3 * ev, 12 * ((size_t)c.nMeshV - ev), cudaMemcpyDeviceToHost, c.stream));
        if ((size_t)c.nMeshT > et) PRB_CUDA(cudaMemcpyAsync(c.hMeshT.p + 3 * et, c.meshT.p + 3 * et, 12 * ((size_t)c.nMeshT - et), cudaMemcpyDeviceToHost, c.stream));
        PRB_CUDA(cudaStreamSynchronize(c.stream));
        c.hMeshValid = true;
    }
    if (v) *v = c.hMeshV.p;
    if (nv) *nv = c.nMeshV;
    if (t) *t = c.hMeshT.p;
    if (nt) *nt = c.nMeshT;
    return PRB_OK;
}

int prb_get_stream(prb_context* h, void** stream) {
    if (!h || !stream) { set_error("prb_get_stream: null argument"); return PRB_ERR_ARG; }
    *stream = (void*)h->c.stream;
    return PRB_OK;
}

int prb_get_stats(prb_context* h, prb_stats* out) {
    if (!h || !out) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    prb_stats& s = c.stats;
    s.n_points = c.N;
    s.depth = c.D;
    s.n_nodes = c.M;
    for (int d = 0; d < 16; d++) { s.nodes_per_depth[d] = d <= c.D ? c.cnt[d] : 0; s.cg_iters[d] = d <= c.D ? c.cgIters[d] : 0; }
    s.n_subdivide = (int)c.subdivide.size();
    s.n_passes = (int)c.passes.size();
    s.n_vertices = c.nGlobalV;       // the whole mesh (multi-GPU: this context holds the pieces listed by "mesh_layout")
    s.n_triangles = c.nGlobalT;
    s.iso_value = c.iso;
    for (int a = 0; a < 3; a++) s.center[a] = c.center[a];
    s.scale = c.scale;
    s.cg_row_iters = c.cgRowIters;
    s.kernel_launches = c.launches;
    auto el = [&](int a, int b) { float ms = 0; if (cudaEventElapsedTime(&ms, c.ev[a], c.ev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0; } return ms; };
    if (c.stage >= 1) s.ms_h2d = el(0, 1);
    if (c.stage >= 2) s.ms_octree = el(1, 2);
    if (c.stage >= 3) { s.ms_splat = el(2, 3); s.ms_divergence = el(3, 4); }
    if (c.stage >= 4) { s.ms_solve = el(4, 5); s.ms_iso = el(5, 6); }
    if (c.stage >= 5) { s.ms_extract = el(6, 7); s.ms_total = el(0, 7); }
    *out = s;
    return PRB_OK;
}

int64_t prb_get_array(prb_context* h, const char* name, void* dst, int64_t cap) {
    if (!h || !name) return PRB_ERR_ARG;
    Context& c = h->c;
    DeviceGuard guard__(c.device);
    if (!guard__.ok) return PRB_ERR_CUDA;
    cudaStreamSynchronize(c.stream);
    std::string s(name);
    const void* dev = nullptr;
    int64_t bytes = -1;
    std::vector<char> host;
    auto D_ = [&](const void* p, size_t b) { dev = p; bytes = (int64_t)b; };
    auto H_ = [&](const void* p, size_t b) { host.assign((const char*)p, (const char*)p + b); bytes = (int64_t)b; };
    const int M = c.M, D = c.D;
    if (s == "points") D_(c.P.p, c.P.bytes());
    else if (s == "normals") D_(c.Nr.p, c.Nr.bytes());
    else if (s == "sorted_idx") D_(c.sortedIdx.p, c.sortedIdx.bytes());
    else if (s == "sorted_key") D_(c.sortedKey.p, c.sortedKey.bytes());
    else if (s == "base") H_(c.base, sizeof(int) * (D + 2));
    else if (s == "count") H_(c.cnt, sizeof(int) * (D + 1));
    else if (s == "key") D_(c.key.p, c.key.bytes());
    else if (s == "pidx") D_(c.pidx.p, c.pidx.bytes());
    else if (s == "pnum") D_(c.pnum.p, c.pnum.bytes());
    else if (s == "parent") D_(c.parent.p, c.parent.bytes());
    else if (s == "didx") D_(c.didx.p, c.didx.bytes());
    else if (s == "dnum") D_(c.dnum.p, c.dnum.bytes());
    else if (s == "child0") D_(c.child0.p, c.child0.bytes());
    else if (s == "neighs") D_(c.neighs.p, c.neighs.bytes());
    else if (s == "sg_table") D_(c.sgTab.p, c.sgTab.bytes());
    else if (s == "p2n") D_(c.p2n.p, c.p2n.bytes());
    else if (s == "vectorfield") D_(c.Vp, c.Vp ? 12 * (size_t)c.cnt[D] : 0);
    else if (s == "divergence") D_(c.divgv, c.divgv ? 4 * (size_t)M : 0);
    else if (s == "x") D_(c.xv, c.xv ? 4 * (size_t)M : 0);
    else if (s == "pointvalue") D_(c.pointValue.p, c.pointValue.bytes());
    else if (s == "vvalue_slots") D_(c.vvalPtr, c.vvalPtr ? 32 * (size_t)M : 0);
    else if (s == "mesh_v") D_(c.meshV.p, 12 * (size_t)c.nMeshV);
    else if (s == "mesh_t") D_(c.meshT.p, 12 * (size_t)c.nMeshT);
    else if (s == "iso") H_(&c.iso, 4);
    else if (s == "iso_modes") { float v[2] = {c.isoPlain, c.isoWeighted}; H_(v, 8); }
    else if (s == "center_scale") { float v[4] = {c.center[0], c.center[1], c.center[2], c.scale}; H_(v, 16); }
    else if (s == "cg_iters") H_(c.cgIters, sizeof(int) * (D + 1));
    else if (s == "cg_phase_ns") H_(c.cgPhaseNs, sizeof(c.cgPhaseNs));
    else if (s == "cascadic_rhs") D_(c.bCas.p ? c.bCas.p + 7 : nullptr, c.bCas.p ? 4 * (size_t)M : 0);
    else if (s == "detail_ms") {
        std::vector<float> ms;
        for (size_t k = 0; k + 1 < c.detailUsed; k++) { float v = 0; if (cudaEventElapsedTime(&v, c.detailEv[k], c.detailEv[k + 1]) != cudaSuccess) { cudaGetLastError(); v = -1; } ms.push_back(v); }
        H_(ms.data(), ms.size() * 4);
    } else if (s == "detail_names") {
        std::string all;
        for (size_t k = 0; k + 1 < c.detailUsed; k++) { all += c.detailName[k] + " -> " + c.detailName[k + 1]; all.push_back('\0'); }
        H_(all.data(), all.size());
    }
    else if (s == "lap_stencil") H_(c.tab.stencil.data(), c.tab.stencil.size() * 4);
    else if (s == "df_table") H_(c.tab.dfT.data(), c.tab.dfT.size() * 4);
    else if (s == "passes") { std::vector<int> v; for (auto& p : c.passes) { v.push_back(p.kind); v.push_back(p.nv); v.push_back(p.nt); } H_(v.data(), v.size() * 4); }
    else if (s == "subdivide") H_(c.subdivide.data(), c.subdivide.size() * 4);
    else if (s == "mesh_layout") H_(c.layout.data(), c.layout.size() * sizeof(Context::PieceRecord));
    else if (s == "children") {
        // expanded [M][8] view of child0 for comparison with the reference layout
        std::vector<int> c0(M);
        if (M && cudaMemcpy(c0.data(), c.child0.p, sizeof(int) * (size_t)M, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); return PRB_ERR_CUDA; }
        std::vector<int> ch(8 * (size_t)M);
        for (int i = 0; i < M; i++) for (int k = 0; k < 8; k++) ch[8 * (size_t)i + k] = c0[i] < 0 ? -1 : c0[i] + k;
        H_(ch.data(), ch.size() * 4);
    } else { set_error("unknown array " + s); return PRB_ERR_ARG; }
    if (dst && cap >= bytes && bytes > 0) {
        if (dev) { if (cudaMemcpy(dst, dev, (size_t)bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy failed"); return PRB_ERR_CUDA; } }
        else std::memcpy(dst, host.data(), (size_t)bytes);
    }
    return bytes;
}

int64_t prb_host_tables(int depth, const char* name, void* dst, int64_t cap) {
    if (!name || depth < 0 || depth > kMaxDepth) { set_error("prb_host_tables: bad argument"); return PRB_ERR_ARG; }
    BSplineTables t;
    build_bspline_tables(depth, t);
    std::string s(name);
    const void* src = nullptr;
    size_t bytes = 0;
    if (s == "gauss") { src = t.gauss; bytes = sizeof(t.gauss); }
    else if (s == "max_depth_fn") { src = t.maxDepthFn; bytes = sizeof(t.maxDepthFn); }
    else if (s == "base_fn") { src = t.baseFn.data(); bytes = t.baseFn.size() * 4; }
    else if (s == "df_table") { src = t.dfT.data(); bytes = t.dfT.size() * 4; }
    else if (s == "df_offset") { src = t.dfOffset.data(); bytes = t.dfOffset.size() * 4; }
    else if (s == "stencil") { src = t.stencil.data(); bytes = t.stencil.size() * 4; }
    else if (s == "ff0") { src = t.ff0.data(); bytes = t.ff0.size() * 8; }
    else if (s == "ff1") { src = t.ff1.data(); bytes = t.ff1.size() * 8; }
    else if (s == "d20") { src = t.d20.data(); bytes = t.d20.size() * 8; }
    else if (s == "d21") { src = t.d21.data(); bytes = t.d21.size() * 8; }
    else if (s == "ff_cross") { src = t.ffX.data(); bytes = t.ffX.size() * 8; }
    else if (s == "d2_cross") { src = t.d2X.data(); bytes = t.d2X.size() * 8; }
    else if (s == "cross_offset") { src = t.crossOff.data(); bytes = t.crossOff.size() * 4; }
    else { set_error("unknown table " + s); return PRB_ERR_ARG; }
    if (dst && cap >= (int64_t)bytes && bytes) std::memcpy(dst, src, bytes);
    return (int64_t)bytes;
}

int prb_set_array(prb_context* h, const char* name, const void* src, int64_t bytes) {
    if (!h || !name || !src) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    std::string s(name);
    void* dst = nullptr;
    size_t want = 0;
    if (s == "vectorfield") { dst = c.Vp; want = c.Vp ? 12 * (size_t)c.cnt[c.D] : 0; }
    else if (s == "divergence") { dst = c.divgv; want = c.divgv ? 4 * (size_t)c.M : 0; }
    else if (s == "x") { dst = c.xv; want = c.xv ? 4 * (size_t)c.M : 0; }
    else if (s == "iso") { if (bytes != 4) return PRB_ERR_ARG; std::memcpy(&c.iso, src, 4); return PRB_OK; }
    else { set_error("unknown array " + s); return PRB_ERR_ARG; }
    if (!dst || (int64_t)want != bytes) { set_error("size mismatch for " + s); return PRB_ERR_ARG; }
    PRB_CUDA(cudaMemcpy(dst, src, want, cudaMemcpyHostToDevice));
    return PRB_OK;
}

}  // extern "C"
