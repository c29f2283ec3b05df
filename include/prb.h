/* prb.h -- C ABI of the B200-native Poisson surface reconstruction pipeline.
 *
 * Drop-in boundary.  The reference (DavidXu-JJ/PoissonRecon_GPU) has no plugin / FFI surface:
 * its only contract is "process in, files out" (main.cu:3247-3252 hard-coded paths,
 * main.cu:517-520 input dispatch by extension, main.cu:4566 PlyWriteTriangles).  This header is
 * therefore the boundary a maintainer of the reference would bind instead of calling
 * pipelineBuildNodeArray / LaplacianIteration / the inline marching-cubes stages of main():
 *
 *   reference stage (file:line)                                   entry point here
 *   ------------------------------------------------------------  ---------------------------
 *   pipelineBuildNodeArray            main.cu:511-841             prb_set_points + prb_build_octree
 *   computeVectorField + divergence   main.cu:3355-3462           prb_splat
 *   LaplacianIteration + CG           main.cu:1223-1332,          prb_solve
 *     solverCG_DeviceToDevice         CG_CUDA.cuh:344-509
 *   iso-value                         main.cu:3480-3499           prb_solve (tail)
 *   vertex/edge/face arrays, MC,      main.cu:3504-4564           prb_extract
 *     refinement passes
 *   insertTriangle / mesh container   main.cu:3220-3245           prb_get_mesh
 *   whole main()                      main.cu:3247-4573           prb_run
 *
 * Conventions: plain C, opaque handle, int status (0 = ok, <0 = error, message from
 * prb_last_error()); no exceptions cross the boundary; the caller owns input buffers; the
 * library owns outputs until prb_destroy or the next prb_set_points; one context per GPU; a
 * context is not thread-safe; all work runs on an internal non-default stream.
 * There is no CPU fallback: every entry point fails with PRB_ERR_CUDA if no sm_100 device.
 */
#ifndef PRB_H_
#define PRB_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct prb_context prb_context;

enum {
    PRB_OK = 0,
    PRB_ERR_ARG = -1,     /* bad argument (depth out of range, null pointer, n <= 0, unknown name) */
    PRB_ERR_CUDA = -2,    /* CUDA runtime failure / no usable device */
    PRB_ERR_STATE = -3,   /* stage called out of order */
    PRB_ERR_NOMEM = -4
};

/* Per-run statistics (counts are the parity probes the reference prints: NodeArray_sz
 * main.cu:3278, VertexArray_sz :3575, EdgeArray_sz :3609, SubdivideNum :3836, isoValue :3499). */
typedef struct prb_stats {
    int64_t n_points;
    int32_t depth;
    int32_t n_nodes;               /* NodeArray_sz */
    int32_t nodes_per_depth[16];   /* NodeArrayCount_h */
    int32_t cg_iters[16];          /* CG iterations per depth */
    int32_t n_subdivide;           /* SubdivideNum */
    int32_t n_passes;              /* mesh passes emitted (main + refinement) */
    int64_t n_vertices, n_triangles;
    float iso_value;
    float center[3], scale;        /* file coords = p*scale + center (plyfile.cu:2801-2803) */
    /* device time per stage, milliseconds (CUDA events on the context stream) */
    float ms_h2d, ms_octree, ms_splat, ms_divergence, ms_solve, ms_iso, ms_extract, ms_total;
    int64_t cg_row_iters;          /* sum_d rows_d * iters_d (roofline numerator) */
    int32_t kernel_launches;       /* kernels launched by the last prb_run */
} prb_stats;

/* Create a context on CUDA device `device` for octree depth `depth` (2..12; the reference's
 * compile-time `maxDepth`, main.cu:69). */
int prb_create(int device, int depth, prb_context** out);
void prb_destroy(prb_context* ctx);
const char* prb_last_error(void);

/* Oriented samples: xyz and normals as float32 [n][3], raw file coordinates.  Pointers may be
 * host (pageable or pinned) or device memory; they are copied ASYNCHRONOUSLY (positions on the
 * context stream, normals behind them on a second stream, under the key generation and the first
 * sort passes): pinned-host and device buffers must stay valid and unchanged until
 * prb_build_octree / prb_run has returned.  Replaces the two PointStream passes + H2D copies of
 * main.cu:530-581. */
int prb_set_points(prb_context* ctx, const float* xyz, const float* normals, int64_t n);

/* Staged execution (each requires the previous stage). */
int prb_build_octree(prb_context* ctx);
int prb_splat(prb_context* ctx);
int prb_solve(prb_context* ctx);
int prb_extract(prb_context* ctx);
/* All four stages. */
int prb_run(prb_context* ctx);

/* Multi-GPU: the mesh is DISTRIBUTED.  Every rank marches the depth-D cells of its Morton range and a share of the refinement
 * passes; prb_get_mesh returns the pieces this rank produced (vertex positions of the pieces back to back, triangles with GLOBAL
 * vertex ids) and prb_get_array("mesh_layout") says where each piece sits in the whole mesh -- writing all ranks' pieces at those
 * offsets gives exactly the single-GPU mesh.  prb_stats.n_vertices / n_triangles / the "passes" array describe the whole mesh. */
/* Triangle mesh of the last prb_extract, in HOST memory owned by the context: vertices are
 * float32 [nv][3] in the normalised unit cube (like the reference's in-core mesh); file
 * coordinates are v*scale + center (prb_stats).  Triangles are int32 [nt][3]. */
int prb_get_mesh(prb_context* ctx, const float** vertices, int64_t* nv, const int32_t** triangles, int64_t* nt);
/* Same data left on the device (no D2H copy). */
int prb_get_mesh_device(prb_context* ctx, const float** d_vertices, int64_t* nv, const int32_t** d_triangles, int64_t* nt);

int prb_get_stats(prb_context* ctx, prb_stats* out);

/* The context's internal CUDA stream (a cudaStream_t), so that a caller can order its own work
 * or record its own events around the stages (bench.py times steps with events on it). */
int prb_get_stream(prb_context* ctx, void** stream);

/* Parity / debug getters: copies a named intermediate array to host memory `dst` (if non-null
 * and cap_bytes is large enough) and returns its size in bytes, or a negative error.  Names:
 *   points normals sorted_idx sorted_key base count key pidx pnum parent didx dnum children
 *   neighs p2n vectorfield divergence x pointvalue iso center_scale cg_iters lap_stencil
 *   vvalue_slots passes mesh_v mesh_t subdivide child0 sg_table df_table
 *   mesh_layout (int64 [pieces][5]: pass, first global vertex, vertices, first global triangle, triangles of every
 *   piece of the mesh this context holds, in the order they sit in prb_get_mesh's arrays)                     */
int64_t prb_get_array(prb_context* ctx, const char* name, void* dst, int64_t cap_bytes);
/* Overwrite an intermediate (teacher forcing in parity tests): vectorfield divergence x iso. */
int prb_set_array(prb_context* ctx, const char* name, const void* src, int64_t bytes);

/* Parity / debug: re-run a single stage ("divergence", "solve", "iso", "extract") on the current
 * intermediates (used with prb_set_array for stage-by-stage comparison against the oracle). */
int prb_run_stage(prb_context* ctx, const char* name);

/* Unit-test hooks of two building blocks (host pointers in and out): the single-pass exclusive scan used by every compaction
 * (replaces the thrust::exclusive_scan / copy_if calls of main.cu:399-4502) and the stable LSD radix sort of (key, index) pairs
 * over the low `key_bits` bits (replaces thrust::sort_by_key, main.cu:598-602): out_idx[i] = input position of the i-th smallest key. */
int prb_debug_scan(prb_context* ctx, const int32_t* in, int64_t n, int32_t* out, int64_t* total);
int prb_debug_sort(prb_context* ctx, const uint64_t* keys, int64_t n, int key_bits, uint64_t* out_keys, int32_t* out_idx);

/* Options: "cg_tol" (default 1e-5, CG_CUDA.cuh:347), "cg_max_iter" (10000, CG_CUDA.cuh:263),
 * "refine" (1 = run the refinement passes, main.cu:3799-4564), "refine_implicit" (1; 0 = materialise
 * the virtual subtrees of every pass: cross-check path), "cg_zigzag" (1 = the CG phases sweep memory
 * in alternating directions for L2 reuse; results do not depend on it), "refine_bound_check" (0; 1 = evaluate
 * every refinement brick and fail if a certified sign is wrong: test mode), "iso_density_weighted" (0; 1 = OPT-IN mode outside reference
 * parity, SURVEY.md 8f-4: the iso value becomes the mean of chi over the samples weighted by 1 / (samples in the sample's ancestor cell
 * at depth D-3), i.e. a mean over the surface instead of over the samples of an unevenly dense scan; prb_get_array("iso_modes") returns
 * [plain mean, weighted mean]; the weighted one is only computed with the option on), "cascadic" (0; 1 = OPT-IN mode outside reference parity, SURVEY.md 8f-3: the depths are solved coarse to fine and the
 * right-hand side of depth d first loses what the coarser solutions explain, b' = b - sum_{e<d} L_{d,e} x_e -- the coupling the
 * reference's independent per-depth systems omit; single GPU; prb_get_array("cascadic_rhs") returns b'), "cg_bulk", "div_mode",
 * "cg_timing", "detail" (INTEGRATION.md), "early_mesh_copy" (0; 1 = the device -> host copy of the main marching-cubes piece into the
 * pinned buffers prb_get_mesh returns starts before the refinement passes and runs under them; same mesh, no net gain measured),
 * "mg_timeout_ms" (2000: how long a rank waits for a peer at a cross-GPU barrier before the run
 * fails with PRB_ERR_CUDA instead of hanging; converted to SM cycles at 2 GHz). */
int prb_set_option(prb_context* ctx, const char* key, double value);

/* ---- Multi-GPU (new: the reference is single-GPU, devID = 0 hard-coded at CG_CUDA.cuh:356).
 * One process and one context per GPU of one NVLink box, 2..8 ranks.  Every rank is given the SAME
 * samples and builds the same octree; the divergence, the CG solve and the iso value are sharded
 * by Morton range (depths with fewer than 65536 nodes stay replicated) and exchange data through
 * a peer-mapped arena: kernels read the other ranks' halo blocks over NVLink directly, phase
 * boundaries are epoch flags in that arena (no host round trip).  Every rank ends with the
 * complete solution and extracts the complete mesh.
 *   prb_mg_init      allocates this rank's arena (cudaMalloc, `arena_bytes`; 8*(nodes+8) bytes
 *                    + 16 KiB are needed per run) and returns its 64-byte CUDA IPC handle;
 *   prb_mg_set_peer  opens the arena of another rank from its handle (exchange the handles with
 *                    any host-side transport, e.g. torch.distributed.all_gather_object);
 *   prb_mg_barrier   box-wide barrier through the arena flags (all ranks must call it);
 *   prb_set_points_sharded / the distributed mesh: see below;
 *   prb_mg_plan      host-only helper: the contiguous split of `count` units over `world` ranks
 *                    used for every sharded range (out[world + 1]). */
int prb_mg_init(prb_context* ctx, int rank, int world, int64_t arena_bytes, void* ipc_handle_out_64_bytes);
/* Multi-GPU input: rank r passes only ITS slice of the cloud -- the samples [plan[r], plan[r+1]) of prb_mg_plan(n_total, world), host
 * or device pointers to the first sample of the slice -- and the slices are gathered over NVLink when the octree is built.  (With
 * prb_set_points every rank passes the whole cloud.)  All ranks must use the same entry point. */
int prb_set_points_sharded(prb_context* ctx, const float* xyz_slice, const float* normals_slice, int64_t n_total);
int prb_mg_set_peer(prb_context* ctx, int peer_rank, const void* ipc_handle_64_bytes);
int prb_mg_barrier(prb_context* ctx);
int prb_mg_plan(int64_t count, int world, int64_t* out);
/* Host-only: the deal of whole refinement passes to ranks used by prb_extract (pass i: count[i] roots at depth depth[i]; largest
 * estimated cost first, to the least loaded rank): owner_out[i] = rank. */
int prb_mg_deal_passes(int depth_max, int n_passes, const int32_t* depth, const int32_t* count, int world, int32_t* owner_out);

/* Host-side B-spline precompute (replaces FunctionData<2,double>::set / setDotTables,
 * FunctionData.inl:112-215, and the table uploads of main.cu:3308-3359).  Needs no GPU: copies
 * the named table for `depth` to `dst` and returns its size in bytes.  Names: gauss (4x4 f32),
 * max_depth_fn (4x4 f32), base_fn (res x 4 x 5 f32), df_table (f32), df_offset (i32, depth+2),
 * stencil ((depth+1) x 27 f32), ff0 ff1 d20 d21 (f64 per depth: same-depth 1-D <F,F> and
 * <F',F'> at centre distance 0 and 1), ff_cross d2_cross (f64) + cross_offset (i32 (depth+1)^2): cross-depth 1-D integrals of the
 * opt-in cascadic mode: entry cross_offset[d][e] + u, u = off_o - 2^(d-e) (off_n - 1), for a depth-d node o and a depth-e node n, e < d). */
int64_t prb_host_tables(int depth, const char* name, void* dst, int64_t cap_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PRB_H_ */
