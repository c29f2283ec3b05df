// Octree construction: normalise -> Morton keys -> radix sort -> unique leaves -> per-depth
// sibling groups -> SoA node slabs -> 27-neighbour tables.
//
// Replaces pipelineBuildNodeArray (main.cu:511-841) and its kernels (main.cu:132-509).  The
// reference builds each level with two 64 MB hash tables, atomics and per-level cudaMalloc /
// memset / D2D concatenation; here every level is a flag + scan + compaction over the sorted
// list of non-empty nodes of the level below (Morton order makes siblings contiguous), and the
// node records are written once, directly into their final slab of a single SoA index space.
// Results follow the reference's INTENDED semantics (the reference's own pidx
// is racy, see tests/golden/ref_sphere100k_d8_report.json "ref_run_to_run").
#include "common.cuh"
#include "scan.cuh"

namespace prb {

int scan_work_ensure(Context& c, size_t tiles) {
    ScanWork& w = c.scanWork;
    if (!c.hScanTotal) PRB_CUDA(cudaMallocHost((void**)&c.hScanTotal, 64));
    if (!w.ticket) {
        PRB_TRY(c.scanTicket.alloc(16, c.stream));
        PRB_CUDA(cudaMemsetAsync(c.scanTicket.p, 0, 16 * sizeof(unsigned), c.stream));
        w.ticket = c.scanTicket.p;
    }
    if (tiles > w.maxTiles) {
        // (stream order: every earlier scan on this stream is complete before the new block is used)
        const size_t want = tiles + tiles / 2 + 1024;
        PRB_TRY(c.scanDesc.alloc(want, c.stream));
        PRB_CUDA(cudaMemsetAsync(c.scanDesc.p, 0, want * sizeof(unsigned long long), c.stream));
        w.desc = c.scanDesc.p;
        w.maxTiles = want;
        w.epoch = 0;                           // fresh zeroed descriptors: status 0 / 1 never match an epoch >= 1
    }
    return PRB_OK;
}

// exclusive scan of an int array (kernels in scan.cuh)
int exclusive_scan(Context& c, const int* in, int* out, i64 n, i64* total_host, int hostSlot) {
    return exclusive_scan_op(c, ScanLoadInt{in}, out, n, total_host, hostSlot);
}

// ------------------------------------------------------------------------------------------
// A0: bounding box (main.cu:530-545)
__global__ void __launch_bounds__(256) k_bbox_partial(const float* __restrict__ xyz, i64 n, float* __restrict__ part) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float v = xyz[3 * i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    __shared__ float sm[6][8];
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) { sm[a][threadIdx.x >> 5] = mn[a]; sm[3 + a][threadIdx.x >> 5] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = sm[threadIdx.x][0];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? fminf(v, sm[threadIdx.x][w]) : fmaxf(v, sm[threadIdx.x][w]);
        part[6 * blockIdx.x + threadIdx.x] = v;
    }
}
__global__ void k_bbox_final(const float* __restrict__ part, int nb, float* __restrict__ out) {
    int a = threadIdx.x;
    if (a >= 6) return;
    float v = part[a];
    for (int b = 1; b < nb; b++) v = a < 3 ? fminf(v, part[6 * b + a]) : fmaxf(v, part[6 * b + a]);
    out[a] = v;
}

// A0 + A1: normalise (main.cu:552-571) and Morton-encode (main.cu:132-164).  Host float
// semantics of the reference are kept: no FMA contraction in the normal length, IEEE division.
// One block per SORT TILE (sort.cu): the digit histogram of the first radix pass is taken on the way.
constexpr int kEncThreads = 256, kEncItems = 16;          // = kSortThreads, kSortItems (sort.cu)
__global__ void __launch_bounds__(kEncThreads) k_normalise_encode_count(const float* __restrict__ xyz, i64 n,
                                                                        float cx, float cy, float cz, float scale, int D, int bits, int nTiles,
                                                                        float* __restrict__ P0, u64* __restrict__ keys, int* __restrict__ idx,
                                                                        int* __restrict__ counts) {
    extern __shared__ int sHist[];
    const int radix = 1 << bits;
    for (int d = threadIdx.x; d < radix; d += kEncThreads) sHist[d] = 0;
    __syncthreads();
    const float ctr[3] = {cx, cy, cz};
    const i64 t0 = (i64)blockIdx.x * (kEncThreads * kEncItems);
    for (int it = 0; it < kEncItems; it++) {
        const i64 i = t0 + it * kEncThreads + threadIdx.x;
        if (i >= n) break;
        float p[3];
#pragma unroll
        for (int a = 0; a < 3; a++) p[a] = __fdiv_rn(__fsub_rn(xyz[3 * i + a], ctr[a]), scale);
        // strict '>' against the running cell centre: a point on a cell boundary goes to the lower cell (Q5)
        float c[3] = {0.5f, 0.5f, 0.5f};
        float w = 0.25f;
        u64 k = 0;
        for (int l = D - 1; l >= 0; --l) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (p[a] > c[a]) { k |= 1ull << (3 * l + 2 - a); c[a] += w; }
                else c[a] -= w;
            }
            w *= 0.5f;
        }
#pragma unroll
        for (int a = 0; a < 3; a++) P0[3 * i + a] = p[a];
        keys[i] = k;
        idx[i] = (int)i;
        atomicAdd(&sHist[(int)(k & (u64)(radix - 1))], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < radix; d += kEncThreads) counts[(size_t)d * nTiles + blockIdx.x] = sHist[d];
}
int sort_tiles(i64 n);
int sort_passes(int keyBits);
int sort_digit_bits(int keyBits);
int radix_sort_gather(Context& c, u64* keys0, int* idx0, u64* keysTmp, int* idxTmp, int* counts, i64 n, int keyBits, const float* P0, const float* N0, float nscale,
                      cudaEvent_t normalsReady, u64* keysOut, int* idxOut, float* P, float* Nr);
// number of non-empty nodes of EVERY depth in one pass over the sorted keys: sample i starts a new node at depth d iff its key differs from
// its predecessor's in the top 3 d bits, i.e. at every depth >= h(i) = the level of the highest differing bit.  hist[h] counts the samples
// by h; U_d = sum_{h <= d} hist[h].  One host round trip for all levels instead of one per level.
__global__ void __launch_bounds__(256) k_level_counts(const u64* __restrict__ key, i64 n, int D, int* __restrict__ hist) {
    __shared__ int sH[kMaxDepth + 2];
    if (threadIdx.x <= kMaxDepth) sH[threadIdx.x] = 0;
    __syncthreads();
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int h = 0;                                      // sample 0 starts a node at every depth
        if (i > 0) {
            const u64 x = key[i] ^ key[i - 1];
            if (x == 0) continue;
            const int top = 63 - __clzll((long long)x);  // highest differing bit; level l owns the bits [3 (D - l), 3 (D - l) + 2]
            h = D - top / 3;
        }
        atomicAdd(&sH[h], 1);
    }
    __syncthreads();
    if (threadIdx.x <= kMaxDepth && sH[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sH[threadIdx.x]);
}
// run heads: first element of every run of equal (key >> shift)
__global__ void __launch_bounds__(256) k_head_flags(const u64* __restrict__ key, i64 n, int shift, int* __restrict__ flag) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        flag[i] = (i == 0) || ((key[i] >> shift) != (key[i - 1] >> shift));
}
__global__ void __launch_bounds__(256) k_compact_leaves(const u64* __restrict__ key, const int* __restrict__ flag, const int* __restrict__ excl, i64 n,
                                                        u64* __restrict__ lkey, int* __restrict__ fp, int nLeaves) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        if (flag[i]) { lkey[excl[i]] = key[i]; fp[excl[i]] = (int)i; }
    if (blockIdx.x == 0 && threadIdx.x == 0) fp[nLeaves] = (int)n;
}
// one level up: parents of the non-empty nodes of depth d (list `lkey`, sorted)
__global__ void __launch_bounds__(256) k_compact_parents(const u64* __restrict__ lkey, const int* __restrict__ fp, const int* __restrict__ fdm1,
                                                         const int* __restrict__ flag, const int* __restrict__ excl, int n, int levelShift, int isDm1,
                                                         u64* __restrict__ pkey, int* __restrict__ pfp, int* __restrict__ pfc, int* __restrict__ pfdm1,
                                                         int* __restrict__ prank, int* __restrict__ slot,
                                                         int nParents, int nPoints, int nDm1) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        int g = excl[r] + flag[r] - 1;
        prank[r] = g;
        slot[r] = 8 * g + (int)((lkey[r] >> levelShift) & 7);
        if (flag[r]) {
            pkey[g] = lkey[r] & ~(7ull << levelShift);
            pfp[g] = fp[r];
            pfc[g] = r;
            pfdm1[g] = isDm1 ? g : fdm1[r];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { pfp[nParents] = nPoints; pfc[nParents] = n; pfdm1[nParents] = nDm1; }
}

struct NodeArrays {
    u64* key;
    int *parent, *child0, *pidx, *pnum, *didx, *dnum;
};
// one thread per sibling group of depth d: writes the 8 node records (main.cu:226-276 depth D,
// 301-359 upper levels, 433-490 prefix pidx/didx + keys of empty siblings)
__global__ void __launch_bounds__(128) k_fill_groups(NodeArrays A, int d, int D, int baseD, int baseParent, int baseChild,
                                                     int nGroups, const u64* __restrict__ pkey, const int* __restrict__ pslot,
                                                     const int* __restrict__ fc, const u64* __restrict__ ckey, const int* __restrict__ cfp,
                                                     const int* __restrict__ cfdm1) {
    int shift = 3 * (D - d);
    // the 8 records of a group are computed by one thread and then transposed through shared
    // memory, so that every store instruction of a warp writes 32 consecutive nodes
    __shared__ int sT[4][32 * 9];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int* tile = sT[wp];
    for (int gw = (blockIdx.x * blockDim.x + threadIdx.x) - lane; gw < nGroups; gw += gridDim.x * blockDim.x) {
        const int g = gw + lane;
        const bool act = g < nGroups;
        int pn[8], pi[8], dn[8], di[8], ch[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { pn[k] = 0; pi[k] = 0; dn[k] = (d == D) ? 1 : 0; di[k] = 0; ch[k] = -1; }
        int firstP = 0, firstD = 0;
        u64 kbase = 0;
        int par = 0;
        if (act) {
            int r0 = fc[g], r1 = fc[g + 1];
            for (int r = r0; r < r1; r++) {
                int cc = (int)((ckey[r] >> shift) & 7);
                int p0 = cfp[r], p1 = cfp[r + 1];
                int dd0 = 0, dd1 = 0;
                if (d < D) { dd0 = 8 * cfdm1[r]; dd1 = 8 * cfdm1[r + 1]; }
                if (r == r0) { firstP = p0; firstD = dd0; }
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k == cc) { pn[k] = p1 - p0; pi[k] = p0; dn[k] = (d == D) ? 1 : dd1 - dd0; di[k] = dd0; ch[k] = (d < D) ? baseChild + 8 * r : -1; }
            }
            kbase = pkey[g];
            par = baseParent + pslot[g];
        }
        int nowP = firstP, nowD = firstD;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            pi[k] = nowP; nowP += pn[k];
            if (d == D) di[k] = 8 * g + k; else { di[k] = nowD; nowD += dn[k]; }
        }
        const int i0w = baseD + 8 * gw;                     // first node of the warp's 32 groups
        const int nw = 8 * min(32, nGroups - gw);            // nodes of this warp step
        auto flush = [&](const int (&v)[8], int* __restrict__ dst) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) tile[lane * 9 + k] = v[k];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int idx = r * 32 + lane;
                if (idx < nw) dst[i0w + idx] = tile[(idx >> 3) * 9 + (idx & 7)];
            }
        };
        flush(ch, A.child0);
        flush(pn, A.pnum);
        flush(pi, A.pidx);
        flush(dn, A.dnum);
        flush(di, A.didx);
        // parent and key: the group's value is shared by its 8 nodes (key: plus the child code)
        __syncwarp();
        tile[lane * 9] = par;
        tile[lane * 9 + 1] = (int)(unsigned)(kbase & 0xffffffffu);
        tile[lane * 9 + 2] = (int)(unsigned)(kbase >> 32);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int idx = r * 32 + lane;
            if (idx < nw) {
                const int gl = idx >> 3, k = idx & 7;
                A.parent[i0w + idx] = tile[gl * 9];
                const u64 kb = (u64)(unsigned)tile[gl * 9 + 1] | ((u64)(unsigned)tile[gl * 9 + 2] << 32);
                A.key[i0w + idx] = kb | ((u64)k << shift);
            }
        }
    }
}
__global__ void k_fill_root(NodeArrays A, int nPoints, int nSlotsD, int hasChildren) {
    A.key[0] = 0; A.parent[0] = -1; A.child0[0] = hasChildren ? 1 : -1; A.pidx[0] = 0; A.pnum[0] = nPoints; A.didx[0] = 0; A.dnum[0] = nSlotsD;
}
__global__ void __launch_bounds__(256) k_node_offsets(const u64* __restrict__ key, int base, int count, int d, int D, ushort4* __restrict__ offs) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < count; l += gridDim.x * blockDim.x) {
        u64 k = key[base + l];
        int ox = 0, oy = 0, oz = 0;
        for (int lv = 1; lv <= d; lv++) {
            int c = (int)((k >> (3 * (D - lv))) & 7);
            ox |= ((c >> 2) & 1) << (d - lv);
            oy |= ((c >> 1) & 1) << (d - lv);
            oz |= (c & 1) << (d - lv);
        }
        offs[base + l] = make_ushort4((unsigned short)ox, (unsigned short)oy, (unsigned short)oz, (unsigned short)d);
    }
}
// neighbours of depth d from the parents' (main.cu:492-509).  One WARP per sibling group: the 8
// siblings share the parent, so its 27 neighbours and their first-child indices are fetched
// once (one lane each) and the group's 216 table entries -- one contiguous 864-byte record --
// are produced from them with shuffles through the closed-form LUT.
__global__ void __launch_bounds__(256) k_neighbours(const int* __restrict__ parent, const int* __restrict__ child0, int* __restrict__ neighs, int base, int count) {
    __shared__ unsigned char sLut[216];          // (c, j) -> pj | cc << 5
    for (int t = threadIdx.x; t < 216; t += blockDim.x) {
        int pj, cc;
        lut_parent_child(t / 27, t % 27, pj, cc);
        sLut[t] = (unsigned char)(pj | (cc << 5));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, nGroups = count >> 3;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < nGroups; g += (gridDim.x * blockDim.x) >> 5) {
        const int i0 = base + 8 * g;
        const int P = parent[i0];
        int c0 = -1;
        if (lane < 27) {
            const int np = neighs[27 * (i64)P + lane];
            if (np >= 0) c0 = child0[np];
        }
        int* out = neighs + 27 * (i64)i0;
#pragma unroll
        for (int r = 0; r < 7; r++) {
            const int t = r * 32 + lane;
            const int e = sLut[t < 216 ? t : 0];
            const int v = __shfl_sync(0xffffffffu, c0, e & 31);
            if (t < 216) out[t] = v >= 0 ? v + (e >> 5) : -1;
        }
    }
}
__global__ void k_root_neighbours(int* __restrict__ neighs) {
    int j = threadIdx.x;
    if (j < 27) neighs[j] = (j == 13) ? 0 : -1;
}
// Super-group table of the stencil SpMV (solver.cu).  Super-group 1+G' = the (up to 8) sibling
// groups whose parents P_k are the 8 nodes of group G' (children of one node Q); entry u =
// ux*16+uy*4+uz (each in 0..3) is the first child of the node at offset 2*off(Q) + u - 1 in the
// P-level grid (-1: absent or childless), i.e. the 4x4x4 cube of 8-row blocks that holds every
// neighbour of every row under Q.  The interior entries (u in {1,2}^3) are the row bases of the
// super-group's own groups.  Super-group 0 is depth 1 (the root's children).
__global__ void __launch_bounds__(256) k_sg_table(const int* __restrict__ parent, const int* __restrict__ child0, const int* __restrict__ neighs,
                                                  int nSg, int* __restrict__ sgTab) {
    i64 total = (i64)nSg * 64;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        int sg = (int)(t >> 6), u = (int)(t & 63);
        int ux = u >> 4, uy = (u >> 2) & 3, uz = u & 3;
        int out = -1;
        if (sg == 0) {
            if (u == 21) out = child0[0];
        } else {
            int Q = parent[1 + 8 * (sg - 1)];
            int j = 9 * ((ux + 1) >> 1) + 3 * ((uy + 1) >> 1) + ((uz + 1) >> 1);
            int np = neighs[27 * (i64)Q + j];
            if (np >= 0) {
                int b = child0[np];
                if (b >= 0) out = child0[b + ((((ux + 1) & 1) << 2) | (((uy + 1) & 1) << 1) | ((uz + 1) & 1))];
            }
        }
        sgTab[t] = out;
    }
}
// Multi-GPU shard plan of the sharded depths.  Rank r's share of depth d starts at the super-group that holds the row at fraction
// r / world of the depth's COST (rows plus a fixed part per super-group, see below: super-groups hold anything from 8 to 64 rows) and its first node is the first existing sibling group at or after that
// super-group (interior table entries in child order).
__global__ void k_shard_plan(const int* __restrict__ sgTab, const int* __restrict__ parent, const int* __restrict__ sgStart /* [D+2] */, const int* __restrict__ base /* [D+2] */,
                             int D, int world, int shardFrom, int* __restrict__ sgLo /* [(D+2)][kMaxRanks+1] */, int* __restrict__ rowLo) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (D + 2) * (kMaxRanks + 1)) return;
    int d = t / (kMaxRanks + 1), r = t % (kMaxRanks + 1);
    if (d < shardFrom || d > D || r > world) return;          // (the host keeps its own values for the replicated depths)
    const int cntD = base[d + 1] - base[d];
    int sg;
    if (r == 0) sg = sgStart[d];
    else if (r == world) sg = sgStart[d + 1];
    else {
        // cost of a share = its rows + 0.3 x (average rows per super-group) x its super-groups: the SpMV pays a fixed part for every
        // super-group it stages, however few rows it holds (measured on 8 GPUs: equal rows left the SpMV times 1.48 .. 2.20 ms apart).
        // The cost of the sibling groups before g is monotone in g: binary search for the target fraction
        const int nG = cntD >> 3, nSg = sgStart[d + 1] - sgStart[d];
        const double cPerSg = 0.3 * (double)cntD / (double)(nSg > 0 ? nSg : 1);
        const double target = ((double)cntD + cPerSg * nSg) * (double)r / (double)world;
        int lo = 0, hi = nG;                                               // smallest g with cost(g) >= target
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int sgm = 1 + (parent[base[d] + 8 * mid] - 1) / 8;
            const double cost = 8.0 * mid + cPerSg * (double)(sgm - sgStart[d]);
            if (cost < target) lo = mid + 1; else hi = mid;
        }
        const int g = lo < nG ? lo : nG - 1;
        sg = 1 + (parent[base[d] + 8 * g] - 1) / 8;
    }
    sgLo[t] = sg;
    int out = base[d + 1];
    if (sg < sgStart[d + 1]) {
        const int idx[8] = {21, 22, 25, 26, 37, 38, 41, 42};
        for (int k = 0; k < 8; k++) {
            int v = sgTab[64 * (i64)sg + idx[k]];
            if (v >= 0) { out = v; break; }
        }
    }
    rowLo[t] = out;
}
__global__ void __launch_bounds__(256) k_point_to_leaf(const int* __restrict__ flag, const int* __restrict__ excl, const int* __restrict__ slotD, i64 n, int* __restrict__ p2n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        p2n[i] = slotD[excl[i] + flag[i] - 1];
}

int stage_octree(Context& c) {
    const int D = c.D;
    const i64 N = c.N;
    cudaStream_t st = c.stream;
    if (N <= 0 || N > 0x7fffffff / 4) { set_error("point count out of range"); return PRB_ERR_ARG; }
    if (c.rawSharded) {
        // multi-GPU: every rank uploaded its slice into its arena; gather the others (24 bytes per sample over NVLink)
        long long lo[kMaxRanks + 1];
        for (int r = 0; r <= c.mg.world; r++) lo[r] = (N * r) / c.mg.world;
        for (int q = 0; q < c.mg.world; q++)
            if (!c.mg.peer[q]) { set_error("multi-GPU: peer arenas not exchanged (prb_mg_set_peer)"); return PRB_ERR_STATE; }
        PRB_TRY(mg_barrier(c));
        {
            const void* src[2 * kMaxRanks];
            void* dst[2 * kMaxRanks];
            size_t bytes[2 * kMaxRanks];
            int n = 0;
            for (int qi = 1; qi < c.mg.world; qi++) {
                const int q = (c.mg.rank + qi) % c.mg.world;
                const size_t a = 12 * (size_t)lo[q], b = 12 * (size_t)lo[q + 1];
                src[n] = c.mg.peer[q] + c.mgRawPOff + a; dst[n] = (char*)c.rawPp + a; bytes[n] = b - a; n++;
                src[n] = c.mg.peer[q] + c.mgRawNOff + a; dst[n] = (char*)c.rawNp + a; bytes[n] = b - a; n++;
            }
            PRB_TRY(mg_pull(c, n, src, dst, bytes));
        }
        PRB_TRY(mg_barrier(c));
        c.rawSharded = false;
    }
    mark(c, "octree:gathered");
    // ---- A0 bounding box -> scale / centre (host float arithmetic of main.cu:539-545)
    {
        int nb = grid_for(c, N, 256, 4);
        DBuf<float> part, box;
        PRB_TRY(part.alloc((size_t)nb * 6, st));
        PRB_TRY(box.alloc(6, st));
        PRB_LAUNCH(c, k_bbox_partial, nb, 256, 0, c.rawPp, N, part.p);
        PRB_LAUNCH(c, k_bbox_final, 1, 32, 0, part.p, nb, box.p);
        float h[6];
        PRB_CUDA(cudaMemcpyAsync(h, box.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        float scale = 1;
        for (int a = 0; a < 3; a++) {
            if (!a || scale < h[3 + a] - h[a]) scale = float(h[3 + a] - h[a]);
            c.center[a] = float(h[3 + a] + h[a]) / 2;
        }
        scale *= 1.25f;
        for (int a = 0; a < 3; a++) c.center[a] -= scale / 2;
        c.scale = scale;
        part.release();
        box.release();
    }
    // ---- A1/A2 keys + sort (thrust::sort_by_key x2 on 64-bit codes in the reference, main.cu:598-602;
    //      here one stable LSD sort over the 3D key bits with the sample index as payload)
    mark(c, "octree:bbox");
    DBuf<float> P0;
    DBuf<u64> keys0;
    DBuf<int> idx0;
    PRB_TRY(P0.alloc(3 * (size_t)N, st));
    PRB_TRY(keys0.alloc((size_t)N, st));
    PRB_TRY(idx0.alloc((size_t)N, st));
    PRB_TRY(c.sortedKey.alloc((size_t)N, st));
    PRB_TRY(c.sortedIdx.alloc((size_t)N, st));
    PRB_TRY(c.P.alloc(3 * (size_t)N, st));
    PRB_TRY(c.Nr.alloc(3 * (size_t)N, st));
    {
        // A2: stable LSD radix sort over the 3 D key bits with the sample index as payload (sort.cu); the first digit histogram comes
        // out of the key generation, the sample gather rides on the last scatter pass
        const int bits = sort_digit_bits(3 * D), nTiles = sort_tiles(N);
        DBuf<u64> keysTmp;
        DBuf<int> idxTmp, counts;
        const bool needTmp = sort_passes(3 * D) > 1;
        PRB_TRY(keysTmp.alloc(needTmp ? (size_t)N : 0, st));
        PRB_TRY(idxTmp.alloc(needTmp ? (size_t)N : 0, st));
        PRB_TRY(counts.alloc(((size_t)1 << bits) * (size_t)nTiles, st));
        PRB_LAUNCH(c, k_normalise_encode_count, nTiles, kEncThreads, sizeof(int) << bits, c.rawPp, N, c.center[0], c.center[1], c.center[2], c.scale, D, bits, nTiles,
                   P0.p, keys0.p, idx0.p, counts.p);
        // the normals are first needed by the gather of the last pass (scaled there): their upload may still be running on the copy stream
        PRB_TRY(radix_sort_gather(c, keys0.p, idx0.p, keysTmp.p, idxTmp.p, counts.p, N, 3 * D, P0.p, c.rawNp, (float)(2 << D), c.normalsPending ? c.evNormals : nullptr,
                                  c.sortedKey.p, c.sortedIdx.p, c.P.p, c.Nr.p));
        c.normalsPending = false;
    }
    P0.release(); keys0.release(); idx0.release();
    mark(c, "octree:sorted");
    // ---- A3 unique leaves
    DBuf<int> flagN, exclN;
    PRB_TRY(flagN.alloc((size_t)N, st));
    PRB_TRY(exclN.alloc((size_t)N, st));
    PRB_LAUNCH(c, k_head_flags, grid_for(c, N, 256), 256, 0, c.sortedKey.p, N, 0, flagN.p);
    PRB_TRY(exclusive_scan(c, flagN.p, exclN.p, N, nullptr));
    // per-level lists of non-empty nodes; their sizes U_d for all depths from one kernel + one host round trip
    std::vector<DBuf<u64>> lkey(D + 1);
    std::vector<DBuf<int>> fp(D + 1), fc(D + 1), fdm1(D + 1), prank(D + 1), slot(D + 1);
    std::vector<int> U(D + 1, 0);
    {
        DBuf<int> hist;
        PRB_TRY(hist.alloc(kMaxDepth + 2, st));
        PRB_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(int) * (kMaxDepth + 2), st));
        PRB_LAUNCH(c, k_level_counts, grid_for(c, N, 256, 4), 256, 0, c.sortedKey.p, N, D, hist.p);
        int h[kMaxDepth + 2];
        PRB_CUDA(cudaMemcpyAsync(h, hist.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        int acc = 0;
        for (int d = 0; d <= D; d++) { acc += h[d]; U[d] = acc; }
    }
    PRB_TRY(lkey[D].alloc((size_t)U[D], st));
    PRB_TRY(fp[D].alloc((size_t)U[D] + 1, st));
    PRB_LAUNCH(c, k_compact_leaves, grid_for(c, N, 256), 256, 0, c.sortedKey.p, flagN.p, exclN.p, N, lkey[D].p, fp[D].p, U[D]);
    // ---- A4/A5 levels D -> 0
    for (int d = D; d >= 1; --d) {
        int n = U[d];
        DBuf<int> fl, ex;
        PRB_TRY(fl.alloc((size_t)n, st));
        PRB_TRY(ex.alloc((size_t)n, st));
        PRB_LAUNCH(c, k_head_flags, grid_for(c, n, 256), 256, 0, lkey[d].p, (i64)n, 3 * (D - d + 1), fl.p);
        const i64 np = U[d - 1];
        PRB_TRY(exclusive_scan(c, fl.p, ex.p, n, nullptr));
        PRB_TRY(lkey[d - 1].alloc((size_t)np, st));
        PRB_TRY(fp[d - 1].alloc((size_t)np + 1, st));
        PRB_TRY(fc[d - 1].alloc((size_t)np + 1, st));
        PRB_TRY(fdm1[d - 1].alloc((size_t)np + 1, st));
        PRB_TRY(prank[d].alloc((size_t)n, st));
        PRB_TRY(slot[d].alloc((size_t)n, st));
        int nDm1 = (D >= 1) ? U[D - 1] : 0;   // U[D-1] is known from the first iteration on
        if (d == D) nDm1 = (int)np;
        PRB_LAUNCH(c, k_compact_parents, grid_for(c, n, 256), 256, 0, lkey[d].p, fp[d].p, fdm1[d].p, fl.p, ex.p, n, 3 * (D - d), (d - 1 == D - 1) ? 1 : 0,
                   lkey[d - 1].p, fp[d - 1].p, fc[d - 1].p, fdm1[d - 1].p, prank[d].p, slot[d].p, (int)np, (int)N, nDm1);
        fl.release();
        ex.release();
    }
    PRB_TRY(slot[0].alloc(1, st));
    PRB_CUDA(cudaMemsetAsync(slot[0].p, 0, sizeof(int), st));
    mark(c, "octree:levels");
    // ---- node slabs
    c.cnt[0] = 1;
    for (int d = 1; d <= D; d++) c.cnt[d] = 8 * U[d - 1];
    c.base[0] = 0;
    for (int d = 1; d <= D + 1; d++) c.base[d] = c.base[d - 1] + c.cnt[d - 1];
    c.M = c.base[D + 1];
    const int M = c.M;
    PRB_TRY(c.key.alloc((size_t)M, st));
    PRB_TRY(c.parent.alloc((size_t)M, st));
    PRB_TRY(c.child0.alloc((size_t)M, st));
    PRB_TRY(c.pidx.alloc((size_t)M, st));
    PRB_TRY(c.pnum.alloc((size_t)M, st));
    PRB_TRY(c.didx.alloc((size_t)M, st));
    PRB_TRY(c.dnum.alloc((size_t)M, st));
    PRB_TRY(c.neighs.alloc(27 * (size_t)M, st));
    PRB_TRY(c.offs.alloc((size_t)M, st));
    PRB_TRY(c.dBase.alloc(kMaxDepth + 2, st));
    PRB_CUDA(cudaMemcpyAsync(c.dBase.p, c.base, sizeof(int) * (kMaxDepth + 2), cudaMemcpyHostToDevice, st));
    NodeArrays A{c.key.p, c.parent.p, c.child0.p, c.pidx.p, c.pnum.p, c.didx.p, c.dnum.p};
    PRB_LAUNCH(c, k_fill_root, 1, 1, 0, A, (int)N, c.cnt[D], D >= 1 ? 1 : 0);
    for (int d = 1; d <= D; d++) {
        int ng = U[d - 1];
        PRB_LAUNCH(c, k_fill_groups, grid_for(c, ng, 128), 128, 0, A, d, D, c.base[d], c.base[d - 1], d < D ? c.base[d + 1] : 0, ng,
                   lkey[d - 1].p, slot[d - 1].p, fc[d - 1].p, lkey[d].p, fp[d].p, fdm1[d].p);
    }
    PRB_TRY(c.p2n.alloc((size_t)N, st));
    PRB_LAUNCH(c, k_point_to_leaf, grid_for(c, N, 256), 256, 0, flagN.p, exclN.p, slot[D].p, N, c.p2n.p);
    for (int d = 0; d <= D; d++) PRB_LAUNCH(c, k_node_offsets, grid_for(c, c.cnt[d], 256), 256, 0, c.key.p, c.base[d], c.cnt[d], d, D, c.offs.p);
    mark(c, "octree:nodes");
    // ---- A6 neighbours, level by level
    PRB_LAUNCH(c, k_root_neighbours, 1, 32, 0, c.neighs.p);
    for (int d = 1; d <= D; d++)
        PRB_LAUNCH(c, k_neighbours, grid_for(c, (i64)c.cnt[d] * 4, 256), 256, 0, c.parent.p, c.child0.p, c.neighs.p, c.base[d], c.cnt[d]);
    mark(c, "octree:neighbours");
    // super-groups: depth 1, then one per sibling group of depths 1..D-1
    c.nSg = 1 + (c.base[D] - 1) / 8;
    PRB_TRY(c.sgTab.alloc(64 * (size_t)c.nSg, st));
    PRB_LAUNCH(c, k_sg_table, grid_for(c, (i64)c.nSg * 64, 256), 256, 0, c.parent.p, c.child0.p, c.neighs.p, c.nSg, c.sgTab.p);
    PRB_TRY(build_cg_table(c));
    mark(c, "octree:sgtable");
    // ---- multi-GPU shard plan: depths with at least minShardRows nodes (and every deeper one) are
    // split into contiguous super-group ranges, the shallower ones are replicated on every rank
    {
        const int W = c.mg.world;
        c.shardFrom = D + 1;
        if (c.mg.active())
            for (int d = D; d >= 2 && c.cnt[d] >= c.mg.minShardRows; --d) c.shardFrom = d;
        int sgStart[kMaxDepth + 2];
        sgStart[0] = sgStart[1] = 0;
        for (int d = 2; d <= D + 1; d++) sgStart[d] = 1 + (c.base[d - 1] - 1) / 8;
        for (int d = 0; d <= D + 1; d++)
            for (int r = 0; r <= kMaxRanks; r++) { c.sgLo[d][r] = 0; c.rowLo[d][r] = 0; }
        for (int d = 1; d <= D; d++) {
            const i64 n = sgStart[d + 1] - sgStart[d];
            for (int r = 0; r <= W; r++) {
                c.sgLo[d][r] = d >= c.shardFrom ? sgStart[d] + (int)((n * r) / W) : (r == 0 ? sgStart[d] : sgStart[d + 1]);
                c.rowLo[d][r] = r == 0 ? c.base[d] : c.base[d + 1];
            }
        }
        if (c.shardFrom <= D) {
            DBuf<int> dSgLo, dRowLo, dSgStart;
            const int nt = (D + 2) * (kMaxRanks + 1);
            PRB_TRY(dSgLo.alloc(nt, st)); PRB_TRY(dRowLo.alloc(nt, st)); PRB_TRY(dSgStart.alloc(kMaxDepth + 2, st));
            int hStart[kMaxDepth + 2] = {0};
            for (int d = 0; d <= D + 1; d++) hStart[d] = sgStart[d];
            PRB_CUDA(cudaMemcpyAsync(dSgStart.p, hStart, sizeof(hStart), cudaMemcpyHostToDevice, st));
            PRB_LAUNCH(c, k_shard_plan, 1, 256, 0, c.sgTab.p, c.parent.p, dSgStart.p, c.dBase.p, D, W, c.shardFrom, dSgLo.p, dRowLo.p);
            int hRow[kMaxDepth + 2][kMaxRanks + 1], hSg[kMaxDepth + 2][kMaxRanks + 1];
            PRB_CUDA(cudaMemcpyAsync(&hRow[0][0], dRowLo.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, st));
            PRB_CUDA(cudaMemcpyAsync(&hSg[0][0], dSgLo.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, st));
            PRB_CUDA(cudaStreamSynchronize(st));
            for (int d = c.shardFrom; d <= D; d++)
                for (int r = 0; r <= W; r++) { c.rowLo[d][r] = hRow[d][r]; c.sgLo[d][r] = hSg[d][r]; }
        }
    }
    flagN.release(); exclN.release();
    for (int d = 0; d <= D; d++) { lkey[d].release(); fp[d].release(); fc[d].release(); fdm1[d].release(); prank[d].release(); slot[d].release(); }
    PRB_CUDA(cudaGetLastError());
    return PRB_OK;
}

}  // namespace prb
