// A2: radix sort of the (Morton key, sample index) pairs + gather of the samples into sorted order.
//
// Replaces the two thrust::sort_by_key calls over 64-bit (key << 32 | index) codes of the reference
// (main.cu:598-602: 8 digit passes each) and the library sort of round 1.  Only the 3 D key bits are
// sorted, by a STABLE least-significant-digit sort with the sample index as payload -- the same total
// order as sorting (key, index) pairs, since the input is in index order.  passes = ceil(3D / 10) digit
// passes of <= 10 bits (3 at depth <= 10, 4 at depth 11 / 12), each
//   count   per-tile digit histogram (shared-memory atomics; pass 0 is fused into the key generation,
//           octree.cu k_normalise_encode_count), written digit-major: counts[digit][tile]
//   scan    ONE exclusive scan over that array (scan.cuh, single pass): digit-major order is the output
//           order, so the scanned value is where the tile's items of that digit go
//   scatter stable ranks inside the tile -- per 32-item row __match_any_sync on the digit, running
//           per-warp digit counters in shared memory, then an 8-step scan over the warps -- and the
//           scatter; the LAST pass also gathers the sample positions / normals (24 + 24 bytes per
//           sample moved once, no separate gather kernel) and scales the raw normals on the way
//           (main.cu:553-571 semantics: unit length times 2^(D+1); no FMA contraction, IEEE division).
// No cross-tile spinning: a tile's base comes from the scan, so the kernels have no forward-progress
// requirements.  Algorithmic bytes per sample (k = 8): passes x (8 r count + 12 r + 12 w
// scatter) + 48 r + 48 w gather.
#include "common.cuh"
#include "scan.cuh"

namespace prb {

constexpr int kSortThreads = 256, kSortItems = 16, kSortTile = kSortThreads * kSortItems, kSortWarpItems = 32 * kSortItems;
constexpr int kSortMaxBits = 10;

__global__ void __launch_bounds__(kSortThreads) k_sort_count(const u64* __restrict__ keys, i64 n, int shift, int bits, int nTiles, int* __restrict__ counts) {
    extern __shared__ int sHist[];
    const int radix = 1 << bits;
    for (int d = threadIdx.x; d < radix; d += kSortThreads) sHist[d] = 0;
    __syncthreads();
    const i64 t0 = (i64)blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const i64 i = t0 + k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&sHist[(int)((keys[i] >> shift) & (u64)(radix - 1))], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < radix; d += kSortThreads) counts[(size_t)d * nTiles + blockIdx.x] = sHist[d];
}

template <bool GATHER>
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const u64* __restrict__ keysIn, const int* __restrict__ idxIn, i64 n, int shift, int bits, int nTiles,
                                                               const int* __restrict__ offsets, u64* __restrict__ keysOut, int* __restrict__ idxOut,
                                                               const float* __restrict__ P0, const float* __restrict__ N0 /* raw normals */, float nscale,
                                                               float* __restrict__ P, float* __restrict__ Nr) {
    extern __shared__ int sCnt[];                           // [8 warps][radix]
    const int radix = 1 << bits, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < 8 * radix; d += kSortThreads) sCnt[d] = 0;
    __syncthreads();
    int* cnt = sCnt + wp * radix;
    // warp-striped: warp wp owns the items [t0 + wp * 512, + 512), row k = 32 consecutive items -> (warp, row, lane) is index order
    const i64 w0 = (i64)blockIdx.x * kSortTile + (i64)wp * kSortWarpItems;
    u64 key[kSortItems];
    int val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const i64 i = w0 + k * 32 + lane;
        const bool ok = i < n;
        key[k] = ok ? keysIn[i] : 0;
        val[k] = ok ? idxIn[i] : 0;
        const unsigned d = ok ? (unsigned)((key[k] >> shift) & (u64)(radix - 1)) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int r = __popc(peers & ((1u << lane) - 1u));
        int prev = 0;
        if (ok) prev = cnt[d];
        __syncwarp();
        if (ok && r == 0) cnt[d] = prev + __popc(peers);
        __syncwarp();
        rank[k] = prev + r;
    }
    __syncthreads();
    // per digit: where this tile's items go (scan of the digit-major counts), then the exclusive prefix over the warps
    for (int d = threadIdx.x; d < radix; d += kSortThreads) {
        int run = offsets[(size_t)d * nTiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < 8; w++) { const int t = sCnt[w * radix + d]; sCnt[w * radix + d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const i64 i = w0 + k * 32 + lane;
        if (i >= n) continue;
        const int d = (int)((key[k] >> shift) & (u64)(radix - 1));
        const i64 pos = (i64)cnt[d] + rank[k];
        keysOut[pos] = key[k];
        idxOut[pos] = val[k];
        if (GATHER) {
            const i64 s = val[k];
            const float q[3] = {N0[3 * s], N0[3 * s + 1], N0[3 * s + 2]};
            const float sq = __fadd_rn(__fadd_rn(__fmul_rn(q[0], q[0]), __fmul_rn(q[1], q[1])), __fmul_rn(q[2], q[2]));
            float len = (float)sqrt((double)sq);
            if (len > 1e-6f) len = __fdiv_rn(1.0f, len);
            len = __fmul_rn(len, nscale);
#pragma unroll
            for (int a = 0; a < 3; a++) { P[3 * pos + a] = P0[3 * s + a]; Nr[3 * pos + a] = __fmul_rn(q[a], len); }
        }
    }
}

int sort_tiles(i64 n) { return div_up(n, kSortTile); }
int sort_passes(int keyBits) { return keyBits <= 0 ? 1 : (keyBits + kSortMaxBits - 1) / kSortMaxBits; }
int sort_digit_bits(int keyBits) { const int p = sort_passes(keyBits); return keyBits <= 0 ? 1 : (keyBits + p - 1) / p; }

// keys0 / idx0: the unsorted pairs (clobbered); counts: [2^bits * tiles] ints, holding the pass-0 histogram on entry.
int radix_sort_gather(Context& c, u64* keys0, int* idx0, u64* keysTmp, int* idxTmp, int* counts, i64 n, int keyBits, const float* P0, const float* N0, float nscale,
                      cudaEvent_t normalsReady /* or null */, u64* keysOut, int* idxOut, float* P, float* Nr) {
    const int passes = sort_passes(keyBits), bits = sort_digit_bits(keyBits), radix = 1 << bits, nTiles = sort_tiles(n);
    const size_t smemCnt = sizeof(int) * 8 * (size_t)radix;
    if (smemCnt > 48 * 1024) { set_error("sort: digit too wide"); return PRB_ERR_ARG; }
    u64* srcK = keys0;
    int* srcI = idx0;
    for (int p = 0; p < passes; p++) {
        const int shift = p * bits;
        const bool last = p == passes - 1;
        if (p > 0) PRB_LAUNCH(c, k_sort_count, nTiles, kSortThreads, sizeof(int) * radix, srcK, n, shift, bits, nTiles, counts);
        PRB_TRY(exclusive_scan(c, counts, counts, (i64)radix * nTiles, nullptr));
        u64* dstK = last ? keysOut : (srcK == keys0 ? keysTmp : keys0);
        int* dstI = last ? idxOut : (srcI == idx0 ? idxTmp : idx0);
        if (last) {
            if (normalsReady) PRB_CUDA(cudaStreamWaitEvent(c.stream, normalsReady, 0));
            PRB_LAUNCH(c, k_sort_scatter<true>, nTiles, kSortThreads, smemCnt, srcK, srcI, n, shift, bits, nTiles, counts, dstK, dstI, P0, N0, nscale, P, Nr);
        } else
            PRB_LAUNCH(c, k_sort_scatter<false>, nTiles, kSortThreads, smemCnt, srcK, srcI, n, shift, bits, nTiles, counts, dstK, dstI, nullptr, nullptr, 0.f, nullptr, nullptr);
        srcK = dstK;
        srcI = dstI;
    }
    return PRB_OK;
}

}  // namespace prb

// ---- unit-test hooks (include/prb.h "debug"): the scan and the sort on caller-provided host data
using namespace prb;
extern "C" {
int prb_debug_scan(prb_context* h, const int32_t* in, int64_t n, int32_t* out, int64_t* total) {
    if (!h || (n > 0 && (!in || !out))) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    DBuf<int> a, b;
    PRB_TRY(a.alloc((size_t)n, c.stream));
    PRB_TRY(b.alloc((size_t)n, c.stream));
    if (n) PRB_CUDA(cudaMemcpyAsync(a.p, in, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
    i64 t = 0;
    PRB_TRY(exclusive_scan(c, a.p, b.p, n, &t));
    if (n) PRB_CUDA(cudaMemcpyAsync(out, b.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    PRB_CUDA(cudaStreamSynchronize(c.stream));
    if (total) *total = t;
    return PRB_OK;
}
int prb_debug_sort(prb_context* h, const uint64_t* keys, int64_t n, int key_bits, uint64_t* out_keys, int32_t* out_idx) {
    if (!h || n <= 0 || !keys || !out_keys || !out_idx || key_bits < 1 || key_bits > 40) return PRB_ERR_ARG;
    Context& c = h->c;
    PRB_DEVICE(c);
    cudaStream_t st = c.stream;
    const int bits = sort_digit_bits(key_bits), nTiles = sort_tiles(n);
    DBuf<u64> k0, k1, ko;
    DBuf<int> i0, i1, io, counts;
    DBuf<float> f0, f1;
    PRB_TRY(k0.alloc((size_t)n, st)); PRB_TRY(k1.alloc((size_t)n, st)); PRB_TRY(ko.alloc((size_t)n, st));
    PRB_TRY(i0.alloc((size_t)n, st)); PRB_TRY(i1.alloc((size_t)n, st)); PRB_TRY(io.alloc((size_t)n, st));
    PRB_TRY(f0.alloc(3 * (size_t)n, st)); PRB_TRY(f1.alloc(6 * (size_t)n, st));
    PRB_TRY(counts.alloc(((size_t)1 << bits) * (size_t)nTiles, st));
    PRB_CUDA(cudaMemcpyAsync(k0.p, keys, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    std::vector<int> iota((size_t)n);
    for (int64_t i = 0; i < n; i++) iota[(size_t)i] = (int)i;
    PRB_CUDA(cudaMemcpyAsync(i0.p, iota.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemsetAsync(f0.p, 0, 12 * (size_t)n, st));
    PRB_LAUNCH(c, k_sort_count, nTiles, kSortThreads, sizeof(int) << bits, k0.p, (i64)n, 0, bits, nTiles, counts.p);
    PRB_TRY(radix_sort_gather(c, k0.p, i0.p, k1.p, i1.p, counts.p, n, key_bits, f0.p, f0.p, 1.f, nullptr, ko.p, io.p, f1.p, f1.p + 3 * (size_t)n));
    PRB_CUDA(cudaMemcpyAsync(out_keys, ko.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaMemcpyAsync(out_idx, io.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    return PRB_OK;
}
}
