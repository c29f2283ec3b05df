// Exclusive scan (three kernels: tile sums -> scan of sums -> tile scan + offset), templated on
// the element loader so that counts can be derived on the fly (popc of an edge mask, a byte
// array, ...) instead of being materialised as int arrays first.
#pragma once
#include "common.cuh"

namespace prb {

int scan_work_ensure(Context& c, size_t tiles);     // octree.cu: descriptors + ticket of the context (grow-only)


constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

struct ScanLoadInt { const int* p; __device__ __forceinline__ int operator()(i64 i) const { return p[i]; } };
struct ScanLoadU8 { const unsigned char* p; __device__ __forceinline__ int operator()(i64 i) const { return p[i]; } };
struct ScanLoadPopc16 { const unsigned short* p; __device__ __forceinline__ int operator()(i64 i) const { return __popc((unsigned)p[i]); } };

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /* >= 33 ints */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        int s = lane < nw ? smem[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < nw) smem[lane] = si - s;
        if (lane == nw - 1) smem[32] = si;
    }
    __syncthreads();
    int r = inc - v + smem[w];
    *total = smem[32];
    __syncthreads();
    return r;
}

template <class Op>
__global__ void __launch_bounds__(kScanBlock) k_scan_tile_sums(Op in, int* __restrict__ sums, i64 n) {
    i64 t0 = (i64)blockIdx.x * kScanTile;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        i64 i = t0 + k * kScanBlock + threadIdx.x;
        if (i < n) s += in(i);
    }
    __shared__ int sm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (kScanBlock >> 5) ? sm[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}
static __global__ void __launch_bounds__(1024) k_scan_sums(int* __restrict__ sums, int nb, int* __restrict__ total) {
    __shared__ int sm[33];
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        int i = b0 + threadIdx.x;
        int v = i < nb ? sums[i] : 0, tot;
        int e = block_exclusive_scan(v, &tot, sm);
        if (i < nb) sums[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}
template <class Op>
__global__ void __launch_bounds__(kScanBlock) k_scan_apply(Op in, int* __restrict__ out, const int* __restrict__ sums, i64 n) {
    __shared__ int sm[33];
    i64 t0 = (i64)blockIdx.x * kScanTile;
    int carry = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        i64 i = t0 + k * kScanBlock + threadIdx.x;
        int v = i < n ? in(i) : 0, tot;
        int e = block_exclusive_scan(v, &tot, sm);
        if (i < n) out[i] = carry + e;
        carry += tot;
    }
}

// ---- single-pass chained scan with decoupled look-back.  Tiles take their index from a ticket counter (a tile only ever waits
// for tiles that have started), publish their aggregate, then walk back over their predecessors -- 32 at a time, one per lane of
// warp 0 -- until one with an inclusive prefix is met.  A descriptor is one 64-bit word (status << 32 | value), written and read
// atomically; status = 2 * epoch + {0: aggregate, 1: inclusive prefix}: the per-call epoch makes clearing the descriptors
// between calls unnecessary, and the last tile rewinds the ticket counter.
__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void scan_st(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory"); }
template <class Op>
__global__ void __launch_bounds__(kScanBlock) k_scan_lookback(Op in, int* out /* may alias the input */, i64 n, int nTiles, unsigned long long* __restrict__ desc, unsigned* __restrict__ ticket,
                                                              unsigned epoch, int* __restrict__ total, int* __restrict__ hostTotal /* pinned, or null */) {
    __shared__ int sm[33];
    __shared__ int sTile, sPrefix;
    if (threadIdx.x == 0) sTile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = sTile;
    // thread t owns the kScanItems consecutive elements [t0 + t * kScanItems, ...): a warp covers one contiguous run
    const i64 i0 = (i64)tile * kScanTile + (i64)threadIdx.x * kScanItems;
    int v[kScanItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = (i0 + k < n) ? in(i0 + k) : 0; sum += v[k]; }
    int agg;
    const int excl = block_exclusive_scan(sum, &agg, sm);
    const unsigned stA = 2u * epoch, stP = 2u * epoch + 1u;
    if (threadIdx.x < 32) {
        int prefix = 0;
        if (tile == 0) {
            if (threadIdx.x == 0) scan_st(desc, ((unsigned long long)stP << 32) | (unsigned)agg);
        } else {
            if (threadIdx.x == 0) scan_st(desc + tile, ((unsigned long long)stA << 32) | (unsigned)agg);
            for (int j = tile - 1;; j -= 32) {
                const int mine = j - (int)threadIdx.x;
                unsigned st = stP;
                int val = 0;
                if (mine >= 0) {
                    unsigned long long d;
                    do { d = scan_ld(desc + mine); st = (unsigned)(d >> 32); } while (st != stA && st != stP);
                    val = (int)(unsigned)(d & 0xffffffffu);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, st == stP);      // (lanes past tile 0 count as inclusive with value 0)
                const int first = incl ? __ffs(incl) - 1 : 31;                      // nearest predecessor with an inclusive prefix (none: the whole window counts)
                int part = ((int)threadIdx.x <= first) ? val : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (incl) break;
            }
            if (threadIdx.x == 0) scan_st(desc + tile, ((unsigned long long)stP << 32) | (unsigned)(prefix + agg));
        }
        if (threadIdx.x == 0) {
            sPrefix = prefix;
            if (tile == nTiles - 1) {                                           // every ticket has been drawn: rewind for the next call
                *total = prefix + agg;
                *ticket = 0u;
                if (hostTotal) *hostTotal = prefix + agg;                       // straight into pinned host memory: no copy-engine round trip
            }
        }
    }
    __syncthreads();
    int run = sPrefix + excl;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (i0 + k < n) out[i0 + k] = run;
        run += v[k];
    }
}

// exclusive scan of in(0..n) into out on the context stream; the grand total is returned through
// *total_host (synchronises the stream) when total_host != nullptr.  The last tile stores the total into the pinned word
// c.hScanTotal[hostSlot] itself (hostSlot >= 0), so the host only waits for the stream; a caller that passes hostSlot > 0 without
// total_host reads c.hScanTotal[hostSlot] after its own later synchronisation.
template <class Op>
int exclusive_scan_op(Context& c, Op in, int* out, i64 n, i64* total_host, int hostSlot = 0) {
    if (n <= 0) {
        if (total_host) *total_host = 0;
        else if (hostSlot > 0 && c.hScanTotal) c.hScanTotal[hostSlot] = 0;
        return PRB_OK;
    }
    int nb = div_up(n, kScanTile);
    PRB_TRY(scan_work_ensure(c, (size_t)nb));
    ScanWork& w = c.scanWork;
    w.epoch++;
    int* hostWord = (total_host || hostSlot > 0) ? c.hScanTotal + hostSlot : nullptr;
    PRB_LAUNCH(c, k_scan_lookback<Op>, nb, kScanBlock, 0, in, out, n, nb, w.desc, w.ticket, w.epoch, (int*)(w.ticket + 1), hostWord);
    if (total_host) {
        PRB_CUDA(cudaStreamSynchronize(c.stream));
        *total_host = ((volatile int*)c.hScanTotal)[hostSlot];
    }
    return PRB_OK;
}

}  // namespace prb
