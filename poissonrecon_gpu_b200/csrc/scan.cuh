// Exclusive scan (three kernels: tile sums -> scan of sums -> tile scan + offset), templated on
// the element loader so that counts can be derived on the fly (popc of an edge mask, a byte
// array, ...) instead of being materialised as int arrays first.
#pragma once
#include "common.cuh"

namespace prb {

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

// exclusive scan of in(0..n) into out on the context stream; the grand total is returned through
// *total_host (synchronises the stream) when total_host != nullptr.
template <class Op>
int exclusive_scan_op(Context& c, Op in, int* out, i64 n, i64* total_host) {
    if (n <= 0) { if (total_host) *total_host = 0; return PRB_OK; }
    int nb = div_up(n, kScanTile);
    DBuf<int> sums;
    PRB_TRY(sums.alloc((size_t)nb + 1, c.stream));
    PRB_LAUNCH(c, k_scan_tile_sums<Op>, nb, kScanBlock, 0, in, sums.p, n);
    PRB_LAUNCH(c, k_scan_sums, 1, 1024, 0, sums.p, nb, sums.p + nb);
    PRB_LAUNCH(c, k_scan_apply<Op>, nb, kScanBlock, 0, in, out, sums.p, n);
    if (total_host) {
        int t = 0;
        PRB_CUDA(cudaMemcpyAsync(&t, sums.p + nb, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        PRB_CUDA(cudaStreamSynchronize(c.stream));
        *total_host = t;
    }
    sums.release();
    return PRB_OK;
}

}  // namespace prb
