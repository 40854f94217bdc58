// Device-wide exclusive prefix sum  uint32 counts -> int64 offsets (n+1 outputs).
// Three small launches (block sums, scan of block sums, block scan + offset); the
// inputs here are per-cell / per-segment counters, a few MB at most.
#pragma once
#include "common.cuh"

namespace yb {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 counters per block

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* s_warp,
                                                          long long& block_total) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long warp_off = (warp == 0) ? 0 : s_warp[warp - 1];
    block_total = s_warp[(blockDim.x >> 5) - 1];
    const long long res = warp_off + inc - v;
    __syncthreads();
    return res;
}

static __global__ void scan_block_sums(const unsigned int* __restrict__ in, long long n,
                                       long long* __restrict__ block_sums) {
    __shared__ long long s_warp[32];
    const long long base = (long long)blockIdx.x * kScanTile;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const long long i = base + (long long)threadIdx.x * kScanItems + k;
        if (i < n) v += in[i];
    }
    long long total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block_sums in place, total appended at [n_blocks]
static __global__ void scan_of_block_sums(long long* __restrict__ block_sums, int n_blocks) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = (i < n_blocks) ? block_sums[i] : 0;
        long long total;
        const long long ex = block_exclusive_scan(v, s_warp, total);
        const long long carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = s_carry;
}

static __global__ void scan_apply(const unsigned int* __restrict__ in, long long n,
                                  const long long* __restrict__ block_sums,
                                  long long* __restrict__ out) {
    __shared__ long long s_warp[32];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    unsigned int item[kScanItems];
    long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        item[k] = (base + k < n) ? in[base + k] : 0u;
        v += item[k];
    }
    long long total;
    long long ex = block_exclusive_scan(v, s_warp, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += item[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[gridDim.x];
}

// small inputs (segment tables): the whole scan in one block, one launch
constexpr int kScanSmallItems = 16;
static __global__ void __launch_bounds__(1024)
scan_single_block(const unsigned int* __restrict__ in, long long n, long long* __restrict__ out) {
    __shared__ long long s_warp[32];
    long long carry = 0;
    for (long long base = 0; base < n; base += (long long)blockDim.x * kScanSmallItems) {
        const long long i0 = base + (long long)threadIdx.x * kScanSmallItems;
        unsigned int item[kScanSmallItems];
        long long v = 0;
        if (i0 + kScanSmallItems <= n) {   // 64-byte aligned run: four 128-bit loads
            const uint4* p = reinterpret_cast<const uint4*>(in + i0);
#pragma unroll
            for (int q = 0; q < kScanSmallItems / 4; ++q) {
                const uint4 t = p[q];
                item[4 * q] = t.x; item[4 * q + 1] = t.y; item[4 * q + 2] = t.z; item[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kScanSmallItems; ++k) item[k] = (i0 + k < n) ? in[i0 + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < kScanSmallItems; ++k) v += item[k];
        long long total;
        long long ex = block_exclusive_scan(v, s_warp, total) + carry;
#pragma unroll
        for (int k = 0; k < kScanSmallItems; ++k) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += item[k];
        }
        carry += total;   // every thread tracks the same running total
    }
    if (threadIdx.x == 0) out[n] = carry;
}

static inline size_t scan_workspace_bytes(long long n) {
    const long long n_blocks = (n + kScanTile - 1) / kScanTile;
    return align_up((size_t)(n_blocks + 2) * sizeof(long long), 256);
}

// out has n+1 entries; out[n] = total.  n >= 1.
static inline int exclusive_scan_u32(const unsigned int* in, long long n, long long* out,
                                     void* workspace, cudaStream_t stream) {
    if (n <= 64 * 1024) {
        scan_single_block<<<1, 1024, 0, stream>>>(in, n, out);
        return (int)cudaGetLastError();
    }
    const int n_blocks = (int)((n + kScanTile - 1) / kScanTile);
    long long* block_sums = reinterpret_cast<long long*>(workspace);
    scan_block_sums<<<n_blocks, kScanThreads, 0, stream>>>(in, n, block_sums);
    scan_of_block_sums<<<1, 1024, 0, stream>>>(block_sums, n_blocks);
    scan_apply<<<n_blocks, kScanThreads, 0, stream>>>(in, n, block_sums, out);
    return (int)cudaGetLastError();
}

}  // namespace yb
