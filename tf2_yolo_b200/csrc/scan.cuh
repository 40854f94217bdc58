// Device-wide exclusive prefix sum  uint32 counts -> int64 offsets (n+1 outputs).
// One launch either way: a single block for segment tables (<= 64k counters), a chained scan with
// decoupled look-back for the per-cell counters of a batch (a few MB).
#pragma once
#include "common.cuh"

namespace yb {


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

// ---- large inputs: ONE launch, chained scan with decoupled look-back -------------------------
// Blocks of 4096 counters (8 warps x 512, coalesced) take a ticket (so a block's predecessors have started),
// publish their aggregate, and warp 0 walks back over the published aggregates / inclusive
// prefixes, 32 predecessors per step.  status[b]: bits 62..63 = 0 empty | 1 aggregate | 2 inclusive.
constexpr int kLbThreads = 256;
constexpr int kLbChunks = 4;                       // 128-counter chunks per warp (one uint4 per lane each)
constexpr int kLbWarpItems = kLbChunks * 128;      // 512 counters per warp, loaded and stored coalesced
constexpr int kLbTile = (kLbThreads / 32) * kLbWarpItems;   // 4096 counters per block
constexpr unsigned long long kLbValueMask = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

static __global__ void __launch_bounds__(kLbThreads)
scan_lookback_kernel(const unsigned int* __restrict__ in, long long n, long long* __restrict__ out,
                     unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket) {
    __shared__ long long s_wtot[kLbThreads / 32];
    __shared__ unsigned int s_bid;
    __shared__ long long s_prefix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned bid = s_bid;
    const long long wbase = (long long)bid * kLbTile + (long long)warp * kLbWarpItems;
    // lane owns counters [wbase + k*128 + lane*4, +4) of chunk k: every load / store instruction of
    // the warp covers one contiguous run
    uint4 item[kLbChunks];
    long long inc[kLbChunks], sum4[kLbChunks];
#pragma unroll
    for (int k = 0; k < kLbChunks; ++k) {
        const long long i0 = wbase + k * 128 + lane * 4;
        if (i0 + 4 <= n) {
            item[k] = *reinterpret_cast<const uint4*>(in + i0);
        } else {
            item[k].x = (i0 < n) ? in[i0] : 0u;
            item[k].y = (i0 + 1 < n) ? in[i0 + 1] : 0u;
            item[k].z = (i0 + 2 < n) ? in[i0 + 2] : 0u;
            item[k].w = 0u;
        }
        sum4[k] = (long long)item[k].x + item[k].y + item[k].z + item[k].w;
    }
    long long carry = 0;   // counters of the warp's earlier chunks
#pragma unroll
    for (int k = 0; k < kLbChunks; ++k) {
        long long v = sum4[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        inc[k] = carry + v - sum4[k];                       // exclusive inside the warp
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) s_wtot[warp] = carry;
    __syncthreads();
    long long warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kLbThreads / 32; ++w) {
        const long long t = s_wtot[w];
        if (w < warp) warp_off += t;
        total += t;
    }
    if (warp == 0) {
        long long prefix = 0;
        if (bid == 0) {
            if (lane == 0) lb_store(&status[0], (2ull << 62) | (unsigned long long)total);
        } else {
            if (lane == 0) lb_store(&status[bid], (1ull << 62) | (unsigned long long)total);
            long long look = (long long)bid - 1;   // newest predecessor of this step
            while (true) {
                const long long b = look - lane;
                unsigned long long st = (b >= 0) ? lb_load(&status[b]) : (2ull << 62);   // before block 0: inclusive 0
                // wait until every predecessor of this window has published something
                while (__any_sync(0xffffffffu, (st >> 62) == 0ull)) {
                    if ((st >> 62) == 0ull) st = lb_load(&status[b]);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
                const int first_inc = incl ? (__ffs(incl) - 1) : 32;   // nearest inclusive prefix
                long long part = (lane <= first_inc) ? (long long)(st & kLbValueMask) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                prefix += part;
                if (incl) break;
                look -= 32;
            }
            if (lane == 0) lb_store(&status[bid], (2ull << 62) | (unsigned long long)(prefix + total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const long long base = s_prefix + warp_off;
#pragma unroll
    for (int k = 0; k < kLbChunks; ++k) {
        const long long i0 = wbase + k * 128 + lane * 4;
        const long long v0 = base + inc[k], v1 = v0 + item[k].x, v2 = v1 + item[k].y, v3 = v2 + item[k].z;
        if (i0 + 4 <= n && ((reinterpret_cast<uintptr_t>(out + i0) & 15) == 0)) {
            *reinterpret_cast<longlong2*>(out + i0) = make_longlong2(v0, v1);
            *reinterpret_cast<longlong2*>(out + i0 + 2) = make_longlong2(v2, v3);
        } else {
            if (i0 < n) out[i0] = v0;
            if (i0 + 1 < n) out[i0 + 1] = v1;
            if (i0 + 2 < n) out[i0 + 2] = v2;
            if (i0 + 3 < n) out[i0 + 3] = v3;
        }
        if (i0 <= n - 1 && n - 1 < i0 + 4) out[n] = v3 + item[k].w;   // the lane that owns the last counter
    }
}

static inline size_t scan_workspace_bytes(long long n) {
    const long long n_blocks = (n + kLbTile - 1) / kLbTile;
    return align_up((size_t)(n_blocks + 2) * sizeof(long long), 256);
}

// out has n+1 entries; out[n] = total.  n >= 1.
// status_zeroed: the caller has already cleared scan_workspace_bytes(n) bytes of the workspace.
static inline int exclusive_scan_u32(const unsigned int* in, long long n, long long* out,
                                     void* workspace, cudaStream_t stream, bool status_zeroed = false) {
    if (n <= 64 * 1024) {
        scan_single_block<<<1, 1024, 0, stream>>>(in, n, out);
        YB_CUDA_TRY(cudaGetLastError());
        return YB_OK;
    }
    const int n_blocks = (int)((n + kLbTile - 1) / kLbTile);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(workspace);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(status + n_blocks);
    if (!status_zeroed)
        YB_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)(n_blocks + 1) * sizeof(long long), stream));
    scan_lookback_kernel<<<n_blocks, kLbThreads, 0, stream>>>(in, n, out, status, ticket);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

}  // namespace yb
