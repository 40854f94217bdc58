// Anchor k-means: one Lloyd assignment + accumulation pass over all boxes.
// Replaces utils/kmeans.py:79-90 (distance matrix, argmin, per-cluster mean) with a
// single streaming kernel; distances follow kmeans.py:9-33 (iou_dist: 1 - area
// ratio, NOT box overlap) and :36-40 (euclidean), float64, one rounding per
// operation (-fmad=false) so assignments are bit-identical to NumPy's argmin
// (first minimum wins).
//
// Data tiles (16 B per box for d=2) are staged into a shared-memory ring with 1-D
// bulk-async copies (TMA); every thread keeps k*(d+1) accumulators in registers
// (compile-time k bound, predicated adds - no atomics in the hot loop), reduced
// warp -> CTA -> global partials -> last CTA in a fixed order (deterministic).
#include "common.cuh"

namespace yb {

constexpr int kKmThreads = 256;
constexpr int kKmStages = 3;
constexpr int kKmTilePts = 2048;  // points per stage (32 KB at d=2)

struct KmLaunch {
    const double* data;
    long long n;
    const double* centers;
    int k, d, kind;
    int bulk_ok;
    int* assign;
    double* sums;       // [k][d]
    long long* counts;  // [k]
    double* partials;   // [grid][k*(d+1)]
    unsigned int* counter;
};

template <int KMAX, int D>
__global__ void __launch_bounds__(kKmThreads)
kmeans_assign_kernel(const __grid_constant__ KmLaunch L) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* ring = reinterpret_cast<double*>(smem);  // [stages][tile*D]
    __shared__ uint64_t full[kKmStages];
    __shared__ double s_center[KMAX * D];
    __shared__ double s_carea[KMAX];
    __shared__ double s_red[(kKmThreads / 32) * KMAX * (D + 1)];
    __shared__ int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = L.k;

    if (tid == 0) {
        for (int i = 0; i < kKmStages; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    if (tid < KMAX * D) s_center[tid] = (tid < k * D) ? L.centers[tid] : 0.0;
    __syncthreads();
    if (tid < KMAX) s_carea[tid] = (tid < k && D >= 2) ? __dmul_rn(s_center[tid * D], s_center[tid * D + 1]) : 0.0;
    __syncthreads();

    double acc[KMAX][D];
    int cnt[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
        cnt[c] = 0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc[c][j] = 0.0;
    }

    const long long n_tiles = (L.n + kKmTilePts - 1) / kKmTilePts;
    const long long n_my = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](long long t) {
        const long long tile = blockIdx.x + t * gridDim.x;
        const long long p0 = tile * kKmTilePts;
        const int np = (int)min((long long)kKmTilePts, L.n - p0);
        const int stage = (int)(t % kKmStages);
        double* dst = ring + (size_t)stage * kKmTilePts * D;
        const uint32_t bytes = (uint32_t)np * D * 8u;
        if (L.bulk_ok && (bytes & 15u) == 0u) {
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(dst, L.data + p0 * D, bytes, &full[stage]);
        } else {
            mbar_arrive(&full[stage]);  // consumers read global memory directly for this tile
        }
    };
    if (tid == 0)
        for (long long t = 0; t < min((long long)(kKmStages - 1), n_my); ++t) issue(t);

    for (long long it = 0; it < n_my; ++it) {
        if (tid == 0 && it + kKmStages - 1 < n_my) issue(it + kKmStages - 1);
        const long long tile = blockIdx.x + it * gridDim.x;
        const long long p0 = tile * kKmTilePts;
        const int np = (int)min((long long)kKmTilePts, L.n - p0);
        const int stage = (int)(it % kKmStages);
        const bool staged = L.bulk_ok && ((((uint32_t)np * D * 8u) & 15u) == 0u);
        const double* src = staged ? ring + (size_t)stage * kKmTilePts * D : L.data + p0 * D;
        mbar_wait(&full[stage], (uint32_t)((it / kKmStages) & 1));
        for (int i = tid; i < np; i += kKmThreads) {
            double v[D];
#pragma unroll
            for (int j = 0; j < D; ++j) v[j] = src[(size_t)i * D + j];
            int best = 0;
            double bd = 0.0;
            if (L.kind == YB_DIST_IOU) {
                const double a = __dmul_rn(v[0], v[1]);
#pragma unroll
                for (int c = 0; c < KMAX; ++c) {
                    if (c < k) {
                        const double ca = s_carea[c];
                        const double dist = 1.0 - fmin(ca, a) / fmax(ca, a);
                        if (c == 0 || dist < bd) {
                            bd = dist;
                            best = c;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < KMAX; ++c) {
                    if (c < k) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            const double df = s_center[c * D + j] - v[j];
                            s = __dadd_rn(s, __dmul_rn(df, df));
                        }
                        const double dist = sqrt(s);
                        if (c == 0 || dist < bd) {
                            bd = dist;
                            best = c;
                        }
                    }
                }
            }
            if (L.assign != nullptr) L.assign[p0 + i] = best;
#pragma unroll
            for (int c = 0; c < KMAX; ++c) {
                const bool hit = (best == c);
                cnt[c] += hit ? 1 : 0;
#pragma unroll
                for (int j = 0; j < D; ++j) acc[c][j] += hit ? v[j] : 0.0;
            }
        }
        __syncthreads();  // stage free for the next bulk load
    }

    // ---- reduction: thread -> warp -> CTA -> global partials -> last CTA ----
    constexpr int NV = KMAX * (D + 1);
#pragma unroll
    for (int c = 0; c < KMAX; ++c) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double s = warp_sum(acc[c][j]);
            if (lane == 0) s_red[warp * NV + c * (D + 1) + j] = s;
        }
        const int cs = warp_sum(cnt[c]);
        if (lane == 0) s_red[warp * NV + c * (D + 1) + D] = (double)cs;  // exact below 2^53
    }
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
        for (int w = 0; w < kKmThreads / 32; ++w) s += s_red[w * NV + tid];
        L.partials[(size_t)blockIdx.x * NV + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(L.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    for (int i = warp; i < NV; i += kKmThreads / 32) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&L.partials[(size_t)b * NV + i]);
        s = warp_sum(s);
        if (lane == 0) {
            const int c = i / (D + 1), j = i - c * (D + 1);
            if (c < k) {
                if (j < D) L.sums[c * D + j] = s;
                else L.counts[c] = (long long)s;
            }
        }
    }
    if (tid == 0) *L.counter = 0u;
}

__global__ void minmax_kernel(const double* __restrict__ x, long long n, double* __restrict__ partials,
                              unsigned int* counter, double* __restrict__ out2) {
    __shared__ double s_lo[32], s_hi[32];
    __shared__ int s_is_last;
    double lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double v = x[i];
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
        partials[2 * blockIdx.x] = lo;
        partials[2 * blockIdx.x + 1] = hi;
        __threadfence();
        s_is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last || threadIdx.x != 0) return;
    __threadfence();
    lo = INFINITY; hi = -INFINITY;
    for (int b = 0; b < (int)gridDim.x; ++b) {
        lo = fmin(lo, __ldcg(&partials[2 * b]));
        hi = fmax(hi, __ldcg(&partials[2 * b + 1]));
    }
    out2[0] = lo;
    out2[1] = hi;
    *counter = 0u;
}

constexpr int kKmGrid = kNumSMs * 2;

template <int KMAX, int D>
static int launch_km(const KmLaunch& L, cudaStream_t stream) {
    const size_t smem = (size_t)kKmStages * kKmTilePts * D * sizeof(double);
    YB_CUDA_TRY(cudaFuncSetAttribute(kmeans_assign_kernel<KMAX, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long n_tiles = (L.n + kKmTilePts - 1) / kKmTilePts;
    const int grid = (int)max(1LL, min((long long)kKmGrid, n_tiles));
    kmeans_assign_kernel<KMAX, D><<<grid, kKmThreads, smem, stream>>>(L);
    return (int)cudaGetLastError();
}

template <int D>
static int launch_km_d(const KmLaunch& L, cudaStream_t stream) {
    if (L.k <= 4) return launch_km<4, D>(L, stream);
    if (L.k <= 8) return launch_km<8, D>(L, stream);
    if (L.k <= 12) return launch_km<12, D>(L, stream);
    return launch_km<16, D>(L, stream);
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_kmeans_workspace_bytes(int64_t n_points, int k, int n_dim) {
    (void)n_points;
    if (k < 1 || n_dim < 1) return 0;
    return align_up((size_t)kKmGrid * 16 * (n_dim + 1) * sizeof(double), 256) + 256;
}

extern "C" int yb_kmeans_assign(const double* data, int64_t n_points, int n_dim, const double* centers,
                                int k, int dist_kind, int32_t* assign, double* sums, int64_t* counts,
                                void* workspace, size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (centers == nullptr || sums == nullptr || counts == nullptr || workspace == nullptr) return YB_E_NULL;
    if (n_points > 0 && data == nullptr) return YB_E_NULL;
    if (n_points < 0 || k < 1 || k > 16 || n_dim < 1 || n_dim > 4) return YB_E_SHAPE;
    if (dist_kind != YB_DIST_IOU && dist_kind != YB_DIST_EUCLID) return YB_E_PARAM;
    if (dist_kind == YB_DIST_IOU && n_dim < 2) return YB_E_SHAPE;
    if ((uintptr_t)data & 7) return YB_E_ALIGN;
    if (workspace_bytes < yb_kmeans_workspace_bytes(n_points, k, n_dim) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    KmLaunch L;
    L.data = data;
    L.n = n_points;
    L.centers = centers;
    L.k = k;
    L.d = n_dim;
    L.kind = dist_kind;
    L.bulk_ok = (((uintptr_t)data & 15) == 0) ? 1 : 0;
    L.assign = assign;
    L.sums = sums;
    L.counts = reinterpret_cast<long long*>(counts);
    L.partials = reinterpret_cast<double*>(workspace);
    L.counter = reinterpret_cast<unsigned int*>(
        (char*)workspace + align_up((size_t)kKmGrid * 16 * (n_dim + 1) * sizeof(double), 256));
    YB_CUDA_TRY(cudaMemsetAsync(L.counter, 0, sizeof(unsigned int), stream));
    switch (n_dim) {
        case 1: return launch_km_d<1>(L, stream);
        case 2: return launch_km_d<2>(L, stream);
        case 3: return launch_km_d<3>(L, stream);
        default: return launch_km_d<4>(L, stream);
    }
}

extern "C" int yb_minmax_f64(const double* data, int64_t n, double* out2, void* workspace,
                             size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (data == nullptr || out2 == nullptr || workspace == nullptr) return YB_E_NULL;
    if (n < 1) return YB_E_SHAPE;
    const int grid = kNumSMs * 4;
    if (workspace_bytes < (size_t)grid * 16 + 256 || ((uintptr_t)workspace & 255)) return YB_E_WORKSPACE;
    double* partials = reinterpret_cast<double*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + align_up((size_t)grid * 16, 256));
    YB_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    minmax_kernel<<<grid, 256, 0, stream>>>(data, n, partials, counter, out2);
    return (int)cudaGetLastError();
}
