// Label-side helpers next to the hot path:
//   down2x labels : utils/tools.py:342-367 (down2xlabel) - the Python triple loop the YOLOv4 data
//                   sequence runs per batch and per extra scale (yolov4/__init__.py:47-53)
//   column sums   : the data-sized part of utils/tools.py:592-627 (get_class_weight)
#include "common.cuh"

namespace yb {

template <typename T>
__device__ __forceinline__ T mul1(T a, T b);
template <>
__device__ __forceinline__ float mul1<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul1<double>(double a, double b) { return __dmul_rn(a, b); }

// one warp per output cell: pick the largest-area entry of the 2x2 block (first maximum in
// row-major order, areas in the INPUT dtype like NumPy), re-express its xy offset in the
// coarser cell, copy the remaining channels; blocks without an object stay zero.
template <typename T>
__global__ void __launch_bounds__(256)
down2x_kernel(const T* __restrict__ in, long long n_img, int gh, int gw, int ch, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int oh = gh / 2, ow = gw / 2;
    const long long total = n_img * oh * ow;
    for (long long o = warp; o < total; o += n_warps) {
        const long long img = o / ((long long)oh * ow);
        const int rem = (int)(o - img * oh * ow);
        const int oy = rem / ow, ox = rem - oy * ow;
        const T* base = in + ((img * gh + 2 * oy) * gw + 2 * ox) * (long long)ch;
        const T* cell[4] = {base, base + ch, base + (long long)gw * ch, base + (long long)gw * ch + ch};
        bool any = false;
        int pick = 0;
        T best = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            any = any || (cell[q][4] == (T)1);
            const T area = mul1<T>(cell[q][2], cell[q][3]);
            if (q == 0 || area > best) {
                best = area;
                pick = q;
            }
        }
        // crop[..., 4].max() == 1: the MAXIMUM must equal one
        T mx = cell[0][4];
#pragma unroll
        for (int q = 1; q < 4; ++q) mx = (cell[q][4] > mx) ? cell[q][4] : mx;
        any = (mx == (T)1);
        double* dst = out + o * (long long)ch;
        if (!any) {
            for (int k = lane; k < ch; k += 32) dst[k] = 0.0;
        } else {
            const T* src = cell[pick];
            for (int k = lane; k < ch; k += 32) {
                double v = (double)src[k];
                if (k == 0) v = (v + (double)(pick & 1)) / 2.0;
                if (k == 1) v = (v + (double)(pick >> 1)) / 2.0;
                dst[k] = v;
            }
        }
    }
}

// per-column sums of a (rows, cols) matrix, fp64 accumulation, deterministic two-stage reduce
template <typename T>
__global__ void __launch_bounds__(256)
column_sums_kernel(const T* __restrict__ data, long long rows, int cols, double* __restrict__ partials,
                   unsigned int* counter, double* __restrict__ out) {
    extern __shared__ double s_part[];  // [warps][cols]
    __shared__ int s_is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int c = threadIdx.x; c < n_warps * cols; c += blockDim.x) s_part[c] = 0.0;
    __syncthreads();
    // a warp walks rows; lanes stride the columns (coalesced)
    for (int c0 = 0; c0 < cols; c0 += 32) {
        const int c = c0 + lane;
        double acc = 0.0;
        if (c < cols)
            for (long long r = (long long)blockIdx.x * n_warps + warp; r < rows; r += (long long)gridDim.x * n_warps)
                acc += (double)data[r * cols + c];
        if (c < cols) s_part[warp * cols + c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < n_warps; ++w) s += s_part[w * cols + c];
        partials[(size_t)blockIdx.x * cols + c] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(&partials[(size_t)b * cols + c]);
        out[c] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;
}

constexpr int kColGrid = kNumSMs * 2;

}  // namespace yb

using namespace yb;

extern "C" int yb_down2x_labels(const void* labels, int is_f64, int64_t n_img, int grid_h, int grid_w,
                                int channels, double* out, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_img < 0 || grid_h <= 0 || grid_w <= 0 || channels < 5) return YB_E_SHAPE;
    if ((grid_h & 1) || (grid_w & 1)) return YB_E_SHAPE;  // the reference indexes out of range on odd grids
    if (n_img == 0) return YB_OK;
    if (labels == nullptr || out == nullptr) return YB_E_NULL;
    const long long total = (long long)n_img * (grid_h / 2) * (grid_w / 2);
    const int threads = 256;
    const int blocks = (int)min((long long)kNumSMs * 8, (total * 32 + threads - 1) / threads);
    if (is_f64)
        down2x_kernel<double><<<blocks, threads, 0, stream>>>(reinterpret_cast<const double*>(labels), n_img, grid_h,
                                                              grid_w, channels, out);
    else
        down2x_kernel<float><<<blocks, threads, 0, stream>>>(reinterpret_cast<const float*>(labels), n_img, grid_h,
                                                             grid_w, channels, out);
    return (int)cudaGetLastError();
}

extern "C" size_t yb_column_sums_workspace_bytes(int cols) {
    if (cols <= 0) return 0;
    return align_up((size_t)kColGrid * cols * sizeof(double), 256) + 256;
}

extern "C" int yb_column_sums(const void* data, int is_f64, int64_t rows, int cols, double* out, void* workspace,
                              size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (rows < 0 || cols <= 0 || cols > 4096) return YB_E_SHAPE;
    if (out == nullptr || workspace == nullptr || (rows > 0 && data == nullptr)) return YB_E_NULL;
    if (workspace_bytes < yb_column_sums_workspace_bytes(cols) || ((uintptr_t)workspace & 255)) return YB_E_WORKSPACE;
    double* partials = reinterpret_cast<double*>(workspace);
    unsigned int* counter =
        reinterpret_cast<unsigned int*>((char*)workspace + align_up((size_t)kColGrid * cols * sizeof(double), 256));
    YB_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    const int threads = 256;
    const size_t smem = (size_t)(threads / 32) * cols * sizeof(double);
    const int grid = (int)max(1LL, min((long long)kColGrid, ((long long)rows + 7) / 8));
    if (is_f64) {
        YB_CUDA_TRY(cudaFuncSetAttribute(column_sums_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        column_sums_kernel<double><<<grid, threads, smem, stream>>>(reinterpret_cast<const double*>(data), rows, cols,
                                                                    partials, counter, out);
    } else {
        YB_CUDA_TRY(cudaFuncSetAttribute(column_sums_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        column_sums_kernel<float><<<grid, threads, smem, stream>>>(reinterpret_cast<const float*>(data), rows, cols,
                                                                   partials, counter, out);
    }
    return (int)cudaGetLastError();
}
