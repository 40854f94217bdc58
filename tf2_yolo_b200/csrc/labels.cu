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

// ---------------------------------------------------------------------------------------------
// Box lists -> label grids on the device: utils/tools.py:179-209 (_encode_to_array) for the
// finest grid and utils/tools.py:342-367 (down2xlabel) for every coarser level, in one launch.
// The grids are almost empty (a handful of boxes per image against 7581 cells x 85 channels at
// v4-608), so nothing dense is ever read: a CTA stages the image's boxes in shared memory,
// derives the sparse list of non-zero cells of every level from them, zero-fills its slice of
// the output with 128-bit stores and drops the few non-zero cells on top.  Python's floored
// division and modulo (float_divmod in CPython, npy_divmod in NumPy: same algorithm) are
// restated on fmod, which is exact.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double py_floordiv(double v, double w) {
    const double mod = fmod(v, w);
    double div = (v - mod) / w;
    if (mod != 0.0 && ((w < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, v / w);
}
__device__ __forceinline__ double py_mod(double v, double w) {
    double mod = fmod(v, w);
    if (mod != 0.0) {
        if ((w < 0.0) != (mod < 0.0)) mod += w;
    } else {
        mod = copysign(0.0, w);
    }
    return mod;
}

struct EncEntry {  // a non-zero cell of one level
    int cell;      // y * grid_w(level) + x
    int src;       // box whose w, h, obj and cell's class bits it carries; -1: copied from an empty cell
    double x, y;   // xy offset in this level's cell
};

template <typename T>
__device__ __forceinline__ void zero_fill(T* p, long long n) {
    // head up to 16-byte alignment, 128-bit body, tail
    const int tid = threadIdx.x, nt = blockDim.x;
    long long head = (long long)(((16 - ((uintptr_t)p & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    for (long long i = tid; i < head; i += nt) p[i] = (T)0;
    constexpr int per = 16 / sizeof(T);
    const long long body = (n - head) / per;
    uint4* q = reinterpret_cast<uint4*>(p + head);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = tid; i < body; i += nt) q[i] = z;
    for (long long i = head + body * per + tid; i < n; i += nt) p[i] = (T)0;
}

template <typename T>
__global__ void __launch_bounds__(256)
encode_labels_kernel(const double* __restrict__ boxes, const int64_t* __restrict__ box_offsets, int max_boxes,
                     double img_h, double img_w, int grid_h, int grid_w, int class_num, int n_levels,
                     T* out0, T* out1, T* out2, T* out3, int n_chunks, unsigned long long* n_bad) {
    extern __shared__ __align__(16) unsigned char enc_smem[];
    // per box: cell on the finest grid (-1: not written), class, [x_off, y_off, w, h]
    double* s_val = reinterpret_cast<double*>(enc_smem);                          // [max_boxes][4]
    EncEntry* s_ent = reinterpret_cast<EncEntry*>(s_val + 4 * (size_t)max_boxes);  // [n_levels][max_boxes]
    int* s_cell = reinterpret_cast<int*>(s_ent + (size_t)n_levels * max_boxes);    // [max_boxes]
    int* s_label = s_cell + max_boxes;                                             // [max_boxes]
    __shared__ int s_count[YB_MAX_SCALES];
    __shared__ unsigned int s_bad;

    const int img = blockIdx.x / n_chunks, chunk = blockIdx.x - img * n_chunks;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, n_warps = nt >> 5;
    const int ch = 5 + class_num;
    const int64_t b0 = box_offsets[img];
    const int64_t n_all = box_offsets[img + 1] - b0;
    const int n = (int)(n_all < (int64_t)max_boxes ? (n_all < 0 ? 0 : n_all) : max_boxes);
    if (tid < YB_MAX_SCALES) s_count[tid] = 0;
    if (tid == 0) s_bad = (n_all > max_boxes) ? (unsigned int)(n_all - max_boxes) : 0u;
    __syncthreads();

    // ---- boxes -> finest-grid cells (tools.py:185-209) -----------------------------------------
    const double cell_h = img_h / (double)grid_h, cell_w = img_w / (double)grid_w;
    for (int b = tid; b < n; b += nt) {
        const double* r = boxes + (b0 + b) * 5;
        const double x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3], lab = r[4];
        const double bx = x1 + (x2 - x1) / 2.0, by = y1 + (y2 - y1) / 2.0;
        const double bw = x2 - x1, bh = y2 - y1;
        int cell = -1;
        // non-finite corners make int() raise in the reference; so does a class outside [0, C)
        const bool finite = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(bx) &&
                            isfinite(by) && lab >= 0.0 && lab < (double)class_num;
        if (!finite) {
            atomicAdd(&s_bad, 1u);
        } else {
            const double xf = py_floordiv(bx, cell_w), yf = py_floordiv(by, cell_h);
            if (xf < (double)grid_w && yf < (double)grid_h) {        // :199
                if (xf < -(double)grid_w || yf < -(double)grid_h) {
                    atomicAdd(&s_bad, 1u);                            // IndexError in the reference
                } else {
                    int xi = (int)xf, yi = (int)yf;
                    if (xi < 0) xi += grid_w;                         // NumPy negative index
                    if (yi < 0) yi += grid_h;
                    cell = yi * grid_w + xi;
                    s_val[4 * b + 0] = py_mod(bx, cell_w) / cell_w;
                    s_val[4 * b + 1] = py_mod(by, cell_h) / cell_h;
                    s_val[4 * b + 2] = bw / img_w;
                    s_val[4 * b + 3] = bh / img_h;
                }
            }
        }
        s_cell[b] = cell;
        s_label[b] = finite ? (int)lab : 0;
    }
    __syncthreads();
    if (chunk == 0 && tid == 0 && s_bad && n_bad) atomicAdd(n_bad, (unsigned long long)s_bad);

    // ---- level 0: the LAST box written to a cell owns its x, y, w, h -----------------------------
    for (int b = tid; b < n; b += nt) {
        const int cell = s_cell[b];
        if (cell < 0) continue;
        bool last = true;
        for (int j = b + 1; j < n; ++j)
            if (s_cell[j] == cell) {
                last = false;
                break;
            }
        if (last) {
            const int k = atomicAdd(&s_count[0], 1);
            s_ent[k] = EncEntry{cell, b, s_val[4 * b + 0], s_val[4 * b + 1]};
        }
    }
    __syncthreads();

    // ---- coarser levels: down2xlabel on the sparse list (tools.py:355-366) ----------------------
    int gw_prev = grid_w;
    for (int l = 1; l < n_levels; ++l) {
        const EncEntry* prev = s_ent + (size_t)(l - 1) * max_boxes;
        EncEntry* cur = s_ent + (size_t)l * max_boxes;
        const int m = s_count[l - 1];
        const int gw_cur = gw_prev >> 1;
        for (int e = tid; e < m; e += nt) {
            if (prev[e].src < 0) continue;                       // obj == 0: cannot make the block's max 1
            const int py = (prev[e].cell / gw_prev) >> 1, px = (prev[e].cell % gw_prev) >> 1;
            bool first = true;                                    // one writer per 2x2 block
            for (int j = 0; j < e; ++j)
                if (prev[j].src >= 0 && ((prev[j].cell / gw_prev) >> 1) == py && ((prev[j].cell % gw_prev) >> 1) == px) {
                    first = false;
                    break;
                }
            if (!first) continue;
            int pick = 0, pick_e = -1;
            double best = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {                         // np.argmax of w*h over the crop, row-major
                const int child = (2 * py + (q >> 1)) * gw_prev + 2 * px + (q & 1);
                int ce = -1;
                for (int j = 0; j < m; ++j)
                    if (prev[j].cell == child) {
                        ce = j;
                        break;
                    }
                double area = 0.0;
                if (ce >= 0 && prev[ce].src >= 0) area = __dmul_rn(s_val[4 * prev[ce].src + 2], s_val[4 * prev[ce].src + 3]);
                if (q == 0 || area > best) {
                    best = area;
                    pick = q;
                    pick_e = ce;
                }
            }
            EncEntry o;
            o.cell = py * gw_cur + px;
            o.src = pick_e >= 0 ? prev[pick_e].src : -1;
            o.x = ((pick_e >= 0 ? prev[pick_e].x : 0.0) + (double)(pick & 1)) / 2.0;
            o.y = ((pick_e >= 0 ? prev[pick_e].y : 0.0) + (double)(pick >> 1)) / 2.0;
            cur[atomicAdd(&s_count[l], 1)] = o;
        }
        __syncthreads();
        gw_prev = gw_cur;
    }

    // ---- this CTA's slice of every level: zero-fill, then the non-zero cells on top -------------
    T* outs[YB_MAX_SCALES] = {out0, out1, out2, out3};
    int gh_l = grid_h, gw_l = grid_w;
    for (int l = 0; l < n_levels; ++l) {
        const long long cells = (long long)gh_l * gw_l;
        const long long lo = cells * chunk / n_chunks, hi = cells * (chunk + 1) / n_chunks;
        T* slab = outs[n_levels - 1 - l] + (long long)img * cells * ch;   // outs[] is coarse first
        zero_fill(slab + lo * ch, (hi - lo) * ch);
        gh_l >>= 1;
        gw_l >>= 1;
    }
    __syncthreads();
    gh_l = grid_h;
    gw_l = grid_w;
    for (int l = 0; l < n_levels; ++l) {
        const long long cells = (long long)gh_l * gw_l;
        const long long lo = cells * chunk / n_chunks, hi = cells * (chunk + 1) / n_chunks;
        T* slab = outs[n_levels - 1 - l] + (long long)img * cells * ch;
        const EncEntry* ent = s_ent + (size_t)l * max_boxes;
        const int m = s_count[l];
        for (int e = warp; e < m; e += n_warps) {
            const EncEntry en = ent[e];
            if (en.cell < lo || en.cell >= hi) continue;
            T* dst = slab + (long long)en.cell * ch;
            if (lane == 0) dst[0] = (T)en.x;
            if (lane == 1) dst[1] = (T)en.y;
            if (en.src >= 0) {
                if (lane == 2) dst[2] = (T)s_val[4 * en.src + 2];
                if (lane == 3) dst[3] = (T)s_val[4 * en.src + 3];
                if (lane == 4) dst[4] = (T)1;
                const int fine = s_cell[en.src];                  // class bits of every box of that cell
                for (int j = lane; j < n; j += 32)
                    if (s_cell[j] == fine) dst[5 + s_label[j]] = (T)1;
            }
        }
        gh_l >>= 1;
        gw_l >>= 1;
    }
}

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
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
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
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" int yb_encode_labels(const double* boxes, const int64_t* box_offsets, int64_t n_img,
                                int max_boxes_per_img, double img_h, double img_w, int grid_h, int grid_w,
                                int class_num, int n_levels, void* const* out_levels_host, int out_f64,
                                unsigned long long* n_bad, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_img < 0 || grid_h <= 0 || grid_w <= 0 || class_num < 1 || n_levels < 1 || n_levels > YB_MAX_SCALES)
        return YB_E_SHAPE;
    if (max_boxes_per_img < 0 || max_boxes_per_img > YB_ENCODE_MAX_BOXES) return YB_E_SHAPE;
    if (!(img_h > 0.0) || !(img_w > 0.0)) return YB_E_PARAM;
    const int div = 1 << (n_levels - 1);
    if (n_levels > 1 && ((grid_h % div) || (grid_w % div))) return YB_E_SHAPE;   // down2xlabel needs even grids
    if (n_img == 0) return YB_OK;
    if (out_levels_host == nullptr || box_offsets == nullptr) return YB_E_NULL;
    void* outs[YB_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
    for (int l = 0; l < n_levels; ++l) {
        outs[l] = out_levels_host[l];
        if (outs[l] == nullptr) return YB_E_NULL;
        if ((uintptr_t)outs[l] & (out_f64 ? 7 : 3)) return YB_E_ALIGN;
    }
    const int cap = max_boxes_per_img < 1 ? 1 : max_boxes_per_img;
    const size_t smem = (size_t)cap * (4 * sizeof(double) + (size_t)n_levels * sizeof(EncEntry) + 2 * sizeof(int));
    // enough CTAs to fill the GPU; every CTA of an image rebuilds the (tiny) sparse lists
    long long chunks = (4LL * kNumSMs + n_img - 1) / n_img;
    const long long coarse_cells = (long long)(grid_h / div) * (grid_w / div);
    if (chunks > coarse_cells) chunks = coarse_cells;
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    if (n_img * chunks > 0x7fffffffLL) return YB_E_SHAPE;
    const unsigned int grid = (unsigned int)(n_img * chunks);
    if (out_f64) {
        static SmemRaised done;
        YB_CUDA_TRY(raise_dynamic_smem_once(encode_labels_kernel<double>, (int)smem, &done));
        encode_labels_kernel<double><<<grid, 256, smem, stream>>>(
            boxes, box_offsets, cap, img_h, img_w, grid_h, grid_w, class_num, n_levels, (double*)outs[0],
            (double*)outs[1], (double*)outs[2], (double*)outs[3], (int)chunks, n_bad);
    } else {
        static SmemRaised done;
        YB_CUDA_TRY(raise_dynamic_smem_once(encode_labels_kernel<float>, (int)smem, &done));
        encode_labels_kernel<float><<<grid, 256, smem, stream>>>(
            boxes, box_offsets, cap, img_h, img_w, grid_h, grid_w, class_num, n_levels, (float*)outs[0],
            (float*)outs[1], (float*)outs[2], (float*)outs[3], (int)chunks, n_bad);
    }
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}
