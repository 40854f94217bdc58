// PR-curve / mAP matching: per (image, class) every detection takes the maximum
// and arg-maximum IoU over the ground truths of its class
// (utils/measurement.py:252-283 PRfunc, :104-130 create_score_mat), float64 in the
// reference's operation order (utils/tools.py:649-666; -fmad=false).
#include "common.cuh"

namespace yb {

constexpr double kMapEps = 1e-07;

__device__ __forceinline__ double map_iou(double tx, double ty, double tw, double th, double px,
                                          double py, double pw, double ph) {
    const double thw = tw / 2.0, thh = th / 2.0, phw = pw / 2.0, phh = ph / 2.0;
    const double iw = fmax(fmin(px + phw, tx + thw) - fmax(px - phw, tx - thw), 0.0);
    const double ih = fmax(fmin(py + phh, ty + thh) - fmax(py - phh, ty - thh), 0.0);
    const double inter = iw * ih;
    const double uni = pw * ph + tw * th - inter;
    return inter / (uni + kMapEps);
}

// one thread per detection row; ground truths of an image are few (tens)
__global__ void map_match_kernel(const double* __restrict__ gt, const long long* __restrict__ gt_off,
                                 const double* __restrict__ det, const long long* __restrict__ det_off,
                                 long long n_img, long long det_cap, double* __restrict__ best_iou,
                                 int* __restrict__ best_gt) {
    long long n_det = det_off[n_img];
    if (n_det > det_cap) n_det = det_cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_det;
         i += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = n_img;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (det_off[mid] <= i) lo = mid; else hi = mid;
        }
        const double* d = det + i * 7;
        const long long cls = (long long)d[5];
        double best = -1.0;
        int arg = -1, seen = 0;
        for (long long g = gt_off[lo]; g < gt_off[lo + 1]; ++g) {
            const double* t = gt + g * 7;
            if ((long long)t[5] != cls) continue;
            const double v = map_iou(t[0], t[1], t[2], t[3], d[0], d[1], d[2], d[3]);
            if (arg < 0 || v > best) {  // first maximum, like np.argmax
                best = v;
                arg = seen;
            }
            ++seen;
        }
        best_iou[i] = best;
        best_gt[i] = arg;
    }
}

__global__ void map_gt_count_kernel(const double* __restrict__ gt, const long long* __restrict__ gt_off,
                                    long long n_img, long long gt_cap, int C, int* __restrict__ counts) {
    long long n_gt = gt_off[n_img];
    if (n_gt > gt_cap) n_gt = gt_cap;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_gt;
         g += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = n_img;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (gt_off[mid] <= g) lo = mid; else hi = mid;
        }
        const double cf = gt[g * 7 + 5];
        const long long cls = (long long)cf;
        if (cls >= 0 && cls < C) atomicAdd(&counts[lo * C + cls], 1);
    }
}

}  // namespace yb

using namespace yb;

extern "C" int yb_map_match(const double* gt_rows, const int64_t* gt_offsets, const double* det_rows,
                            const int64_t* det_offsets, int64_t n_img, int class_num, int64_t gt_cap,
                            int64_t det_cap, double* best_iou, int32_t* best_gt, int32_t* gt_class_counts,
                            yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (gt_offsets == nullptr || det_offsets == nullptr || gt_class_counts == nullptr) return YB_E_NULL;
    if (n_img < 0 || class_num <= 0 || gt_cap < 0 || det_cap < 0) return YB_E_SHAPE;
    if (n_img == 0) return YB_OK;
    if ((gt_cap > 0 && gt_rows == nullptr) ||
        (det_cap > 0 && (det_rows == nullptr || best_iou == nullptr || best_gt == nullptr)))
        return YB_E_NULL;
    YB_CUDA_TRY(cudaMemsetAsync(gt_class_counts, 0, sizeof(int32_t) * n_img * class_num, stream));
    const int threads = 256;
    if (gt_cap > 0) {
        const int blocks = (int)min((long long)kNumSMs * 8, ((long long)gt_cap + threads - 1) / threads);
        map_gt_count_kernel<<<blocks, threads, 0, stream>>>(
            gt_rows, reinterpret_cast<const long long*>(gt_offsets), n_img, gt_cap, class_num, gt_class_counts);
        YB_CUDA_TRY(cudaGetLastError());
    }
    if (det_cap > 0) {
        const int blocks = (int)min((long long)kNumSMs * 8, ((long long)det_cap + threads - 1) / threads);
        map_match_kernel<<<blocks, threads, 0, stream>>>(
            gt_rows, reinterpret_cast<const long long*>(gt_offsets), det_rows,
            reinterpret_cast<const long long*>(det_offsets), n_img, det_cap, best_iou, best_gt);
        YB_CUDA_TRY(cudaGetLastError());
    }
    return YB_OK;
}
