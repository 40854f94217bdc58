// PR-curve / mAP matching: per (image, class) every detection takes the maximum
// and arg-maximum IoU over the ground truths of its class
// (utils/measurement.py:252-283 PRfunc, :104-130 create_score_mat), float64 in the
// reference's operation order (utils/tools.py:649-666; -fmad=false).
#include "common.cuh"
#include "nms_pair.cuh"
#include "scan.cuh"

namespace yb {

// the IoU of utils/tools.py:649-666 exactly as the NMS kernels evaluate it: np.maximum / np.minimum
// PROPAGATE NaN (fmax / fmin would drop it), so a non-finite box yields the NaN NumPy yields
__device__ __forceinline__ double map_iou(double tx, double ty, double tw, double th, double px,
                                          double py, double pw, double ph) {
    return pair_iou<1>(tx, ty, tw, th, px, py, pw, ph);
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
        bool nan_seen = false;
        for (long long g = gt_off[lo]; g < gt_off[lo + 1]; ++g) {
            const double* t = gt + g * 7;
            if ((long long)t[5] != cls) continue;
            const double v = map_iou(t[0], t[1], t[2], t[3], d[0], d[1], d[2], d[3]);
            if (v != v) {               // np.max -> NaN, np.argmax -> the FIRST NaN
                if (!nan_seen) {
                    nan_seen = true;
                    best = v;
                    arg = seen;
                }
            } else if (!nan_seen && (arg < 0 || v > best)) {  // first maximum, like np.argmax
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

// ============================================================================
// PR accumulation: detections -> (conf, gt_id, tp flag) triples grouped by class
// (utils/measurement.py:252-292), per-class score-matrix counters (:104-130), and the
// final per-class confidence sort + distinct-TP prefix (:297-321).
// ============================================================================
namespace yb {

// ---- A: per (image, class) segment sizes, transposed to class-major ----------
__global__ void map_segment_sizes_kernel(const long long* __restrict__ seg_off, const int* __restrict__ gt_counts,
                                         long long n_img, int C, long long max_per_img,
                                         unsigned int* __restrict__ kept_T, unsigned int* __restrict__ gt_T) {
    const long long n_seg = n_img * C;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_seg;
         s += (long long)gridDim.x * blockDim.x) {
        const long long img = s / C;
        const int c = (int)(s - img * C);
        long long n = seg_off[s + 1] - seg_off[s];
        if (max_per_img > 0 && n > max_per_img) n = max_per_img;
        kept_T[(long long)c * n_img + img] = (unsigned)n;
        gt_T[(long long)c * n_img + img] = (unsigned)gt_counts[s];
    }
}

constexpr int kTopCap = 1024;   // truncated segments up to this size are ranked by a warp-wide sort

__device__ __forceinline__ unsigned long long orderable(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);  // ascending with value
}

// ---- B: one warp per segment: triples, per-image top-k, score counters ----------
__global__ void __launch_bounds__(256)
map_triples_kernel(const double* __restrict__ det, const long long* __restrict__ seg_off,
                   const double* __restrict__ best_iou, const int* __restrict__ best_gt,
                   const int* __restrict__ gt_counts, long long n_img, int C, double iou_thr,
                   long long max_per_img, const long long* __restrict__ gt_base,
                   const long long* __restrict__ out_pos_T, const long long* __restrict__ gt_pos_T,
                   double* __restrict__ conf_out, long long* __restrict__ gtid_out,
                   unsigned char* __restrict__ flag_out, int* __restrict__ cls_out,
                   long long* __restrict__ class_offsets, unsigned long long* __restrict__ acc /* [3][C] pp,tpp,tp */) {
    // per-warp sort scratch for the top-k of a truncated segment: kTopCap keys + indices
    extern __shared__ __align__(16) unsigned char topk_smem[];
    const int lane = threadIdx.x & 31;
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(topk_smem) + (size_t)(threadIdx.x >> 5) * kTopCap;
    unsigned short* s_idx = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned long long*>(topk_smem) +
                                                              (size_t)(blockDim.x >> 5) * kTopCap) +
                            (size_t)(threadIdx.x >> 5) * kTopCap;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_seg = n_img * C;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= C; c += (long long)gridDim.x * blockDim.x)
        class_offsets[c] = out_pos_T[c * n_img];  // c == C -> total
    for (long long s = warp; s < n_seg; s += n_warps) {
        const long long img = s / C;
        const int c = (int)(s - img * C);
        const long long d0 = seg_off[s];
        const int n = (int)(seg_off[s + 1] - d0);
        if (n == 0) continue;
        const int n_gt = gt_counts[s];
        const long long tpos = (long long)c * n_img + img;
        const long long out0 = out_pos_T[tpos];
        // running ground-truth count of this class before this image (measurement.py:256-258)
        const long long gbase = gt_base[c] + (gt_pos_T[tpos] - gt_pos_T[(long long)c * n_img]);
        const bool trunc = max_per_img > 0 && n > max_per_img;
        const bool sorted = trunc && n <= kTopCap;
        if (sorted) {
            // visit order of np.argsort(conf)[::-1][:max_per_img] (measurement.py:285-289): a warp-wide
            // bitonic sort of (descending confidence, higher index first) in shared memory gives every
            // detection its rank; the ranks below max_per_img are the output positions
            int P = 32;
            while (P < n) P <<= 1;
            for (int j = lane; j < P; j += 32) {
                unsigned long long k = ~0ull;   // padding sorts last
                if (j < n) {
                    const double* r = det + (d0 + j) * 7;
                    k = ~orderable(__dmul_rn(r[4], r[6]));   // ascending key == descending confidence
                }
                s_key[j] = k;
                s_idx[j] = (unsigned short)j;
            }
            __syncwarp();
            for (int k = 2; k <= P; k <<= 1) {
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
                    for (int t = lane; t < P; t += 32) {
                        const int u = t ^ jj;
                        if (u > t) {
                            const unsigned long long ka = s_key[t], kb = s_key[u];
                            const unsigned short ia = s_idx[t], ib = s_idx[u];
                            // a before b: smaller key, ties -> higher index first
                            const bool a_first = ka < kb || (ka == kb && ia > ib);
                            const bool up = (t & k) == 0;
                            if (up ? !a_first : a_first) {
                                s_key[t] = kb; s_key[u] = ka;
                                s_idx[t] = ib; s_idx[u] = ia;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            // rank -> position lookup by original index (the key array is dead now)
            unsigned short* s_rank = reinterpret_cast<unsigned short*>(s_key);
            __syncwarp();
            unsigned short mine[kTopCap / 32];
#pragma unroll
            for (int q = 0; q < kTopCap / 32; ++q) mine[q] = (lane + 32 * q < n) ? s_idx[lane + 32 * q] : (unsigned short)0;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < kTopCap / 32; ++q)
                if (lane + 32 * q < n) s_rank[mine[q]] = (unsigned short)(lane + 32 * q);
            __syncwarp();
        }
        int tpp = 0, tp = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            double cf = 0.0;
            long long gid = 0;
            bool fl = false;
            if (j < n) {
                const double* r = det + (d0 + j) * 7;
                cf = __dmul_rn(r[4], r[6]);
                if (n_gt > 0) {
                    fl = best_iou[d0 + j] >= iou_thr;
                    gid = (long long)best_gt[d0 + j] + gbase;
                }
            }
            // score-matrix counters: flagged detections and distinct matched ground truths
            bool first = fl;
            if (fl) {
                const int g = best_gt[d0 + j];
                for (int i = 0; i < j; ++i)
                    if (best_iou[d0 + i] >= iou_thr && best_gt[d0 + i] == g) { first = false; break; }
            }
            tpp += __popc(__ballot_sync(0xffffffffu, fl));
            tp += __popc(__ballot_sync(0xffffffffu, first));
            if (j < n) {
                long long dst;
                if (!trunc) {
                    dst = out0 + j;
                } else if (sorted) {
                    const int rank = (int)reinterpret_cast<const unsigned short*>(s_key)[j];
                    dst = (rank < max_per_img) ? out0 + rank : -1;
                } else {  // top max_per_img by confidence, in sorted order, ties: higher index first
                    int rank = 0;
                    for (int i = 0; i < n; ++i) {
                        const double* q = det + (d0 + i) * 7;
                        const double ci = __dmul_rn(q[4], q[6]);
                        rank += (ci > cf || (ci == cf && i > j)) ? 1 : 0;
                    }
                    dst = (rank < max_per_img) ? out0 + rank : -1;
                }
                if (dst >= 0) {
                    conf_out[dst] = cf;
                    gtid_out[dst] = gid;
                    flag_out[dst] = fl ? 1 : 0;
                    cls_out[dst] = c;
                }
            }
        }
        __syncwarp();   // the sort scratch is reused by the warp's next segment
        if (lane == 0) {
            atomicAdd(&acc[c], (unsigned long long)n);
            if (n_gt > 0) {
                atomicAdd(&acc[C + c], (unsigned long long)tpp);
                atomicAdd(&acc[2 * C + c], (unsigned long long)tp);
            }
        }
    }
}

// ---- C: sort records (class asc, confidence desc, position desc) -----------------------
// record = two u64 compared lexicographically: A = class<<32 | hi32(~key), B = lo32(~key)<<32 | ~pos
__device__ __forceinline__ bool rec_less(const ulonglong2& a, const ulonglong2& b) {
    return a.x < b.x || (a.x == b.x && a.y < b.y);
}

__global__ void pr_pack_kernel(const double* __restrict__ conf, const int* __restrict__ cls, long long n,
                               long long P, ulonglong2* __restrict__ rec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P;
         i += (long long)gridDim.x * blockDim.x) {
        ulonglong2 r;
        if (i < n) {
            const unsigned long long k = ~orderable(conf[i]);  // ascending k == descending confidence
            r.x = ((unsigned long long)(unsigned)cls[i] << 32) | (k >> 32);
            r.y = (k << 32) | (unsigned long long)(~(unsigned)i);
        } else {
            r.x = ~0ull;  // padding sorts last
            r.y = ~0ull;
        }
        rec[i] = r;
    }
}

constexpr int kSortTile = 2048;  // records per CTA in the shared-memory stages (32 KB)

// all (k, j) stages with j < kSortTile for k <= k_hi, starting from stage (k_lo, j_lo)
__global__ void __launch_bounds__(512)
bitonic_smem_kernel(ulonglong2* __restrict__ rec, long long k_lo, long long k_hi, long long j_start) {
    __shared__ ulonglong2 s[kSortTile];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) s[t] = rec[base + t];
    __syncthreads();
    for (long long k = k_lo; k <= k_hi; k <<= 1) {
        long long j = (k == k_lo) ? j_start : (k >> 1);
        if (j >= kSortTile) j = kSortTile >> 1;
        for (; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) {
                const int u = t ^ (int)j;
                if (u > t) {
                    const bool up = ((base + t) & k) == 0;
                    const ulonglong2 a = s[t], b = s[u];
                    if (up ? rec_less(b, a) : rec_less(a, b)) {
                        s[t] = b;
                        s[u] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) rec[base + t] = s[t];
}

__global__ void bitonic_global_kernel(ulonglong2* __restrict__ rec, long long P, long long k, long long j) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P;
         t += (long long)gridDim.x * blockDim.x) {
        const long long u = t ^ j;
        if (u > t) {
            const bool up = (t & k) == 0;
            const ulonglong2 a = rec[t], b = rec[u];
            if (up ? rec_less(b, a) : rec_less(a, b)) {
                rec[t] = b;
                rec[u] = a;
            }
        }
    }
}

// ---- D: first occurrence of every matched ground truth in sorted order ----------------
__global__ void pr_first_pos_kernel(const ulonglong2* __restrict__ rec, long long n,
                                    const long long* __restrict__ gtid, const unsigned char* __restrict__ flag,
                                    const long long* __restrict__ gt_table_base,
                                    unsigned long long* __restrict__ first_pos, long long* __restrict__ order) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 r = rec[i];
        const long long src = (long long)(unsigned)(~(unsigned)(r.y & 0xffffffffull));
        order[i] = src;
        if (flag[src]) {
            const int c = (int)(r.x >> 32);
            atomicMin(&first_pos[gt_table_base[c] + gtid[src]], (unsigned long long)i);
        }
    }
}

__global__ void pr_flags_kernel(const ulonglong2* __restrict__ rec, long long n, const long long* __restrict__ order,
                                const long long* __restrict__ gtid, const unsigned char* __restrict__ flag,
                                const long long* __restrict__ gt_table_base,
                                const unsigned long long* __restrict__ first_pos,
                                unsigned int* __restrict__ is_first, unsigned int* __restrict__ is_flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long src = order[i];
        const bool fl = flag[src] != 0;
        const int c = (int)(rec[i].x >> 32);
        is_flag[i] = fl ? 1u : 0u;
        is_first[i] = (fl && first_pos[gt_table_base[c] + gtid[src]] == (unsigned long long)i) ? 1u : 0u;
    }
}

// Append the valid prefix of one chunk's records (count read on the device) to the rank's record
// arrays at the running total (read and advanced on the device): no host round trip per chunk.
__global__ void map_append_kernel(const double* __restrict__ conf, const long long* __restrict__ gid,
                                  const unsigned char* __restrict__ flag, const int* __restrict__ cls,
                                  const long long* __restrict__ n_src, double* __restrict__ conf_dst,
                                  long long* __restrict__ gid_dst, unsigned char* __restrict__ flag_dst,
                                  int* __restrict__ cls_dst, long long dst_cap, long long* __restrict__ total,
                                  unsigned int* __restrict__ done) {
    const long long n = *n_src, base = *total;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long d = base + i;
        if (d < dst_cap) {
            conf_dst[d] = conf[i];
            gid_dst[d] = gid[i];
            flag_dst[d] = flag[i];
            cls_dst[d] = cls[i];
        }
    }
    // the last CTA to finish advances the total (every CTA has read it by then)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *total = base + n;   // may exceed dst_cap: the caller checks and regrows
        *done = 0u;
    }
}

// precision / recall of every prefix of every class (utils/measurement.py:302-319), from the running
// counts of yb_pr_curve: record i of the sorted array belongs to the class whose extent
// [class_start[c], class_start[c+1]) contains it; num_dets = i - start + 1, num_tp / num_tpp =
// counts since the class start; precision by mode (0: tpp/dets, 1: tp/(tp+fp), 2: tp/dets),
// recall = tp / gts - int64 / int64 true divisions, as NumPy does them (exact conversions below 2^53).
__global__ void pr_points_kernel(const long long* __restrict__ tp_cum, const long long* __restrict__ tpp_cum,
                                 const long long* __restrict__ class_start, const long long* __restrict__ gts,
                                 int C, long long n, int mode, double* __restrict__ precision,
                                 double* __restrict__ recall) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = C;   // class_start[lo] <= i < class_start[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (class_start[mid] <= i) lo = mid; else hi = mid;
        }
        const long long a = class_start[lo];
        const long long tp = tp_cum[i + 1] - tp_cum[a], tpp = tpp_cum[i + 1] - tpp_cum[a];
        const long long dets = i - a + 1, fp = dets - tpp;
        double p;
        if (mode == 0) p = (double)tpp / (double)dets;
        else if (mode == 1) p = (double)tp / (double)(tp + fp);
        else p = (double)tp / (double)dets;
        precision[i] = p;
        recall[i] = (double)tp / (double)gts[lo];
    }
}

static long long next_pow2(long long n) {
    long long p = kSortTile;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace yb

extern "C" size_t yb_map_accumulate_workspace_bytes(int64_t n_img, int class_num) {
    if (n_img < 0 || class_num <= 0) return 0;
    const long long n = (long long)n_img * class_num;
    return 4 * align_up((size_t)(n + 2) * sizeof(long long), 256) + scan_workspace_bytes(n + 1) + 1024;
}

extern "C" int yb_map_accumulate(const double* det_rows, const int64_t* det_seg_offsets, const double* best_iou,
                                 const int32_t* best_gt, const int32_t* gt_class_counts, int64_t n_img,
                                 int class_num, double iou_threshold, int64_t max_per_img,
                                 const int64_t* gt_base, double* conf, int64_t* gt_id, uint8_t* flag,
                                 int32_t* cls, int64_t* class_offsets, uint64_t* score_acc, void* workspace,
                                 size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (det_seg_offsets == nullptr || gt_class_counts == nullptr || gt_base == nullptr ||
        class_offsets == nullptr || score_acc == nullptr || workspace == nullptr)
        return YB_E_NULL;
    if (n_img <= 0 || class_num <= 0) return YB_E_SHAPE;
    if (workspace_bytes < yb_map_accumulate_workspace_bytes(n_img, class_num) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    const long long n = (long long)n_img * class_num;
    const size_t slab = align_up((size_t)(n + 2) * sizeof(long long), 256);
    char* w = reinterpret_cast<char*>(workspace);
    unsigned int* kept_T = reinterpret_cast<unsigned int*>(w);
    unsigned int* gt_T = reinterpret_cast<unsigned int*>(w + slab);
    long long* out_pos_T = reinterpret_cast<long long*>(w + 2 * slab);
    long long* gt_pos_T = reinterpret_cast<long long*>(w + 3 * slab);
    void* scan_ws = w + 4 * slab;
    const int threads = 256;
    const int blocks = (int)min((long long)kNumSMs * 8, (n + threads - 1) / threads);
    map_segment_sizes_kernel<<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const long long*>(det_seg_offsets), gt_class_counts, n_img, class_num, max_per_img, kept_T, gt_T);
    YB_CUDA_TRY(cudaGetLastError());
    int rc = exclusive_scan_u32(kept_T, n, out_pos_T, scan_ws, stream);
    if (rc != 0) return rc;
    rc = exclusive_scan_u32(gt_T, n, gt_pos_T, scan_ws, stream);
    if (rc != 0) return rc;
    const int wblocks = (int)min((long long)kNumSMs * 8, (n * 32 + threads - 1) / threads);
    const size_t topk_smem = (size_t)(threads / 32) * kTopCap * (sizeof(unsigned long long) + sizeof(unsigned short));
    static SmemRaised done;
    YB_CUDA_TRY(raise_dynamic_smem_once(map_triples_kernel, (int)topk_smem, &done));
    map_triples_kernel<<<wblocks, threads, topk_smem, stream>>>(
        det_rows, reinterpret_cast<const long long*>(det_seg_offsets), best_iou, best_gt, gt_class_counts, n_img,
        class_num, iou_threshold, max_per_img, reinterpret_cast<const long long*>(gt_base), out_pos_T, gt_pos_T,
        conf, reinterpret_cast<long long*>(gt_id), flag, cls, reinterpret_cast<long long*>(class_offsets),
        reinterpret_cast<unsigned long long*>(score_acc));
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" int yb_map_append(const double* conf, const int64_t* gt_id, const uint8_t* flag, const int32_t* cls,
                             const int64_t* n_src, int64_t src_capacity, double* conf_dst, int64_t* gt_id_dst,
                             uint8_t* flag_dst, int32_t* cls_dst, int64_t dst_capacity, int64_t* total,
                             uint32_t* counter, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_src == nullptr || total == nullptr || counter == nullptr) return YB_E_NULL;
    if (src_capacity < 0 || dst_capacity < 0) return YB_E_SHAPE;
    if (src_capacity > 0 && (conf == nullptr || gt_id == nullptr || flag == nullptr || cls == nullptr)) return YB_E_NULL;
    if (dst_capacity > 0 && (conf_dst == nullptr || gt_id_dst == nullptr || flag_dst == nullptr || cls_dst == nullptr))
        return YB_E_NULL;
    const int threads = 256;
    const int blocks = (int)max(1LL, min((long long)kNumSMs * 4, ((long long)src_capacity + threads - 1) / threads));
    map_append_kernel<<<blocks, threads, 0, stream>>>(
        conf, reinterpret_cast<const long long*>(gt_id), flag, cls, reinterpret_cast<const long long*>(n_src), conf_dst,
        reinterpret_cast<long long*>(gt_id_dst), flag_dst, cls_dst, dst_capacity, reinterpret_cast<long long*>(total),
        counter);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" int yb_pr_points(const int64_t* tp_cum, const int64_t* tpp_cum, const int64_t* class_start,
                            const int64_t* gts, int class_num, int64_t n_det, int precision_mode,
                            double* precision, double* recall, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_det < 0 || class_num < 1) return YB_E_SHAPE;
    if (precision_mode < 0 || precision_mode > 2) return YB_E_PARAM;
    if (n_det == 0) return YB_OK;
    if (tp_cum == nullptr || tpp_cum == nullptr || class_start == nullptr || gts == nullptr || precision == nullptr ||
        recall == nullptr)
        return YB_E_NULL;
    const int threads = 256;
    const int blocks = (int)min((long long)kNumSMs * 8, ((long long)n_det + threads - 1) / threads);
    pr_points_kernel<<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const long long*>(tp_cum), reinterpret_cast<const long long*>(tpp_cum),
        reinterpret_cast<const long long*>(class_start), reinterpret_cast<const long long*>(gts), class_num, n_det,
        precision_mode, precision, recall);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" size_t yb_pr_curve_workspace_bytes(int64_t n_det, int64_t n_gt_total) {
    if (n_det < 0 || n_gt_total < 0) return 0;
    const long long P = next_pow2(n_det > 0 ? n_det : 1);
    return align_up((size_t)P * sizeof(ulonglong2), 256) + align_up((size_t)(n_gt_total + 1) * 8, 256) +
           2 * align_up((size_t)(n_det + 1) * 4, 256) + scan_workspace_bytes(n_det + 1) + 1024;
}

// order[i] = index (into conf/gt_id/flag/cls) of the i-th record after sorting by
// (class asc, confidence desc, position desc); tp_cum / tpp_cum have n_det+1 entries:
// exclusive running counts of first-occurrence true positives / flagged detections over the
// WHOLE sorted array (the caller subtracts the value at each class start).
extern "C" int yb_pr_curve(const double* conf, const int32_t* cls, const int64_t* gt_id, const uint8_t* flag,
                           int64_t n_det, const int64_t* gt_table_base, int64_t n_gt_total, int64_t* order,
                           int64_t* tp_cum, int64_t* tpp_cum, void* workspace, size_t workspace_bytes,
                           yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n_det < 0 || n_gt_total < 0) return YB_E_SHAPE;
    if (tp_cum == nullptr || tpp_cum == nullptr || workspace == nullptr) return YB_E_NULL;
    if (n_det == 0) {
        YB_CUDA_TRY(cudaMemsetAsync(tp_cum, 0, 8, stream));
        YB_CUDA_TRY(cudaMemsetAsync(tpp_cum, 0, 8, stream));
        return YB_OK;
    }
    if (conf == nullptr || cls == nullptr || gt_id == nullptr || flag == nullptr || gt_table_base == nullptr ||
        order == nullptr)
        return YB_E_NULL;
    if (n_det > 0xfffffff0ll) return YB_E_SHAPE;
    if (workspace_bytes < yb_pr_curve_workspace_bytes(n_det, n_gt_total) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    const long long P = next_pow2(n_det);
    char* w = reinterpret_cast<char*>(workspace);
    ulonglong2* rec = reinterpret_cast<ulonglong2*>(w);
    w += align_up((size_t)P * sizeof(ulonglong2), 256);
    unsigned long long* first_pos = reinterpret_cast<unsigned long long*>(w);
    w += align_up((size_t)(n_gt_total + 1) * 8, 256);
    unsigned int* is_first = reinterpret_cast<unsigned int*>(w);
    w += align_up((size_t)(n_det + 1) * 4, 256);
    unsigned int* is_flag = reinterpret_cast<unsigned int*>(w);
    w += align_up((size_t)(n_det + 1) * 4, 256);
    void* scan_ws = w;

    const int threads = 256;
    const int gblocks = (int)min((long long)kNumSMs * 8, (P + threads - 1) / threads);
    pr_pack_kernel<<<gblocks, threads, 0, stream>>>(conf, cls, n_det, P, rec);
    YB_CUDA_TRY(cudaGetLastError());
    const int tiles = (int)(P / kSortTile);
    bitonic_smem_kernel<<<tiles, 512, 0, stream>>>(rec, 2, kSortTile, 1);  // sorts every tile
    YB_CUDA_TRY(cudaGetLastError());
    for (long long k = 2 * kSortTile; k <= P; k <<= 1) {
        for (long long j = k >> 1; j >= kSortTile; j >>= 1)
            bitonic_global_kernel<<<gblocks, threads, 0, stream>>>(rec, P, k, j);
        bitonic_smem_kernel<<<tiles, 512, 0, stream>>>(rec, k, k, kSortTile >> 1);
        YB_CUDA_TRY(cudaGetLastError());
    }
    YB_CUDA_TRY(cudaMemsetAsync(first_pos, 0xff, (size_t)(n_gt_total + 1) * 8, stream));
    const int dblocks = (int)min((long long)kNumSMs * 8, ((long long)n_det + threads - 1) / threads);
    pr_first_pos_kernel<<<dblocks, threads, 0, stream>>>(rec, n_det, reinterpret_cast<const long long*>(gt_id), flag,
                                                         reinterpret_cast<const long long*>(gt_table_base), first_pos,
                                                         reinterpret_cast<long long*>(order));
    pr_flags_kernel<<<dblocks, threads, 0, stream>>>(rec, n_det, reinterpret_cast<const long long*>(order),
                                                     reinterpret_cast<const long long*>(gt_id), flag,
                                                     reinterpret_cast<const long long*>(gt_table_base), first_pos,
                                                     is_first, is_flag);
    YB_CUDA_TRY(cudaGetLastError());
    int rc = exclusive_scan_u32(is_first, n_det, reinterpret_cast<long long*>(tp_cum), scan_ws, stream);
    if (rc != 0) return rc;
    return exclusive_scan_u32(is_flag, n_det, reinterpret_cast<long long*>(tpp_cum), scan_ws, stream);
}
