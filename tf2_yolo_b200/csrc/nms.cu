// Per-class greedy NMS / DIoU-NMS, batched over images, bit-exact against the
// reference's float64 arithmetic (utils/tools.py:630-684 cal_iou, :687-733 nms).
// Compiled with -fmad=false: NumPy rounds every multiply/add/subtract separately.
//
// Pipeline (all on the caller's stream, no host round trip; the true row count is
// read on the device from row_offsets[n_img]):
//   classify  : row -> segment (image, class); per-segment counts (atomics)
//   scatter   : per image: scan of its C segment counts (start offsets), row ids grouped by
//               segment, segments binned into work lists
//   sweep     : ONE launch.  The first CTAs take the big segments (> 32 boxes, one CTA each):
//               boxes ordered (rank counting up to 128 boxes, bitonic sorts above), corners
//               and areas staged once in shared memory in visit order, blocked greedy sweep
//               (64x64 suppression bitmask per block, serial resolve of the block, the
//               block's kept boxes against every later box).  All other CTAs work through
//               the warp lists: segments of <= 8 boxes four to a warp (8-lane groups),
//               9..32 boxes one warp each - rank by confidence with shuffles, suppression
//               masks from group-broadcast boxes, mask sweep in registers.
//               The pair test is division-free (see suppresses_fast).
//   emit      : per image: prefix of the survivor counts of earlier images + scan of its own C
//               segments, survivors written in the reference's order (class-major, original
//               order inside a class)
#include <climits>
#include <cstring>

#include "common.cuh"
#include "nms_pair.cuh"
#include "scan.cuh"

namespace yb {

constexpr int kBigThreads = 512;   // 2 CTAs/SM: 32 warps hide the fp64 latency of the sweeps
constexpr int kMaskParts = kBigThreads / 64;   // threads per row of the 64x64 block mask
constexpr int kBigCap = 1536;      // boxes per segment held in shared memory
constexpr int kBigP = 2048;        // power-of-two padding of the index arrays of such a segment
constexpr int kSweep = 64;         // sweep block (one 64-bit mask word per box)
constexpr int kCountSort = 128;    // segments up to this size are ordered by rank counting, not bitonic sorts
constexpr int kTiny = 8;           // segments up to this size share a warp (one 8-lane group each)

struct NmsWs {
    int* row_seg;            // [R]
    unsigned int* seg_count; // [n_seg + 1]
    unsigned int* seg_fill;  // [n_seg]
    unsigned int* seg_kept;  // [n_seg + 1]
    unsigned int* img_kept;  // [n_img + 1] survivors per image (atomics from the sweeps)
    unsigned int* ctrl;      // [8]: n_small (9..32 boxes), n_big, n_tiny (<= 8 boxes)
    long long* seg_start;    // [n_seg + 2]
    long long* out_start;    // [n_seg + 2]
    int* members;            // [R]
    int* local_rank;         // [R] by ROW id: rank among its segment's survivors, -1 = removed
    int* small_list;         // [n_seg]: tiny segments from the front, one-warp segments from the back
    int* big_list;           // [n_seg]
    double* gbox;            // [7][R]: planes X(2) Y(2) C(2) A(1) of segments > kBigCap
    int* gmem_pad;           // [2R] padded handle array
    int* gord_pad;           // [2R]
    unsigned char* gremoved; // [R]
};

__device__ __forceinline__ long long device_row_count(const long long* row_offsets, long long n_img,
                                                      long long cap) {
    const long long n = row_offsets[n_img];
    return n < cap ? n : cap;
}

__global__ void nms_classify_kernel(const double* __restrict__ rows, const long long* __restrict__ row_offsets,
                                    long long n_img, long long cap, int C, NmsWs W,
                                    unsigned char* __restrict__ keep) {
    const long long n = device_row_count(row_offsets, n_img, cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        // image of row i: last img with row_offsets[img] <= i
        long long lo = 0, hi = n_img;  // invariant: row_offsets[lo] <= i < row_offsets[hi]
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (row_offsets[mid] <= i) lo = mid; else hi = mid;
        }
        const double cf = rows[i * 7 + 5];
        const long long cls = (long long)cf;  // astype("int"): truncation toward zero
        int seg = -1;
        if (cls >= 0 && cls < C && cf == cf) {
            seg = (int)(lo * C + cls);
            atomicAdd(&W.seg_count[seg], 1u);
        }
        W.row_seg[i] = seg;
        keep[i] = 0;
    }
}

// grid (n_img, Y): every CTA of image i scans the image's C segment counts (start offsets are
// row_offsets[i] + exclusive scan: no device-wide scan, rows of an image stay inside its row range),
// then scatters its share of the image's rows; the y == 0 CTA also bins the segments.
__global__ void __launch_bounds__(256)
nms_scatter_kernel(const long long* __restrict__ row_offsets, long long n_img, long long cap, int C, NmsWs W) {
    __shared__ long long s_warp[32];
    const long long n = device_row_count(row_offsets, n_img, cap);
    const long long img = blockIdx.x;
    const long long r0 = min(row_offsets[img], n), r1 = min(row_offsets[img + 1], n);
    const long long seg0 = img * C;
    const int tid = threadIdx.x;
    long long carry = 0;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + tid;
        const unsigned cnt = (c < C) ? W.seg_count[seg0 + c] : 0u;
        long long total;
        const long long ex = block_exclusive_scan((long long)cnt, s_warp, total);
        if (c < C) {
            W.seg_start[seg0 + c] = r0 + carry + ex;   // same value from every CTA of the image
            if (blockIdx.y == 0) {
                const long long sg = seg0 + c;
                if (cnt == 0u) {
                    W.seg_kept[sg] = 0u;
                } else if (cnt <= (unsigned)kTiny) {   // tiny segments fill the list from the front ...
                    W.small_list[atomicAdd(&W.ctrl[2], 1u)] = (int)sg;
                } else if (cnt <= 32u) {               // ... one-warp segments from the back
                    W.small_list[n_img * C - 1 - (long long)atomicAdd(&W.ctrl[0], 1u)] = (int)sg;
                } else {
                    W.big_list[atomicAdd(&W.ctrl[1], 1u)] = (int)sg;
                }
            }
        }
        carry += total;
    }
    __syncthreads();   // this CTA's seg_start writes are visible to its own threads
    for (long long row = r0 + (long long)blockIdx.y * blockDim.x + tid; row < r1; row += (long long)gridDim.y * blockDim.x) {
        const int seg = W.row_seg[row];
        if (seg >= 0) {
            const long long pos = W.seg_start[seg] + atomicAdd(&W.seg_fill[seg], 1u);
            W.members[pos] = (int)row;
        }
    }
}

// ---- small segments: one warp each ---------------------------------------------
template <int GW>
__device__ __forceinline__ BoxC shfl_box(const BoxC& b, int src, bool with_centre) {
    BoxC r;
    r.x0 = __shfl_sync(0xffffffffu, b.x0, src, GW); r.x1 = __shfl_sync(0xffffffffu, b.x1, src, GW);
    r.y0 = __shfl_sync(0xffffffffu, b.y0, src, GW); r.y1 = __shfl_sync(0xffffffffu, b.y1, src, GW);
    r.area = __shfl_sync(0xffffffffu, b.area, src, GW);
    r.cx = with_centre ? __shfl_sync(0xffffffffu, b.cx, src, GW) : 0.0;
    r.cy = with_centre ? __shfl_sync(0xffffffffu, b.cy, src, GW) : 0.0;
    return r;
}

// One segment per GROUP of GW lanes (GW = 32: one per warp; GW = 8: four tiny segments share a
// warp).  seg < 0 = idle group.  Every lane of the warp runs every shuffle (loops go to the
// largest segment of the warp), so the full-mask shuffles stay converged.
template <int MODE, int GW>
__device__ __forceinline__ void nms_group_segment(const double* __restrict__ rows, double thr, double conf_thr,
                                                  double sigma, const NmsWs& W, int seg, bool pos_thr, int C,
                                                  unsigned char* __restrict__ keep) {
    constexpr unsigned GM = (GW == 32) ? 0xffffffffu : ((1u << GW) - 1u);
    const int lane = threadIdx.x & 31;
    const int gl = lane & (GW - 1);        // lane inside the group
    const int gshift = lane & ~(GW - 1);   // first lane of the group
    const int n = (seg >= 0) ? (int)W.seg_count[seg] : 0;
    const long long start = (seg >= 0) ? W.seg_start[seg] : 0;
    int nmax = n;
    if (GW < 32) {
#pragma unroll
        for (int o = GW; o < 32; o <<= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    }
    const unsigned valid = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);

    // restore original order: rank the (atomically scattered) row ids, permute by shuffle
    int m = (gl < n) ? W.members[start + gl] : INT_MAX;
    {
        int rk = 0;
        for (int j = 0; j < nmax; ++j) {
            const int mj = __shfl_sync(0xffffffffu, m, j, GW);   // every lane shuffles, whatever its n
            rk += (j < n && mj < m) ? 1 : 0;
        }
        int src = gl;
        for (int j = 0; j < nmax; ++j) {
            const int rj = __shfl_sync(0xffffffffu, rk, j, GW);
            if (rj == gl && j < n) src = j;
        }
        m = __shfl_sync(0xffffffffu, m, src, GW);
        if (gl < n) W.members[start + gl] = m;
    }

    double x = 0, y = 0, bw = 0, bh = 0, conf = 0;
    if (gl < n) {
        const double* r = rows + (long long)m * 7;
        x = r[0]; y = r[1]; bw = r[2]; bh = r[3];
        conf = __dmul_rn(r[4], r[6]);
    }
    const BoxC mine = make_box(x, y, bw, bh);
    // visit rank: descending confidence, ties -> higher original index first
    int vis = 0;
    unsigned sup = 0;
    for (int j = 0; j < nmax; ++j) {
        const double cj = __shfl_sync(0xffffffffu, conf, j, GW);
        vis += (j < n && j != gl && visited_before(cj, j, conf, gl)) ? 1 : 0;
        if (MODE != 3) {
            const BoxC other = shfl_box<GW>(mine, j, MODE == 2);
            const int mj = __shfl_sync(0xffffffffu, m, j, GW);
            if (gl < n && j < n && suppresses<(MODE == 3 ? 1 : MODE)>(mine, other, thr, pos_thr, rows, m, mj))
                sup |= 1u << j;
        }
    }
    unsigned dead = 0, seen = 0;
    if (MODE == 3) {
        // soft-NMS (utils/tools.py:736-786): the visit order is fixed up front and deleted boxes keep
        // decaying others, so box j's final confidence is its own product over the boxes visited
        // before it, taken in visit order - independent of every other box.
        bool decayed = false;
        for (int v = 0; v < nmax; ++v) {
            const unsigned who = (__ballot_sync(0xffffffffu, gl < n && vis == v) >> gshift) & GM;
            const int src = who ? (__ffs(who) - 1) : 0;
            const double xs = __shfl_sync(0xffffffffu, x, src, GW), ys = __shfl_sync(0xffffffffu, y, src, GW);
            const double ws = __shfl_sync(0xffffffffu, bw, src, GW), hs = __shfl_sync(0xffffffffu, bh, src, GW);
            if (who != 0u && gl < n && vis > v) {
                const double iou = pair_iou<1>(xs, ys, ws, hs, x, y, bw, bh);
                if (iou >= thr) {
                    conf = conf * exp(-1.0 * (iou * iou) / sigma);
                    decayed = true;
                }
            }
        }
        dead = (__ballot_sync(0xffffffffu, gl < n && decayed && conf < conf_thr) >> gshift) & GM;
    } else {
        // greedy sweep, state replicated in every lane of the group
        for (int v = 0; v < nmax; ++v) {
            const unsigned who = (__ballot_sync(0xffffffffu, gl < n && vis == v) >> gshift) & GM;
            const int src = who ? (__ffs(who) - 1) : 0;
            const unsigned sv = __shfl_sync(0xffffffffu, sup, src, GW);
            if (who != 0u) {   // who == 0 only with NaN confidences
                seen |= 1u << src;
                if (!((dead >> src) & 1u)) dead |= sv & ~seen & valid;
            }
        }
    }
    const unsigned alive = valid & ~dead;
    if (gl < n) {
        const bool kf = (alive >> gl) & 1u;
        keep[m] = kf ? 1 : 0;
        W.local_rank[m] = kf ? __popc(alive & ((1u << gl) - 1u)) : -1;
    }
    if (gl == 0 && seg >= 0) {
        W.seg_kept[seg] = (unsigned)__popc(alive);
        atomicAdd(&W.img_kept[seg / C], (unsigned)__popc(alive));
    }
}

// Work of one CTA on the warp-sized lists: tiny segments four to a warp, then 9..32-box segments
// one per warp; static round-robin over `n_workers` CTAs (a work-stealing atomic per segment
// costs more than the segment).
template <int MODE>
__device__ __forceinline__ void nms_small_work(const double* __restrict__ rows, double thr, double conf_thr, double sigma,
                                               const NmsWs& W, long long n_seg, int C, unsigned worker, unsigned n_workers,
                                               unsigned char* __restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const bool pos_thr = thr > 0.0;
    const unsigned wpb = blockDim.x >> 5;
    const unsigned warp_id = worker * wpb + (threadIdx.x >> 5), n_warps = n_workers * wpb;
    constexpr int kPerWarp = 32 / kTiny;
    const unsigned n_tiny = W.ctrl[2];
    for (unsigned w = warp_id; w * kPerWarp < n_tiny; w += n_warps) {
        const unsigned e = w * kPerWarp + (lane / kTiny);
        const int seg = (e < n_tiny) ? W.small_list[e] : -1;
        nms_group_segment<MODE, kTiny>(rows, thr, conf_thr, sigma, W, seg, pos_thr, C, keep);
    }
    const unsigned n_small = W.ctrl[0];
    for (unsigned w = warp_id; w < n_small; w += n_warps)
        nms_group_segment<MODE, 32>(rows, thr, conf_thr, sigma, W, W.small_list[n_seg - 1 - (long long)w], pos_thr, C, keep);
}

// ---- big segments: one CTA each ---------------------------------------------------
template <typename Less>
__device__ __forceinline__ void block_bitonic(int* h, int P, Less less) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P; t += blockDim.x) {
                const int u = t ^ j;
                if (u > t) {
                    const int a = h[t], b = h[u];
                    const bool up = (t & k) == 0;
                    if (up ? less(b, a) : less(a, b)) {
                        h[t] = b;
                        h[u] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Scratch shared by all segments of a CTA (static shared memory).
struct BigShared {
    unsigned long long mask[kSweep];
    double2 kX[kSweep], kY[kSweep], kC[kSweep];  // kept boxes of the current sweep block, compact
    double kA[kSweep];
    int kept[kSweep];
    int krow[kSweep];
    int nkept;
    int nlist;
    long long scan[32];
    unsigned short bpos[kSweep];            // visit positions of the current block (alive-list sweep)
    int wc[2][2][kBigThreads / 32];         // per-pass, per-item-row, per-warp survivor counts (ordered compaction)
};

// One segment.  The planes hold the boxes in VISIT order: X = (x0, x1), Y = (y0, y1), C = (x, y)
// (DIoU only), A = area; soft-NMS (MODE 3) keeps X = (x, y), Y = (w, h) instead.  cf (confidences,
// later the keep flags) aliases A.  SMEM only tells the compiler which address space the plane
// pointers live in (the function is inlined once per space).
template <int MODE, bool SMEM>
__device__ __forceinline__ void nms_segment(const double* __restrict__ rows, double thr, double conf_thr, double sigma,
                                            const NmsWs& W, int seg, int n_class, int n, long long start, int P, int* mem, int* ord,
                                            double2* X, double2* Y, double2* C, double* A, unsigned char* rem,
                                            unsigned short* alist, BigShared& S, unsigned char* __restrict__ keep) {
    constexpr int M = (MODE == 3) ? 1 : MODE;
    constexpr int kJpt = (M == 2) ? 1 : 2;   // boxes a thread carries at once through the kept list
    const int tid = threadIdx.x;
    const bool pos_thr = thr > 0.0;
    double* cf = A;

    if (n <= kCountSort) {
        // short segment: both orders by all-pairs rank counting (2 barriers each instead of a
        // barrier per bitonic pass - this path is the tail of every sparse-scene launch)
        const int my = (tid < n) ? W.members[start + tid] : INT_MAX;
        if (tid < n) ord[tid] = my;   // ord is scratch here
        __syncthreads();
        if (tid < n) {
            int rk = 0;
            for (int j = 0; j < n; ++j) rk += (ord[j] < my) ? 1 : 0;
            mem[rk] = my;             // 1. original order: row ids ascending
        }
        __syncthreads();
        double ck = 0.0;
        if (tid < n) {
            const double* r = rows + (long long)mem[tid] * 7;
            ck = __dmul_rn(r[4], r[6]);
            cf[tid] = ck;
        }
        __syncthreads();
        if (tid < n) {
            // 2. visit order: confidence descending, ties -> higher original index first
            int vis = 0;
            for (int j = 0; j < n; ++j) vis += (j != tid && visited_before(cf[j], j, ck, tid)) ? 1 : 0;
            ord[vis] = tid;
        }
        __syncthreads();
    } else {
        // 1. original order: sort row ids ascending
        for (int i = tid; i < P; i += kBigThreads) mem[i] = (i < n) ? W.members[start + i] : INT_MAX;
        __syncthreads();
        block_bitonic(mem, P, [](int a, int b) { return a < b; });
        for (int i = tid; i < n; i += kBigThreads) {
            const double* r = rows + (long long)mem[i] * 7;
            cf[i] = __dmul_rn(r[4], r[6]);
            ord[i] = i;
        }
        for (int i = n + tid; i < P; i += kBigThreads) ord[i] = INT_MAX;
        __syncthreads();
        // 2. visit order: confidence descending, ties -> higher original index first
        block_bitonic(ord, P, [cf, n](int a, int b) {
            if (a >= n || b >= n) return a < b;
            return visited_before(cf[a], a, cf[b], b);
        });
    }
    if (MODE == 3) {
        // soft-NMS: every box multiplies its confidence by exp(-IoU^2/sigma) for each earlier-visited
        // box it overlaps (in visit order); deleted <=> it was decayed below conf_thr.
        for (int v = tid; v < n; v += kBigThreads) {
            const double* r = rows + (long long)mem[ord[v]] * 7;
            X[v] = make_double2(r[0], r[1]);
            Y[v] = make_double2(r[2], r[3]);
        }
        __syncthreads();
        for (int v = tid; v < n; v += kBigThreads) {
            double c = cf[ord[v]];
            const double2 xy = X[v], wh = Y[v];
            bool decayed = false;
            for (int i = 0; i < v; ++i) {
                const double2 pxy = X[i], pwh = Y[i];
                const double iou = pair_iou<1>(pxy.x, pxy.y, pwh.x, pwh.y, xy.x, xy.y, wh.x, wh.y);
                if (iou >= thr) {
                    c = c * exp(-1.0 * (iou * iou) / sigma);
                    decayed = true;
                }
            }
            rem[v] = (decayed && c < conf_thr) ? 1 : 0;
        }
        __syncthreads();
    } else {
        // boxes in visit order (the confidence plane is dead from here on)
        __syncthreads();
        for (int v = tid; v < n; v += kBigThreads) {
            const double* r = rows + (long long)mem[ord[v]] * 7;
            const BoxC b = make_box(r[0], r[1], r[2], r[3]);
            X[v] = make_double2(b.x0, b.x1);
            Y[v] = make_double2(b.y0, b.y1);
            if (M == 2) C[v] = make_double2(b.cx, b.cy);
            A[v] = b.area;
            rem[v] = 0;
        }
        __syncthreads();
    }
    auto load_box = [&](int v) {
        BoxC b;
        const double2 x = X[v], y = Y[v];
        b.x0 = x.x; b.x1 = x.y; b.y0 = y.x; b.y1 = y.y;
        b.area = A[v];
        if (M == 2) { const double2 c = C[v]; b.cx = c.x; b.cy = c.y; } else { b.cx = 0.0; b.cy = 0.0; }
        return b;
    };
    // 3. blocked greedy sweep.  `alist` (shared-memory segments only) holds the visit positions
    //    that are still alive behind the last finished block, compacted after every block, so the
    //    threads share the remaining boxes evenly and no lane carries a dead box; without it the
    //    threads stride over all positions behind the block.
    if (SMEM && MODE != 3) {
        // Blocks of the next kSweep boxes that are STILL ALIVE, in visit order (the list of alive
        // positions is compacted in order after every block): a dense scene kills most boxes early,
        // so the number of blocks - each a mask build, a serial resolve and a handful of barriers -
        // follows the boxes that survive until they are visited, not the segment size.
        const unsigned short* cur = nullptr;       // nullptr: the identity list 0 .. n-1
        unsigned short* bufs[2] = {alist, alist + kBigCap};
        int which = 0, n_list = n;
        const int lane = tid & 31, wrp = tid >> 5;
        while (n_list > 0) {
            const int m = min(kSweep, n_list);
            if (tid < kSweep) {
                S.mask[tid] = 0ull;
                if (tid < m) S.bpos[tid] = cur ? cur[tid] : (unsigned short)tid;
            }
            __syncthreads();
            {   // 64x64 upper-triangular mask over the block's boxes, kMaskParts threads per row
                constexpr int kCols = 64 / kMaskParts;
                const int i = tid / kMaskParts, part = tid % kMaskParts;
                if (i < m) {
                    const int vi = S.bpos[i];
                    const BoxC bi = load_box(vi);
                    unsigned long long bits = 0ull;
                    const int j0 = max(part * kCols, i + 1), j1 = min(part * kCols + kCols, m);
                    for (int j = j0; j < j1; ++j) {
                        const int vj = S.bpos[j];
                        const BoxC bj = load_box(vj);
                        int r = suppresses_fast<M>(bi, bj, thr, pos_thr);
                        if (r < 0) r = suppresses_exact<M>(rows, mem[ord[vi]], mem[ord[vj]], thr) ? 1 : 0;
                        if (r) bits |= 1ull << j;
                    }
                    if (bits) atomicOr(&S.mask[i], bits);
                }
            }
            __syncthreads();
            const bool last_block = n_list == m;
            if (tid < 32) {
                // resolve the block in warp 0: lane l owns rows l and l+32; the serial chain over the
                // 64 rows runs on register shuffles, replicated in every lane
                const unsigned long long m0 = S.mask[tid], m1 = S.mask[tid + 32];
                unsigned long long dead = (m >= 64) ? 0ull : (~0ull << m);   // rows >= m do not exist
#pragma unroll
                for (int i = 0; i < kSweep; ++i) {
                    const unsigned long long mi = __shfl_sync(0xffffffffu, (i < 32) ? m0 : m1, i & 31);
                    if (!((dead >> i) & 1ull)) dead |= mi;
                }
                const unsigned long long kept_bits = ~dead;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int bp = tid + 32 * h;
                    if (bp < m) {
                        const bool kf = (kept_bits >> bp) & 1ull;
                        const int v = S.bpos[bp];
                        rem[v] = kf ? 0 : 1;
                        if (kf && !last_block) {   // kept boxes of this block, compact
                            const int q = __popcll(kept_bits & ((1ull << bp) - 1ull));
                            S.kX[q] = X[v];
                            S.kY[q] = Y[v];
                            if (M == 2) S.kC[q] = C[v];
                            S.kA[q] = A[v];
                            S.krow[q] = mem[ord[v]];
                        }
                    }
                }
                if (tid == 0) S.nkept = __popcll(kept_bits);
            }
            __syncthreads();
            if (last_block) break;   // nothing behind the block (uniform)
            const int nk = S.nkept;
            const int n_items = n_list - m;
            unsigned short* next = bufs[which];
            int out_base = 0;        // survivors written so far: the same value in every thread
            // every listed box behind the block against the block's kept boxes, kJpt boxes per
            // thread in flight (independent fp64 chains; the kept box is one broadcast load for all)
            for (int base = 0, pass = 0; base < n_items; base += kBigThreads * kJpt, ++pass) {
                double2 jx[kJpt], jy[kJpt];   // corners only: area / centre are fetched for overlapping pairs
                int jj[kJpt];
                bool alive[kJpt];
                bool any = false;
#pragma unroll
                for (int u = 0; u < kJpt; ++u) {
                    const int slot = base + u * kBigThreads + tid;
                    alive[u] = slot < n_items;
                    jj[u] = 0;
                    if (alive[u]) {
                        jj[u] = cur ? (int)cur[m + slot] : m + slot;
                        jx[u] = X[jj[u]];
                        jy[u] = Y[jj[u]];
                    }
                    any |= alive[u];
                }
                if (any) {
                    for (int q = 0; q < nk; ++q) {
                        const double2 ix = S.kX[q], iy = S.kY[q];
                        any = false;
#pragma unroll
                        for (int u = 0; u < kJpt; ++u) {
                            if (alive[u]) {
                                const double iw = sel_min(ix.y, jx[u].y) - sel_max(ix.x, jx[u].x);
                                const double ih = sel_min(iy.y, jy[u].y) - sel_max(iy.x, jy[u].x);
                                int r = -1;
                                if (pos_thr) {
                                    r = 0;
                                    if (iw > 0.0 && ih > 0.0) {
                                        double ew = 0.0, eh = 0.0, dx = 0.0, dy = 0.0;
                                        if (M == 2) {
                                            ew = sel_max(ix.y, jx[u].y) - sel_min(ix.x, jx[u].x);
                                            eh = sel_max(iy.y, jy[u].y) - sel_min(iy.x, jy[u].x);
                                            const double2 ic = S.kC[q], jc = C[jj[u]];
                                            dx = ic.x - jc.x;
                                            dy = ic.y - jc.y;
                                        }
                                        r = decide_overlapping<M>(iw, ih, S.kA[q], A[jj[u]], ew, eh, dx, dy, thr);
                                    }
                                }
                                if (r < 0) r = suppresses_exact<M>(rows, S.krow[q], mem[ord[jj[u]]], thr) ? 1 : 0;
                                if (r) {
                                    alive[u] = false;
                                    rem[jj[u]] = 1;
                                }
                            }
                            any |= alive[u];
                        }
                        if (!any) break;
                    }
                }
                // survivors of this pass go to the next list IN ORDER: per-warp counts, then every
                // thread walks the (item row, warp) table up to its own entry
                unsigned bal[kJpt];
#pragma unroll
                for (int u = 0; u < kJpt; ++u) {
                    bal[u] = __ballot_sync(0xffffffffu, alive[u]);
                    if (lane == 0) S.wc[pass & 1][u][wrp] = __popc(bal[u]);
                }
                __syncthreads();
                int run = out_base, my_off[kJpt];
#pragma unroll
                for (int u = 0; u < kJpt; ++u) {
                    for (int w = 0; w < kBigThreads / 32; ++w) {
                        if (w == wrp) my_off[u] = run;
                        run += S.wc[pass & 1][u][w];
                    }
                }
                out_base = run;
#pragma unroll
                for (int u = 0; u < kJpt; ++u)
                    if (alive[u]) next[my_off[u] + __popc(bal[u] & ((1u << lane) - 1u))] = (unsigned short)jj[u];
            }
            __syncthreads();   // the next list is complete
            cur = next;
            n_list = out_base;
            which ^= 1;
        }
    }
    unsigned short* cur_list = nullptr;   // nullptr: implicit list = every position behind the block
    int n_list = 0;
    for (int blk = 0; !SMEM && MODE != 3 && blk < n; blk += kSweep) {
        const int m = min(kSweep, n - blk);
        if (tid < kSweep) S.mask[tid] = 0ull;
        __syncthreads();
        {   // 64x64 upper-triangular mask, kMaskParts threads per row
            constexpr int kCols = 64 / kMaskParts;
            const int i = tid / kMaskParts, part = tid % kMaskParts;
            if (i < m && !rem[blk + i]) {
                const BoxC bi = load_box(blk + i);
                unsigned long long bits = 0ull;
                const int j0 = max(part * kCols, i + 1), j1 = min(part * kCols + kCols, m);
                for (int j = j0; j < j1; ++j) {
                    const BoxC bj = load_box(blk + j);
                    int r = suppresses_fast<M>(bi, bj, thr, pos_thr);
                    if (r < 0) r = suppresses_exact<M>(rows, mem[ord[blk + i]], mem[ord[blk + j]], thr) ? 1 : 0;
                    if (r) bits |= 1ull << j;
                }
                if (bits) atomicOr(&S.mask[i], bits);
            }
        }
        __syncthreads();
        const bool last_block = blk + m >= n;
        if (tid < 32) {
            // resolve the block in warp 0: lane l owns rows l and l+32; the serial chain over the 64
            // rows runs on register shuffles, replicated in every lane
            const unsigned long long m0 = S.mask[tid], m1 = S.mask[tid + 32];
            const bool d0 = (tid < m) ? (rem[blk + tid] != 0) : true;
            const bool d1 = (tid + 32 < m) ? (rem[blk + tid + 32] != 0) : true;
            unsigned long long dead = (unsigned long long)__ballot_sync(0xffffffffu, d0) |
                                      ((unsigned long long)__ballot_sync(0xffffffffu, d1) << 32);
#pragma unroll
            for (int i = 0; i < kSweep; ++i) {
                const unsigned long long mi = __shfl_sync(0xffffffffu, (i < 32) ? m0 : m1, i & 31);
                if (!((dead >> i) & 1ull)) dead |= mi;
            }
            const unsigned long long kept_bits = ~dead;   // rows >= m were dead from the start
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int bpos = tid + 32 * h;
                if (bpos < m) {
                    const bool kf = (kept_bits >> bpos) & 1ull;
                    rem[blk + bpos] = kf ? 0 : 1;
                    if (kf && !last_block) {   // kept boxes of this block, compact
                        const int q = __popcll(kept_bits & ((1ull << bpos) - 1ull));
                        const int v = blk + bpos;
                        S.kX[q] = X[v];
                        S.kY[q] = Y[v];
                        if (M == 2) S.kC[q] = C[v];
                        S.kA[q] = A[v];
                        S.krow[q] = mem[ord[v]];
                    }
                }
            }
            if (tid == 0) {
                S.nkept = __popcll(kept_bits);
                S.nlist = 0;
            }
        }
        __syncthreads();
        if (last_block) break;   // nothing left to suppress (uniform)
        const int nk = S.nkept;
        const int first = blk + m;
        const int n_items = (cur_list != nullptr) ? n_list : n - first;
        unsigned short* next_list = (alist == nullptr) ? nullptr : ((cur_list == alist) ? alist + kBigCap : alist);
        // every surviving box behind the block against the block's kept boxes, kJpt boxes per
        // thread in flight (independent fp64 chains; the kept box is one broadcast load for all)
        for (int base = 0; base < n_items; base += kBigThreads * kJpt) {
            double2 jx[kJpt], jy[kJpt];   // corners only: area / centre are fetched for overlapping pairs
            int jj[kJpt];
            bool alive[kJpt];
            bool any = false;
#pragma unroll
            for (int u = 0; u < kJpt; ++u) {
                const int slot = base + u * kBigThreads + tid;
                jj[u] = n;
                if (slot < n_items) jj[u] = (cur_list != nullptr) ? (int)cur_list[slot] : first + slot;
                alive[u] = jj[u] < n && jj[u] >= first && !rem[jj[u]];
                if (alive[u]) {
                    jx[u] = X[jj[u]];
                    jy[u] = Y[jj[u]];
                }
                any |= alive[u];
            }
            if (any) {
                for (int q = 0; q < nk; ++q) {
                    const double2 ix = S.kX[q], iy = S.kY[q];
                    any = false;
#pragma unroll
                    for (int u = 0; u < kJpt; ++u) {
                        if (alive[u]) {
                            const double iw = sel_min(ix.y, jx[u].y) - sel_max(ix.x, jx[u].x);
                            const double ih = sel_min(iy.y, jy[u].y) - sel_max(iy.x, jy[u].x);
                            int r = -1;
                            if (pos_thr) {
                                r = 0;
                                if (iw > 0.0 && ih > 0.0) {
                                    double ew = 0.0, eh = 0.0, dx = 0.0, dy = 0.0;
                                    if (M == 2) {
                                        ew = sel_max(ix.y, jx[u].y) - sel_min(ix.x, jx[u].x);
                                        eh = sel_max(iy.y, jy[u].y) - sel_min(iy.x, jy[u].x);
                                        const double2 ic = S.kC[q], jc = C[jj[u]];
                                        dx = ic.x - jc.x;
                                        dy = ic.y - jc.y;
                                    }
                                    r = decide_overlapping<M>(iw, ih, S.kA[q], A[jj[u]], ew, eh, dx, dy, thr);
                                }
                            }
                            if (r < 0) r = suppresses_exact<M>(rows, S.krow[q], mem[ord[jj[u]]], thr) ? 1 : 0;
                            if (r) {
                                alive[u] = false;
                                rem[jj[u]] = 1;
                            }
                        }
                        any |= alive[u];
                    }
                    if (!any) break;
                }
            }
            if (next_list != nullptr) {   // survivors of this pass go to the next block's list (any order)
#pragma unroll
                for (int u = 0; u < kJpt; ++u) {
                    const unsigned bal = __ballot_sync(0xffffffffu, alive[u]);
                    if (bal) {
                        const int lane = tid & 31;
                        int pos0 = 0;
                        if (lane == 0) pos0 = atomicAdd(&S.nlist, __popc(bal));
                        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
                        if (alive[u]) next_list[pos0 + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)jj[u];
                    }
                }
            }
        }
        __syncthreads();
        if (next_list != nullptr) {
            cur_list = next_list;
            n_list = S.nlist;
        }
    }
    // 4. survivors back in original order: keep flag per original position goes into
    //    the (now dead) confidence plane, then a block scan gives each survivor its rank
    for (int v = tid; v < n; v += kBigThreads) cf[ord[v]] = rem[v] ? 0.0 : 1.0;
    __syncthreads();
    long long carry = 0;
    for (int base = 0; base < n; base += kBigThreads) {
        const int i = base + tid;
        const int kf = (i < n && cf[i] != 0.0) ? 1 : 0;
        long long total;
        const long long ex = block_exclusive_scan((long long)kf, S.scan, total);
        if (i < n) {
            W.local_rank[mem[i]] = kf ? (int)(carry + ex) : -1;
            keep[mem[i]] = (unsigned char)kf;
        }
        carry += total;
    }
    if (tid == 0) {
        W.seg_kept[seg] = (unsigned)carry;
        atomicAdd(&W.img_kept[seg / n_class], (unsigned)carry);
    }
}

// All suppression work of one yb_nms call in ONE launch: the first CTAs take the big segments
// (longest jobs first, one CTA each), every other CTA works through the warp-sized lists.
template <int MODE>
__global__ void __launch_bounds__(kBigThreads, 2)
nms_sweep_kernel(const double* __restrict__ rows, double thr, double conf_thr, double sigma, NmsWs W, long long R,
                 long long n_seg, int C, unsigned big_blocks, unsigned char* __restrict__ keep) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s_X = reinterpret_cast<double2*>(smem_raw);                  // [kBigCap] each
    double2* s_Y = s_X + kBigCap;
    double2* s_C = s_Y + kBigCap;
    double* s_A = reinterpret_cast<double*>(s_C + kBigCap);
    int* s_mem = reinterpret_cast<int*>(s_A + kBigCap);                   // [kBigP]
    int* s_ord = s_mem + kBigP;                                           // [kBigP]
    unsigned char* s_rem = reinterpret_cast<unsigned char*>(s_ord + kBigP);  // [kBigCap]
    unsigned short* s_alist = reinterpret_cast<unsigned short*>(s_rem + kBigCap);   // [2][kBigCap] alive lists
    __shared__ BigShared S;

    const unsigned n_big = W.ctrl[1];
    const unsigned big_used = min(n_big, big_blocks);   // CTAs that have a big segment to start with
    if (blockIdx.x >= big_used) {
        nms_small_work<MODE>(rows, thr, conf_thr, sigma, W, n_seg, C, blockIdx.x - big_used, gridDim.x - big_used, keep);
        return;
    }
    for (unsigned wi = blockIdx.x; wi < n_big; wi += big_used) {
        __syncthreads();
        const int seg = W.big_list[wi];
        const int n = (int)W.seg_count[seg];
        const long long start = W.seg_start[seg];
        int P = 1;
        while (P < n) P <<= 1;
        if (n <= kBigCap) {
            nms_segment<MODE, true>(rows, thr, conf_thr, sigma, W, seg, C, n, start, P, s_mem, s_ord, s_X, s_Y, s_C, s_A,
                                    s_rem, s_alist, S, keep);
        } else {   // planes in the global scratch: [X 2R | Y 2R | C 2R | A R] doubles
            double2* gX = reinterpret_cast<double2*>(W.gbox) + start;
            double2* gY = reinterpret_cast<double2*>(W.gbox + 2 * R) + start;
            double2* gC = reinterpret_cast<double2*>(W.gbox + 4 * R) + start;
            double* gA = W.gbox + 6 * R + start;
            nms_segment<MODE, false>(rows, thr, conf_thr, sigma, W, seg, C, n, start, P, W.gmem_pad + 2 * start,
                                     W.gord_pad + 2 * start, gX, gY, gC, gA, W.gremoved + start, nullptr, S, keep);
        }
    }
}

// grid (n_img, Y): every CTA of image i sums the survivor counts of the earlier images, scans the
// image's C segments (per-(image, class) output extents) and writes its share of the survivors.
__global__ void __launch_bounds__(256)
nms_emit_kernel(const double* __restrict__ rows, const long long* __restrict__ row_offsets, long long n_img,
                long long cap, int C, NmsWs W, double* __restrict__ out_rows, long long* __restrict__ out_offsets,
                long long* __restrict__ out_seg_offsets) {
    __shared__ long long s_warp[32];
    const long long n = device_row_count(row_offsets, n_img, cap);
    const long long img = blockIdx.x, seg0 = img * C;
    const int tid = threadIdx.x;
    long long mine = 0;
    for (long long j = tid; j < img; j += blockDim.x) mine += W.img_kept[j];
    long long base;
    block_exclusive_scan(mine, s_warp, base);   // block total = survivors of the earlier images
    long long carry = 0;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + tid;
        const unsigned cnt = (c < C) ? W.seg_kept[seg0 + c] : 0u;
        long long total;
        const long long ex = block_exclusive_scan((long long)cnt, s_warp, total);
        if (c < C) {
            W.out_start[seg0 + c] = base + carry + ex;   // same value from every CTA of the image
            if (blockIdx.y == 0 && out_seg_offsets != nullptr) out_seg_offsets[seg0 + c] = base + carry + ex;
        }
        carry += total;
    }
    if (blockIdx.y == 0 && tid == 0) {
        if (out_offsets != nullptr) out_offsets[img] = base;
        if (img == n_img - 1) {
            if (out_offsets != nullptr) out_offsets[n_img] = base + carry;
            if (out_seg_offsets != nullptr) out_seg_offsets[n_img * C] = base + carry;
        }
    }
    if (out_rows == nullptr) return;
    __syncthreads();   // this CTA's out_start writes are visible to its own threads
    const long long r0 = min(row_offsets[img], n), r1 = min(row_offsets[img + 1], n);
    for (long long row = r0 + (long long)blockIdx.y * blockDim.x + tid; row < r1; row += (long long)gridDim.y * blockDim.x) {
        const int seg = W.row_seg[row];
        if (seg < 0) continue;
        const int r = W.local_rank[row];
        if (r < 0) continue;
        const double* src = rows + row * 7;
        double* dst = out_rows + (W.out_start[seg] + r) * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) dst[k] = src[k];
    }
}

// (na x nb) IoU / DIoU matrix.  A block owns kPairRows rows x 256 columns: the row boxes are staged
// in shared memory with their corners precomputed, each thread keeps one column box in registers,
// so a pair costs the min/max/area arithmetic and the IEEE divisions only - no index division, no
// reload.  Pairs with a non-finite corner or area take the NaN-propagating reference expression.
constexpr int kPairRows = 64;
struct PairRow {
    double x, y, w, h;         // the original row (exact path)
    double x0, x1, y0, y1, area;
    int finite;
};

__device__ __forceinline__ bool box_finite(const BoxC& c) {
    const double probe = ((c.x0 + c.x1) + (c.y0 + c.y1)) + c.area;   // inf or NaN anywhere -> not finite
    return fabs(probe) < 1.7e308;
}

template <int MODE>
__global__ void __launch_bounds__(256)
pairwise_iou_kernel(const double* __restrict__ a, long long na, int sa, const double* __restrict__ b,
                    long long nb, int sb, double* __restrict__ out) {
    __shared__ PairRow s_row[kPairRows];
    const long long ia0 = (long long)blockIdx.y * kPairRows;
    const int rows_here = (int)min((long long)kPairRows, na - ia0);
    if ((int)threadIdx.x < rows_here) {
        const double* t = a + (ia0 + threadIdx.x) * sa;
        PairRow r;
        r.x = t[0]; r.y = t[1]; r.w = t[2]; r.h = t[3];
        const double hw = r.w / 2.0, hh = r.h / 2.0;
        BoxC c;
        c.x0 = r.x - hw; c.x1 = r.x + hw; c.y0 = r.y - hh; c.y1 = r.y + hh; c.area = r.w * r.h; c.cx = r.x; c.cy = r.y;
        r.x0 = c.x0; r.x1 = c.x1; r.y0 = c.y0; r.y1 = c.y1; r.area = c.area;
        r.finite = box_finite(c) ? 1 : 0;
        s_row[threadIdx.x] = r;
    }
    __syncthreads();
    const long long ib = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ib >= nb) return;
    const double* p = b + ib * sb;
    const double px = p[0], py = p[1], pw = p[2], ph = p[3];
    BoxC pc;
    {
        const double hw = pw / 2.0, hh = ph / 2.0;
        pc.x0 = px - hw; pc.x1 = px + hw; pc.y0 = py - hh; pc.y1 = py + hh; pc.area = pw * ph; pc.cx = px; pc.cy = py;
    }
    const bool p_finite = box_finite(pc);
    double* o = out + ia0 * nb + ib;
    for (int r = 0; r < rows_here; ++r, o += nb) {
        const PairRow& t = s_row[r];
        double v;
        if (p_finite && t.finite) {
            const double iw = sel_max(sel_min(pc.x1, t.x1) - sel_max(pc.x0, t.x0), 0.0);
            const double ih = sel_max(sel_min(pc.y1, t.y1) - sel_max(pc.y0, t.y0), 0.0);
            const double inter = iw * ih;
            const double uni = pc.area + t.area - inter;
            v = inter / (uni + kIouEps);
            if (MODE == 2) {
                const double ew = sel_max(pc.x1, t.x1) - sel_min(pc.x0, t.x0), eh = sel_max(pc.y1, t.y1) - sel_min(pc.y0, t.y0);
                const double c2 = ew * ew + eh * eh;
                const double dx = t.x - px, dy = t.y - py;
                const double rho2 = dx * dx + dy * dy;
                v = v - rho2 / c2;
            }
        } else {
            v = pair_iou<MODE>(t.x, t.y, t.w, t.h, px, py, pw, ph);
        }
        *o = v;
    }
}

template <int MODE>
__global__ void elementwise_iou_kernel(const double* __restrict__ a, int sa, const double* __restrict__ b,
                                       int sb, long long n, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double* t = a + i * sa;
        const double* p = b + i * sb;
        out[i] = pair_iou<MODE>(t[0], t[1], t[2], t[3], p[0], p[1], p[2], p[3]);
    }
}

static size_t carve(size_t& off, size_t bytes) {
    const size_t at = off;
    off += align_up(bytes, 256);
    return at;
}

static size_t nms_layout(long long R, long long n_seg, NmsWs* W, char* base) {
    size_t off = 0;
    if (R < 1) R = 1;
    if (n_seg < 1) n_seg = 1;
#define YB_CARVE(field, type, count)                                   \
    {                                                                  \
        const size_t at = carve(off, sizeof(type) * (size_t)(count));  \
        if (W) W->field = reinterpret_cast<type*>(base + at);          \
    }
    // zero-initialised block first: seg_count, seg_fill, ctrl
    YB_CARVE(seg_count, unsigned int, n_seg + 1)
    YB_CARVE(seg_fill, unsigned int, n_seg)
    YB_CARVE(ctrl, unsigned int, 8)
    YB_CARVE(img_kept, unsigned int, n_seg + 1)   // n_img + 1 would do; n_seg bounds it without n_img here
    const size_t zero_bytes = off;
    YB_CARVE(seg_kept, unsigned int, n_seg + 1)
    YB_CARVE(row_seg, int, R)
    YB_CARVE(seg_start, long long, n_seg + 2)
    YB_CARVE(out_start, long long, n_seg + 2)
    YB_CARVE(members, int, R)
    YB_CARVE(local_rank, int, R)
    YB_CARVE(small_list, int, n_seg)
    YB_CARVE(big_list, int, n_seg)
    YB_CARVE(gbox, double, 7 * R)
    YB_CARVE(gmem_pad, int, 2 * R)
    YB_CARVE(gord_pad, int, 2 * R)
    YB_CARVE(gremoved, unsigned char, R)
#undef YB_CARVE
    (void)zero_bytes;
    return off;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_nms_workspace_bytes(int64_t n_rows, int64_t n_img, int class_num) {
    if (n_rows < 0 || n_img < 0 || class_num <= 0) return 0;
    return nms_layout(n_rows, n_img * class_num, nullptr, nullptr);
}

static int nms_impl(const double* rows, const int64_t* row_offsets_, int64_t n_rows, int64_t n_img,
                    int class_num, double nms_threshold, int iou_mode, double conf_thr, double sigma,
                    uint8_t* keep, double* out_rows, int64_t* out_offsets_, int64_t* out_seg_offsets_,
                    void* workspace, size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const long long* row_offsets = reinterpret_cast<const long long*>(row_offsets_);
    long long* out_offsets = reinterpret_cast<long long*>(out_offsets_);
    long long* out_seg_offsets = reinterpret_cast<long long*>(out_seg_offsets_);
    if (row_offsets == nullptr || workspace == nullptr) return YB_E_NULL;
    if (n_rows < 0 || n_img < 0 || class_num <= 0) return YB_E_SHAPE;
    if (iou_mode != 1 && iou_mode != 2 && iou_mode != 3) return YB_E_PARAM;
    if (n_rows > 0 && (rows == nullptr || keep == nullptr)) return YB_E_NULL;
    if (n_rows > (int64_t)INT_MAX / 2 || n_img > (int64_t)INT_MAX / 2) return YB_E_SHAPE;
    if (workspace_bytes < yb_nms_workspace_bytes(n_rows, n_img, class_num) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    const long long n_seg = n_img * class_num;
    if (n_img == 0 || n_rows == 0) {
        if (out_offsets != nullptr)
            YB_CUDA_TRY(cudaMemsetAsync(out_offsets, 0, sizeof(long long) * (n_img + 1), stream));
        if (out_seg_offsets != nullptr)
            YB_CUDA_TRY(cudaMemsetAsync(out_seg_offsets, 0, sizeof(long long) * (n_seg + 1), stream));
        return YB_OK;
    }
    NmsWs W;
    nms_layout(n_rows, n_seg, &W, reinterpret_cast<char*>(workspace));
    const size_t zero_bytes = (char*)W.seg_kept - (char*)workspace;
    YB_CUDA_TRY(cudaMemsetAsync(workspace, 0, zero_bytes, stream));

    const int threads = 256;
    const int row_blocks = (int)min((long long)kNumSMs * 8, ((long long)n_rows + threads - 1) / threads);
    nms_classify_kernel<<<row_blocks, threads, 0, stream>>>(rows, row_offsets, n_img, n_rows, class_num, W, keep);
    YB_CUDA_TRY(cudaGetLastError());
    // per-image kernels: grid (n_img, Y) with Y CTAs sharing the rows of one image
    const long long per_img = (n_rows + n_img - 1) / n_img;
    const dim3 img_grid((unsigned)n_img, (unsigned)max(1LL, min(64LL, (per_img + 2047) / 2048)));
    nms_scatter_kernel<<<img_grid, threads, 0, stream>>>(row_offsets, n_img, n_rows, class_num, W);
    YB_CUDA_TRY(cudaGetLastError());

    // one launch: up to 2 CTAs/SM worth of big-segment CTAs first, then the warp-list workers
    const size_t big_smem = sizeof(double) * 7 * kBigCap + sizeof(int) * 2 * kBigP + kBigCap + 2 * sizeof(unsigned short) * kBigCap;
    const unsigned big_blocks = (unsigned)min((long long)kNumSMs * 2, n_seg);
    const long long warp_jobs = (min((long long)n_rows, n_seg) + 7) / 8;   // <= one job per 8 segments... per warp
    const unsigned small_blocks = (unsigned)max(1LL, min((long long)kNumSMs * 4, (warp_jobs + 7) / 8));
#define YB_NMS_LAUNCH(M)                                                                                       \
    do {                                                                                                       \
        static SmemRaised done;                                                                    \
        YB_CUDA_TRY(raise_dynamic_smem_once(nms_sweep_kernel<M>, (int)big_smem, &done));                       \
        nms_sweep_kernel<M><<<big_blocks + small_blocks, kBigThreads, big_smem, stream>>>(                     \
            rows, nms_threshold, conf_thr, sigma, W, n_rows, n_seg, class_num, big_blocks, keep);                         \
    } while (0)
    if (iou_mode == 1) YB_NMS_LAUNCH(1);
    else if (iou_mode == 2) YB_NMS_LAUNCH(2);
    else YB_NMS_LAUNCH(3);
#undef YB_NMS_LAUNCH
    YB_CUDA_TRY(cudaGetLastError());
    if (out_rows != nullptr || out_offsets != nullptr || out_seg_offsets != nullptr) {
        nms_emit_kernel<<<img_grid, threads, 0, stream>>>(rows, row_offsets, n_img, n_rows, class_num, W,
                                                          out_rows, out_offsets, out_seg_offsets);
        YB_CUDA_TRY(cudaGetLastError());
    }
    return YB_OK;
}

extern "C" int yb_nms(const double* rows, const int64_t* row_offsets, int64_t n_rows, int64_t n_img,
                      int class_num, double nms_threshold, int iou_mode, uint8_t* keep, double* out_rows,
                      int64_t* out_offsets, int64_t* out_seg_offsets, void* workspace, size_t workspace_bytes,
                      yb_stream_t stream) {
    if (iou_mode != 1 && iou_mode != 2) return YB_E_PARAM;
    return nms_impl(rows, row_offsets, n_rows, n_img, class_num, nms_threshold, iou_mode, 0.0, 1.0, keep,
                    out_rows, out_offsets, out_seg_offsets, workspace, workspace_bytes, stream);
}

extern "C" int yb_soft_nms(const double* rows, const int64_t* row_offsets, int64_t n_rows, int64_t n_img,
                           int class_num, double nms_threshold, double conf_threshold, double sigma,
                           uint8_t* keep, double* out_rows, int64_t* out_offsets, int64_t* out_seg_offsets,
                           void* workspace, size_t workspace_bytes, yb_stream_t stream) {
    if (!(sigma != 0.0)) return YB_E_PARAM;
    return nms_impl(rows, row_offsets, n_rows, n_img, class_num, nms_threshold, 3, conf_threshold, sigma, keep,
                    out_rows, out_offsets, out_seg_offsets, workspace, workspace_bytes, stream);
}

extern "C" int yb_pairwise_iou(const double* a, int64_t na, int stride_a, const double* b, int64_t nb,
                               int stride_b, int iou_mode, double* out, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (na < 0 || nb < 0 || stride_a < 4 || stride_b < 4) return YB_E_SHAPE;
    if (iou_mode != 1 && iou_mode != 2) return YB_E_PARAM;
    if (na == 0 || nb == 0) return YB_OK;
    if (a == nullptr || b == nullptr || out == nullptr) return YB_E_NULL;
    const int threads = 256;
    const long long gx = (nb + threads - 1) / threads, gy = (na + kPairRows - 1) / kPairRows;
    if (gx > 0x7fffffffLL || gy > 65535) return YB_E_SHAPE;   // 4 M rows x 5e11 columns
    const dim3 grid((unsigned)gx, (unsigned)gy);
    if (iou_mode == 1)
        pairwise_iou_kernel<1><<<grid, threads, 0, stream>>>(a, na, stride_a, b, nb, stride_b, out);
    else
        pairwise_iou_kernel<2><<<grid, threads, 0, stream>>>(a, na, stride_a, b, nb, stride_b, out);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" int yb_elementwise_iou(const double* a, int stride_a, const double* b, int stride_b, int64_t n,
                                  int iou_mode, double* out, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n < 0 || stride_a < 4 || stride_b < 4) return YB_E_SHAPE;
    if (iou_mode != 1 && iou_mode != 2) return YB_E_PARAM;
    if (n == 0) return YB_OK;
    if (a == nullptr || b == nullptr || out == nullptr) return YB_E_NULL;
    const int threads = 256;
    const int blocks = (int)min((long long)kNumSMs * 8, ((long long)n + threads - 1) / threads);
    if (iou_mode == 1)
        elementwise_iou_kernel<1><<<blocks, threads, 0, stream>>>(a, stride_a, b, stride_b, n, out);
    else
        elementwise_iou_kernel<2><<<blocks, threads, 0, stream>>>(a, stride_a, b, stride_b, n, out);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}
