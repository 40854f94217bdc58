// Anchor k-means: one Lloyd assignment + accumulation pass over all boxes.
// Replaces utils/kmeans.py:79-90 (distance matrix, argmin, per-cluster mean) with a
// single streaming kernel; distances follow kmeans.py:9-33 (iou_dist: 1 - area
// ratio, NOT box overlap) and :36-40 (euclidean), float64, one rounding per
// operation (-fmad=false) so assignments are bit-identical to NumPy's argmin
// (first minimum wins).
//
// Every warp streams its own tiles (16 B per box for d=2) through its own shared-memory
// ring with 1-D bulk-async copies (TMA) - no block barrier in the loop; every thread owns a
// column of k*(d+1) accumulators in shared memory (plain load/add/store at a dynamic slot,
// no atomics), reduced warp -> CTA -> global partials -> last CTA in a fixed order
// (deterministic).  For (w,h) boxes with iou_dist the assignment needs no division for
// all but a vanishing fraction of the boxes (see the kernel).
#include <climits>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace yb {

constexpr int kKmThreads = 256;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmMaxStages = 4;

struct KmLaunch {
    const double* data;
    long long n;
    const double* centers;
    int k, d, kind;
    int bulk_ok;
    int tile_pts, n_stages;   // per-WARP tile (points) and ring depth
    int* assign;
    double* sums;       // [k][d]
    long long* counts;  // [k]
    double* partials;   // [grid][K*(D+1)]
    unsigned int* counter;
    // Lloyd loop on the device (yb_kmeans_lloyd_step): all optional
    long long* state;   // Lloyd state words; a non-zero status freezes the loop: the kernel returns at once
    double* packed;     // [k*d sums | k counts as doubles]: the payload of the all-reduce when sharded
    double* centers_rw; // non-null: the last CTA applies the Lloyd update itself (single rank)
    double stop_dist;
    long long max_iter;
    // peer exchange (yb_kmeans_lloyd_step_peers): the all-reduce of the k*(d+1) partial sums over
    // NVLink peer memory, inside the same launch
    unsigned char* mailbox[YB_MAX_PEERS];   // mailbox of every rank (own at [rank]); null = no exchange
    int rank, world;
};

// Mailbox of one rank: box[parity][source rank] = kPeerSlotDoubles payload doubles + a flag word.
constexpr int kPeerSlotDoubles = 16 * 5;                       // k <= 16, d <= 4: k*(d+1) <= 80
constexpr size_t kPeerSlotBytes = (kPeerSlotDoubles + 2) * 8;  // payload | flag | pad
__host__ __device__ inline size_t peer_slot_offset(int parity, int src, int world) {
    return ((size_t)parity * world + src) * kPeerSlotBytes;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All-reduce of fin[0..nv) over the ranks through their mailboxes (whole CTA; the last CTA of the
// launch): every rank stores its partial sums into slot [parity][rank] of EVERY mailbox (peer
// stores over NVLink), publishes them with a system-scope release of the exchange number, waits for
// the same number in all slots of its own mailbox and adds the slots in rank order - the same
// order on every rank, so all ranks hold the same bits.  Two parities: a rank can be at most one
// exchange ahead of a peer that is still reading.  A peer that never arrives (its host died) ends
// the wait after 2 s with status 4 instead of hanging the GPU.
__device__ inline bool km_peer_allreduce(const KmLaunch& L, double* fin, int nv) {
    const int tid = threadIdx.x;
    const unsigned long long seq = (unsigned long long)L.state[2] + 1ull;
    const int parity = (int)(seq & 1ull);
    __shared__ int s_ok;
    if (tid == 0) s_ok = 1;
    for (int i = tid; i < nv * L.world; i += blockDim.x) {
        const int peer = i / nv, j = i - peer * nv;
        double* slot = reinterpret_cast<double*>(L.mailbox[peer] + peer_slot_offset(parity, L.rank, L.world));
        slot[j] = fin[j];
    }
    __syncthreads();
    if (tid < L.world) {
        // the CTA's stores happen before this thread's fence (barrier), the fence before the flag: a
        // peer that acquires the flag sees them all (the pattern of a cooperative grid barrier) -
        // one fence per peer instead of one per thread
        __threadfence_system();
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(
            L.mailbox[tid] + peer_slot_offset(parity, L.rank, L.world) + kPeerSlotDoubles * 8);
        st_release_sys(flag, seq);
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(
            L.mailbox[L.rank] + peer_slot_offset(parity, tid, L.world) + kPeerSlotDoubles * 8);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(mine) != seq) {
            if (global_timer_ns() - t0 > 2000000000ull) {
                s_ok = 0;
                break;
            }
            __nanosleep(20);
        }
    }
    __syncthreads();
    if (!s_ok) {
        if (tid == 0) L.state[0] = 4;   // exchange timed out
        return false;
    }
    if (tid < nv) {
        double sum = 0.0;
        for (int src = 0; src < L.world; ++src)
            sum += __ldcg(reinterpret_cast<const double*>(L.mailbox[L.rank] + peer_slot_offset(parity, src, L.world)) + tid);
        fin[tid] = sum;
    }
    __syncthreads();
    if (tid == 0) L.state[2] = (long long)seq;
    return true;
}

// ---- the Lloyd update, utils/kmeans.py:84-97, as device code -------------------------------------
// state words (8 bytes each): [0] status 0 running | 1 loss < stop_dist | 2 iteration cap | 3 empty
// cluster (the host redraws: numpy.random owns that, kmeans.py:89), [1] completed updates,
// [2], [3] reserved, [4 .. 4+YB_KMEANS_HIST) loss of update e at slot (e-1) % YB_KMEANS_HIST,
// then k*(d+1) doubles: the global sums / counts of the iteration that found an empty cluster.
__device__ __forceinline__ double km_np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ double km_np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// np.add.reduce over a contiguous float64 vector (NumPy's pairwise summation, n <= 128)
__device__ inline double km_np_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// The first max(k, 1) threads of a CTA, all of them calling (one __syncthreads inside).  fin:
// [c*(d+1)+j] sums, [c*(d+1)+d] counts (as doubles, exact); dist: k doubles of shared scratch.
// Thread c owns cluster c (its three IEEE divisions run beside the other clusters'), thread 0
// closes the iteration.  Leaves the centres untouched when it freezes on an empty cluster.
__device__ inline void km_lloyd_update(int k, int d, int kind, const double* fin, const double* old_centers,
                                       double* centers, long long* state, double stop_dist, long long max_iter,
                                       double* dist) {
    const int c = threadIdx.x;
    double* hist = reinterpret_cast<double*>(state + 4);
    double* saved = hist + YB_KMEANS_HIST;
    bool empty = false;
    for (int q = 0; q < k; ++q) empty = empty || !(fin[q * (d + 1) + d] > 0.0);   // kmeans.py:85: len(index) > 0
    if (empty) {                                          // uniform: every thread read the same sums
        if (c == 0) {
            for (int i = 0; i < k * (d + 1); ++i) saved[i] = fin[i];
            __threadfence();
            state[0] = 3;
        }
        return;
    }
    if (c < k) {
        double nc[4];
        for (int j = 0; j < d; ++j) nc[j] = fin[c * (d + 1) + j] / fin[c * (d + 1) + d];   // cluster mean
        if (kind == YB_DIST_IOU) {                       // kmeans.py:12-22, 32
            const double ca = __dmul_rn(old_centers[c * d], old_centers[c * d + 1]);
            const double na = __dmul_rn(nc[0], nc[1]);
            dist[c] = 1.0 - km_np_min(ca, na) / km_np_max(ca, na);
        } else {                                         // kmeans.py:39
            double sq = 0.0;
            for (int j = 0; j < d; ++j) {
                const double df = old_centers[c * d + j] - nc[j];
                sq = __dadd_rn(sq, __dmul_rn(df, df));
            }
            dist[c] = sqrt(sq);
        }
        for (int j = 0; j < d; ++j) centers[c * d + j] = nc[j];
    }
    __syncthreads();
    if (c != 0) return;
    const double loss = km_np_sum(dist, k) / (double)k;  // np.mean, kmeans.py:92
    const long long done = state[1] + 1;
    hist[(done - 1) % YB_KMEANS_HIST] = loss;
    state[1] = done;                                     // (read by later launches and by the host after a sync)
    if (loss < stop_dist) state[0] = 1;                  // kmeans.py:97 (epoch = done + 1)
    else if (done + 1 > max_iter) state[0] = 2;
}

// NumPy's argmin over the rounded iou_dist values (kmeans.py:12-22,32,80), first minimum wins;
// returns the SORTED slot of the winner.  Cold path (near ties, degenerate centroids, NaN).
__device__ __noinline__ int km_exact_slot(double a, int k, const double* carea, const int* rank) {
    double bd = 0.0;
    int best = 0;
    for (int c = 0; c < k; ++c) {
        const double ca = carea[c];
        const double dist = 1.0 - fmin(ca, a) / fmax(ca, a);
        if (c == 0 || dist < bd) {
            bd = dist;
            best = c;
        }
    }
    return rank[best];
}

constexpr int kLutShift = 13;                 // high word >> 13: sign, exponent, 7 mantissa bits
constexpr int kLutCells = 2048;               // 16 octaves x 128 cells: areas in [2^-15, 2)
constexpr int kLutBase = (1023 - 15) << 7;    // cell index of 2^-15
constexpr int kCntBits = 7;                   // packed per-lane counters: 9 slots x 7 bits

// Shared memory: [ring: 8 warps x n_stages x tile_pts x D doubles][acc: K x 256 x D doubles][cnt: K x 256 ints]
// Every WARP streams its own tiles through its own ring (lane 0 issues the bulk copies, the warp
// waits on its own mbarriers): no block-wide barrier anywhere in the streaming loop.  Every thread
// owns one column of the accumulator planes, so the per-box update is a plain load / add / store
// at a dynamic slot index (no atomics, no k-way predicated adds).
#ifdef YB_KM_PROF
__device__ unsigned long long g_km_prof[1024 * 8];
#define KPROF(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_km_prof[blockIdx.x * 8 + (k)] = global_timer_ns(); } while (0)
extern "C" int yb_debug_km_prof(unsigned long long* host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, g_km_prof, sizeof(unsigned long long) * (size_t)n_words);
}
#else
#define KPROF(k) do { } while (0)
#endif

template <int K, int D, bool kIou, bool kAssign>
__global__ void __launch_bounds__(kKmThreads, 2)
kmeans_assign_kernel(const __grid_constant__ KmLaunch L) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* ring = reinterpret_cast<double*>(smem);
    double* s_acc = ring + (size_t)kKmWarps * L.n_stages * L.tile_pts * D;    // [K][256][D]
    int* s_cnt = reinterpret_cast<int*>(s_acc + (size_t)K * D * kKmThreads);  // [K][256]
    double* s_red = ring;                  // [8][K*(D+1)], after the streaming loop (ring is dead)
    constexpr bool kBoxes = kIou && D == 2;   // the anchor case: (w,h) boxes, iou_dist
    __shared__ uint64_t full[kKmWarps][kKmMaxStages];
    __shared__ double s_center[K * D];
    __shared__ double s_carea[K];
    __shared__ double s_sarea[K];          // centroid areas, ascending
    __shared__ int s_sidx[K];              // original index of the s-th smallest area
    __shared__ int s_rank[K];              // sorted slot of original index c
    __shared__ int s_thr_hi[K];            // high words of the k-1 decision thresholds (INT_MAX beyond)
    __shared__ int2 s_win[K];              // slot s is certain for win.x < hiword(area) < win.y
    __shared__ unsigned char s_lut[kBoxes ? kLutCells : 16];  // cell of hiword(area) -> certain slot | 0xFF
    __shared__ int s_is_last, s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = L.k;  // <= K
    const int n_stages = L.n_stages, tile_pts = L.tile_pts;
    if (L.state != nullptr && __ldcg(&L.state[0]) != 0) return;   // frozen Lloyd loop: nothing to do
    KPROF(0);

    if (tid == 0) {
        for (int w = 0; w < kKmWarps; ++w)
            for (int i = 0; i < kKmMaxStages; ++i) mbar_init(&full[w][i], 1);
        mbar_fence_init();
        s_bad = (D < 2) ? 1 : 0;
    }
    if (tid < K * D) s_center[tid] = (tid < k * D) ? L.centers[tid] : 0.0;
    for (int i = tid; i < K * D * kKmThreads; i += kKmThreads) s_acc[i] = 0.0;
    for (int i = tid; i < K * kKmThreads; i += kKmThreads) s_cnt[i] = 0;
    __syncthreads();
    // the first tiles of every warp's ring are on their way while the CTA builds its tables
    const long long n_tiles = (L.n + tile_pts - 1) / tile_pts;
    const long long wg = (long long)blockIdx.x * kKmWarps + warp, n_wg = (long long)gridDim.x * kKmWarps;
    const long long n_my = (n_tiles > wg) ? (n_tiles - wg + n_wg - 1) / n_wg : 0;
    double* my_ring = ring + (size_t)warp * n_stages * tile_pts * D;
    uint64_t* my_full = full[warp];
    const long long tile_step = n_wg * tile_pts;
    auto issue = [&](long long p0, int stage) {   // lane 0 only
        const int np = (int)min((long long)tile_pts, L.n - p0);
        const uint32_t bytes = (uint32_t)np * D * 8u;
        if (L.bulk_ok && (bytes & 15u) == 0u) {
            mbar_arrive_expect_tx(&my_full[stage], bytes);
            bulk_g2s(my_ring + (size_t)stage * tile_pts * D, L.data + p0 * D, bytes, &my_full[stage]);
        } else {
            mbar_arrive(&my_full[stage]);  // the warp reads global memory directly for this tile
        }
    };
    long long p_issue = wg * tile_pts;   // first point of the next tile to issue
    if (lane == 0)
        for (int t = 0; t < n_stages - 1 && t < n_my; ++t, p_issue += tile_step) issue(p_issue, t);
    if (tid < K) s_carea[tid] = (tid < k && D >= 2) ? __dmul_rn(s_center[tid * D], s_center[tid * D + 1]) : 0.0;
    __syncthreads();
    // iou_dist is a function of the AREA only and monotone in it on either side of the box's area
    // a, so the nearest centroid is one of the two sorted areas lo <= a < hi bracketing the box, and
    // lo/a > a/hi  <=>  a < sqrt(lo*hi).  The assignment is therefore a step function of a with
    // k-1 thresholds g_j = sqrt(A_j * A_j+1) over the sorted areas A.  The shortcut is accepted only
    // where it provably agrees with NumPy's first-minimum argmin over the ROUNDED distances
    // fl(1 - fl(min/max)):  slot s is certain for  g_{s-1}(1+d) < a < g_s(1-d)  with
    // d = 2e-15 + 1e-15 sqrt(A_j+1 / A_j): there the two candidate ratios differ by > 2e-15, more
    // than the roundings of the ratios and of 1 - r can hide (4.4e-16), so the rounded distances
    // are ordered like the exact ones, strictly; the window is also clamped so that the winning
    // ratio is >= 2^-20, and neighbouring areas must differ by > 1e-6 relative, so every farther
    // centroid's rounded distance is strictly larger still (no tie to break).
    // Everything is decided on the HIGH WORD of the area (positive doubles order like their bit
    // patterns): a table over cells of 2^-7 relative width maps a cell that lies entirely inside
    // one slot's certainty window to that slot (one byte load per box); a cell that touches a
    // threshold counts the thresholds below the high word and tests the slot's window; what is
    // still uncertain (near ties, duplicate / non-finite / extreme centroids, far-away boxes, NaN)
    // takes the exact k-way loop with IEEE divisions.
    // (thread c ranks centroid c, thread sl derives slot sl's window: the square roots and divisions
    // of the k slots run side by side - on one thread they were ~10 us in front of every launch)
    if (tid < k) {
        const double a = s_carea[tid];
        if (!(a > 1e-100) || !(a < 1e100)) s_bad = 1;
        int r = 0;
        for (int q = 0; q < k; ++q) r += (s_carea[q] < a || (s_carea[q] == a && q < tid)) ? 1 : 0;
        s_sarea[r] = a;
        s_sidx[r] = tid;
        s_rank[tid] = r;
    }
    __syncthreads();
    if (tid + 1 < k && !(s_sarea[tid + 1] - s_sarea[tid] > 1e-6 * s_sarea[tid + 1])) s_bad = 1;
    __syncthreads();
    if (tid < K) {
        const int sl = tid;
        const bool bad = s_bad != 0;
        const double tiny = 9.5367431640625e-07, huge = 1048576.0;  // 2^-20, 2^20
        double lo = INFINITY, hi = -INFINITY;   // never certain
        int th = INT_MAX;
        if (!bad && sl < k) {
            lo = s_sarea[sl] * tiny * (1.0 + 1e-12);
            hi = s_sarea[sl] * huge * (1.0 - 1e-12);
            if (sl > 0) {
                const double A0 = s_sarea[sl - 1], A1 = s_sarea[sl];
                const double g = sqrt(A0 * A1), dl = 2e-15 + 1e-15 * sqrt(A1 / A0);
                lo = fmax(lo, g * (1.0 + dl));
            }
            if (sl + 1 < k) {
                const double A0 = s_sarea[sl], A1 = s_sarea[sl + 1];
                const double g = sqrt(A0 * A1), dl = 2e-15 + 1e-15 * sqrt(A1 / A0);
                hi = fmin(hi, g * (1.0 - dl));
                th = __double2hiint(g);
            }
        }
        // hiword(a) > hiword(lo) => a > lo;  hiword(a) < hiword(hi) => a < hi
        s_win[sl] = (lo < hi) ? make_int2(__double2hiint(lo), __double2hiint(hi)) : make_int2(INT_MAX, INT_MIN);
        s_thr_hi[sl] = th;
    }
    __syncthreads();
    int thr_hi[K > 1 ? K - 1 : 1];   // threshold high words in registers
#pragma unroll
    for (int c = 0; c + 1 < K; ++c) thr_hi[c] = s_thr_hi[c];
    auto count_slot = [&](int ah) {
        int sl = 0;
#pragma unroll
        for (int c = 0; c + 1 < K; ++c) sl += (ah > thr_hi[c]) ? 1 : 0;
        return sl;
    };
    if (kBoxes) {
        for (int cell = tid; cell < kLutCells; cell += kKmThreads) {
            const int lo_h = (kLutBase + cell) << kLutShift, hi_h = lo_h + ((1 << kLutShift) - 1);
            const int s1 = count_slot(lo_h), s2 = count_slot(hi_h);
            const int2 w = s_win[s1];
            s_lut[cell] = (s1 == s2 && lo_h > w.x && hi_h < w.y) ? (unsigned char)s1 : (unsigned char)0xFF;
        }
        __syncthreads();
    }

    double* my_acc = s_acc + tid * D;
    int* my_cnt = s_cnt + tid;

    // per-lane box counters: 7-bit fields of one 64-bit register (K <= 9), flushed to the shared
    // counter plane before a field can overflow; larger K counts in shared memory directly
    constexpr bool kPacked = kBoxes && K * kCntBits <= 64;
    unsigned long long cnt_pack = 0ull;
    int cnt_pending = 0;
    auto flush_counts = [&]() {
#pragma unroll
        for (int c = 0; c < K; ++c) my_cnt[c * kKmThreads] += (int)((cnt_pack >> (kCntBits * c)) & ((1u << kCntBits) - 1u));
        cnt_pack = 0ull;
        cnt_pending = 0;
    };

    // one batch of kU boxes of a lane: loads, areas and slots of the whole batch are independent
    // chains; only the accumulator updates are ordered
    constexpr int kU = 4;
    auto process_boxes = [&](const double2 (&bx)[kU], const bool (&have)[kU], long long first) {
        int slot[kU];
        double area[kU];
        bool all_sure = true;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            area[u] = __dmul_rn(bx[u].x, bx[u].y);
            const unsigned cell = (unsigned)((__double2hiint(area[u]) >> kLutShift) - kLutBase);
            slot[u] = (cell < (unsigned)kLutCells) ? (int)s_lut[cell] : 0xFF;
            all_sure = all_sure && (slot[u] != 0xFF || !have[u]);
        }
        if (!all_sure) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (have[u] && slot[u] == 0xFF) {
                    const int ah = __double2hiint(area[u]);
                    const int sl = count_slot(ah);
                    const int2 w = s_win[sl];
                    slot[u] = (ah > w.x && ah < w.y) ? sl : km_exact_slot(area[u], k, s_carea, s_rank);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (have[u]) {
                if (kAssign) L.assign[first + u * 32] = s_sidx[slot[u]];
                double2* acc = reinterpret_cast<double2*>(my_acc + slot[u] * (2 * kKmThreads));
                double2 t2 = *acc;
                t2.x += bx[u].x;
                t2.y += bx[u].y;
                *acc = t2;
                if (kPacked) cnt_pack += 1ull << (kCntBits * slot[u]);
                else my_cnt[slot[u] * kKmThreads] += 1;
            }
        }
        if (kPacked) {
            cnt_pending += kU;
            if (cnt_pending > (1 << kCntBits) - 1 - kU) flush_counts();
        }
    };
    // generic point (any D, euclidean or iou_dist): exact k-way loop
    auto process_point = [&](const double* src, int i, long long p0) {
        double v[D];
#pragma unroll
        for (int j = 0; j < D; ++j) v[j] = src[(size_t)i * D + j];
        int slot = 0;   // accumulator slot: SORTED position for iou_dist, cluster index otherwise
        if (kIou) {
            const double a = __dmul_rn(v[0], v[D > 1 ? 1 : 0]);
            slot = km_exact_slot(a, k, s_carea, s_rank);
            if (kAssign) L.assign[p0 + i] = s_sidx[slot];
        } else {
            double bd = 0.0;
            for (int c = 0; c < k; ++c) {
                double sq = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    const double df = s_center[c * D + j] - v[j];
                    sq = __dadd_rn(sq, __dmul_rn(df, df));
                }
                const double dist = sqrt(sq);
                if (c == 0 || dist < bd) {
                    bd = dist;
                    slot = c;
                }
            }
            if (kAssign) L.assign[p0 + i] = slot;
        }
        double* acc = my_acc + slot * (D * kKmThreads);
#pragma unroll
        for (int j = 0; j < D; ++j) acc[j] += v[j];
        my_cnt[slot * kKmThreads] += 1;
    };

    KPROF(1);
    int stage = 0, issue_stage = n_stages - 1;
    uint32_t parity = 0;
    long long p0 = wg * tile_pts;
    for (long long it = 0; it < n_my; ++it, p0 += tile_step) {
        // the stage refilled here held tile it-1: every lane left it before the __syncwarp below
        if (lane == 0 && it + n_stages - 1 < n_my) {
            issue(p_issue, issue_stage);
            p_issue += tile_step;
        }
        issue_stage = (issue_stage + 1 == n_stages) ? 0 : issue_stage + 1;
        const int np = (int)min((long long)tile_pts, L.n - p0);
        const bool staged = L.bulk_ok && ((((uint32_t)np * D * 8u) & 15u) == 0u);
        const double* src = staged ? my_ring + (size_t)stage * tile_pts * D : L.data + p0 * D;
        mbar_wait(&my_full[stage], parity);
        if (kBoxes) {
            if (np == tile_pts && (tile_pts % (32 * kU)) == 0) {   // full tile of whole batches: no bounds tests
                const bool have[kU] = {true, true, true, true};
                for (int base = 0; base < tile_pts; base += 32 * kU) {
                    double2 bx[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u)
                        bx[u] = *reinterpret_cast<const double2*>(src + (size_t)(base + u * 32 + lane) * 2);
                    process_boxes(bx, have, p0 + base + lane);
                }
            } else {
                for (int base = 0; base < np; base += 32 * kU) {
                    double2 bx[kU];
                    bool have[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const int i = base + u * 32 + lane;
                        have[u] = i < np;
                        bx[u] = have[u] ? *reinterpret_cast<const double2*>(src + (size_t)i * 2) : make_double2(0.0, 0.0);
                    }
                    process_boxes(bx, have, p0 + base + lane);
                }
            }
        } else {
            for (int i = lane; i < np; i += 32) process_point(src, i, p0);
        }
        __syncwarp();  // stage free for the next bulk load
        if (++stage == n_stages) {
            stage = 0;
            parity ^= 1u;
        }
    }
    if (kPacked) flush_counts();
    KPROF(2);
    __syncthreads();   // every warp has left its ring: s_red may reuse it
    KPROF(3);

    // ---- reduction: thread columns -> warp -> CTA -> global partials -> last CTA ----
    constexpr int NV = K * (D + 1);
    // value v = (cluster c, coordinate j | count) lives in 256 thread columns of one accumulator plane
    // (cluster c's plane is its sorted rank for boxes).  Thread (v, seg) adds 32 columns, in a fixed
    // order that starts at its own lane (spreads the shared-memory banks), then eight segments are
    // added per value: plain shared-memory loads instead of 27 x 5 rounds of fp64 shuffles per warp.
    constexpr int kSegs = kKmThreads / 32;
    for (int idx = tid; idx < NV * kSegs; idx += kKmThreads) {
        const int v = idx / kSegs, seg = idx - v * kSegs;
        const int c = v / (D + 1), j = v - c * (D + 1);
        const int sl = (kBoxes && c < k) ? s_rank[c] : c;
        double sres = 0.0;
        if (j < D) {
            const double* plane = s_acc + (size_t)sl * (D * kKmThreads) + j;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) sres += plane[(size_t)(seg * 32 + ((r + lane) & 31)) * D];
        } else {
            const int* plane = s_cnt + sl * kKmThreads;
            int cs = 0;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) cs += plane[seg * 32 + ((r + lane) & 31)];
            sres = (double)cs;   // exact below 2^53
        }
        s_red[seg * NV + v] = sres;
    }
    __syncthreads();
    if (tid < NV) {
        double sres = 0.0;
        for (int w = 0; w < kSegs; ++w) sres += s_red[w * NV + tid];
        L.partials[(size_t)blockIdx.x * NV + tid] = sres;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(L.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    KPROF(4);
    if (!s_is_last) return;
    __threadfence();
    double* s_fin = s_red;   // [K*(D+1)]: the global sums / counts (s_red is dead: partials are in global memory)
    __syncthreads();
    // the partials of all CTAs: a [gridDim][NV] matrix.  Thread (g, i) adds the rows g, g + G, ... of
    // column i (neighbouring threads read neighbouring doubles, many rows in flight per thread), then
    // the G groups are added per column - a fixed order, so every run and every rank gets the same bits
    {
        constexpr int G = kKmThreads / NV;
        double* s_grp = s_red + NV * 2;                  // [G][NV], behind s_fin and the update's scratch
        if (tid < G * NV) {
            const int g = tid / NV, i = tid - g * NV;
            double sres = 0.0;
#pragma unroll 8
            for (int b2 = g; b2 < (int)gridDim.x; b2 += G) sres += __ldcg(&L.partials[(size_t)b2 * NV + i]);
            s_grp[g * NV + i] = sres;
        }
        __syncthreads();
        if (tid < NV) {
            const int i = tid;
            double sres = 0.0;
            for (int g = 0; g < G; ++g) sres += s_grp[g * NV + i];
            const int c = i / (D + 1), j = i - c * (D + 1);
            s_fin[i] = sres;
            if (c < k) {
                if (j < D) {
                    if (L.sums != nullptr) L.sums[c * D + j] = sres;
                    if (L.packed != nullptr) L.packed[c * D + j] = sres;
                } else {
                    if (L.counts != nullptr) L.counts[c] = (long long)sres;
                    if (L.packed != nullptr) L.packed[k * D + c] = sres;
                }
            }
        }
    }
    __syncthreads();
    KPROF(5);
    if (tid == 0) *L.counter = 0u;
    bool ok = true;
    if (L.mailbox[0] != nullptr) {
        // s_fin is laid out [c*(D+1)+j] over K slots; the exchange covers the first k*(D+1) ... K may
        // exceed k, so send all NV entries (unused ones are zero)
        ok = km_peer_allreduce(L, s_fin, NV);
    }
    KPROF(6);
    // every CTA has long copied the centres: update them in place (ok is the same in every thread)
    if (ok && L.centers_rw != nullptr)
        km_lloyd_update(k, D, L.kind, s_fin, s_center, L.centers_rw, L.state, L.stop_dist, L.max_iter, s_fin + NV);
    KPROF(7);
}

// sharded Lloyd loop: the update alone, on the all-reduced [sums | counts] (one thread)
__global__ void kmeans_update_kernel(const double* __restrict__ packed, double* centers, int k, int d, int kind,
                                     long long* state, double stop_dist, long long max_iter) {
    __shared__ double fin[16 * 5], dist[16];
    if (state[0] != 0) return;
    for (int i = threadIdx.x; i < k * (d + 1); i += blockDim.x) {
        const int c = i / (d + 1), j = i - c * (d + 1);
        fin[i] = j < d ? packed[c * d + j] : packed[k * d + c];
    }
    __syncthreads();
    km_lloyd_update(k, d, kind, fin, centers, centers, state, stop_dist, max_iter, dist);
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

// utils/kmeans.py:9-40 as a plain distance evaluation (the module's iou / iou_dist /
// euclidean_dist called directly on data-sized arrays): outer = every centre against every point
// -> (k, M), the broadcast of kmeans.py:79; otherwise element by element -> (M).
// what: 0 iou (area ratio), 1 iou_dist = 1 - iou, 2 euclidean
__global__ void kmeans_dist_kernel(const double* __restrict__ a, long long na, const double* __restrict__ b,
                                   long long nb, int d, int what, int outer, double* __restrict__ out) {
    const long long total = outer ? na * nb : nb;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long ia = outer ? t / nb : t, ib = outer ? t - ia * nb : t;
        const double* pa = a + ia * d;
        const double* pb = b + ib * d;
        double v;
        if (what == 2) {
            double sq = 0.0;
            for (int j = 0; j < d; ++j) {
                const double df = pa[j] - pb[j];
                sq = __dadd_rn(sq, __dmul_rn(df, df));
            }
            v = sqrt(sq);
        } else {
            const double ca = __dmul_rn(pa[0], pa[1]), da = __dmul_rn(pb[0], pb[1]);
            v = km_np_min(ca, da) / km_np_max(ca, da);
            if (what == 1) v = 1.0 - v;
        }
        out[t] = v;
    }
}

constexpr int kKmGrid = kNumSMs * 4;  // upper bound of the launch grid (partials sizing)

template <int K, int D>
static int launch_km(KmLaunch L, cudaStream_t stream) {
    // accumulator planes + 8 per-warp rings must fit: prefer 2 stages of 256-point tiles per warp and
    // 2 CTAs per SM (benchmarks/km_sweep.sh)
    const size_t acc = (size_t)K * kKmThreads * (D * sizeof(double) + sizeof(int));
    const size_t budget2 = (227 * 1024) / 2 - 4 * 1024;   // per CTA with 2 CTAs/SM (static smem + reservation)
    const size_t budget1 = 227 * 1024 - 8 * 1024;
    int tile = 256, stages = 2;
    auto need = [&](int t, int st) { return acc + (size_t)kKmWarps * st * t * D * sizeof(double); };
    while (need(tile, stages) > budget2 && stages > 2) --stages;
    while (need(tile, stages) > budget2 && tile > 32) tile >>= 1;
    if (need(tile, stages) > budget2) {   // fat k*d: one CTA per SM
        tile = 128;
        stages = 3;
        while (need(tile, stages) > budget1 && tile > 32) tile >>= 1;
        if (need(tile, stages) > budget1) return YB_E_SHAPE;
    }
    int ctas = (need(tile, stages) <= budget2) ? 2 : 1;
    {   // tuning hooks (debug): YB_KM_TILE / YB_KM_STAGES / YB_KM_CTAS, read once per process
        struct Env {
            int tile, stages, ctas;
            static int get(const char* n) { const char* e = getenv(n); return (e && *e) ? atoi(e) : 0; }
            Env() : tile(get("YB_KM_TILE")), stages(get("YB_KM_STAGES")), ctas(get("YB_KM_CTAS")) {}
        };
        static const Env env;
        if (env.tile >= 32) tile = env.tile / 32 * 32;
        if (env.stages >= 2 && env.stages <= kKmMaxStages) stages = env.stages;
        if (env.ctas >= 1 && env.ctas <= 4) ctas = env.ctas;
        if (need(tile, stages) > budget1) return YB_E_SHAPE;
    }
    L.tile_pts = tile;
    L.n_stages = stages;
    const size_t smem = need(tile, stages);
    const long long n_tiles = (L.n + tile - 1) / tile;
    const int grid = (int)max(1LL, min((long long)kNumSMs * ctas, (n_tiles + kKmWarps - 1) / kKmWarps));
#define YB_KM_LAUNCH(IOU, ASSIGN)                                                                          \
    do {                                                                                                   \
        static SmemRaised done;                                                                \
        YB_CUDA_TRY(raise_dynamic_smem_once(kmeans_assign_kernel<K, D, IOU, ASSIGN>, (int)smem, &done));  \
        kmeans_assign_kernel<K, D, IOU, ASSIGN><<<grid, kKmThreads, smem, stream>>>(L);                    \
    } while (0)
    const bool iou = L.kind == YB_DIST_IOU, asg = L.assign != nullptr;
    if (iou && asg) YB_KM_LAUNCH(true, true);
    else if (iou) YB_KM_LAUNCH(true, false);
    else if (asg) YB_KM_LAUNCH(false, true);
    else YB_KM_LAUNCH(false, false);
#undef YB_KM_LAUNCH
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

template <int D>
static int launch_km_d(const KmLaunch& L, cudaStream_t stream) {
    if (L.k <= 4) return launch_km<4, D>(L, stream);
    if (L.k <= 8) return launch_km<8, D>(L, stream);
    if (L.k <= 12) return launch_km<12, D>(L, stream);
    return launch_km<16, D>(L, stream);
}

// (w, h) boxes - the anchor case: exact k so the register accumulators are not padded
static int launch_km_boxes(const KmLaunch& L, cudaStream_t stream) {
    switch (L.k) {
        case 1: return launch_km<1, 2>(L, stream);
        case 2: return launch_km<2, 2>(L, stream);
        case 3: return launch_km<3, 2>(L, stream);
        case 4: return launch_km<4, 2>(L, stream);
        case 5: return launch_km<5, 2>(L, stream);
        case 6: return launch_km<6, 2>(L, stream);
        case 7: return launch_km<7, 2>(L, stream);
        case 8: return launch_km<8, 2>(L, stream);
        case 9: return launch_km<9, 2>(L, stream);
        case 10: return launch_km<10, 2>(L, stream);
        case 11: return launch_km<11, 2>(L, stream);
        case 12: return launch_km<12, 2>(L, stream);
        case 13: return launch_km<13, 2>(L, stream);
        case 14: return launch_km<14, 2>(L, stream);
        case 15: return launch_km<15, 2>(L, stream);
        default: return launch_km<16, 2>(L, stream);
    }
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
    L.state = nullptr;
    L.packed = nullptr;
    L.centers_rw = nullptr;
    L.stop_dist = 0.0;
    L.max_iter = 0;
    for (int i = 0; i < YB_MAX_PEERS; ++i) L.mailbox[i] = nullptr;
    L.rank = 0;
    L.world = 1;
    YB_CUDA_TRY(cudaMemsetAsync(L.counter, 0, sizeof(unsigned int), stream));
    switch (n_dim) {
        case 1: return launch_km_d<1>(L, stream);
        case 2: return launch_km_boxes(L, stream);
        case 3: return launch_km_d<3>(L, stream);
        default: return launch_km_d<4>(L, stream);
    }
}

extern "C" int yb_kmeans_dist(const double* a, int64_t na, const double* b, int64_t nb, int n_dim, int what,
                              int outer, double* out, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (na < 0 || nb < 0 || n_dim < 1 || n_dim > 7) return YB_E_SHAPE;   // sums of < 8 terms: sequential in NumPy too
    if (what < 0 || what > 2) return YB_E_PARAM;
    if (what != 2 && n_dim < 2) return YB_E_SHAPE;
    if (!outer && na != nb) return YB_E_SHAPE;
    const long long total = outer ? na * nb : nb;
    if (total == 0) return YB_OK;
    if (a == nullptr || b == nullptr || out == nullptr) return YB_E_NULL;
    const int blocks = (int)min((long long)kNumSMs * 8, (total + 255) / 256);
    kmeans_dist_kernel<<<blocks, 256, 0, stream>>>(a, na, b, nb, n_dim, what, outer, out);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

extern "C" size_t yb_kmeans_state_bytes(int k, int n_dim) {
    if (k < 1 || k > 16 || n_dim < 1 || n_dim > 4) return 0;
    return sizeof(long long) * (size_t)(4 + YB_KMEANS_HIST + k * (n_dim + 1));
}

static int lloyd_step_impl(const double* data, int64_t n_points, int n_dim, double* centers, int k, int dist_kind,
                           double stop_dist, int64_t max_iternum, int64_t* state, double* packed, int32_t* assign,
                           void* workspace, size_t workspace_bytes, void* const* mailboxes, int rank, int world,
                           yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (centers == nullptr || state == nullptr || workspace == nullptr) return YB_E_NULL;
    if (n_points > 0 && data == nullptr) return YB_E_NULL;
    if (n_points < 0 || k < 1 || k > 16 || n_dim < 1 || n_dim > 4) return YB_E_SHAPE;
    if (dist_kind != YB_DIST_IOU && dist_kind != YB_DIST_EUCLID) return YB_E_PARAM;
    if (dist_kind == YB_DIST_IOU && n_dim < 2) return YB_E_SHAPE;
    if (((uintptr_t)data & 7) || ((uintptr_t)state & 7)) return YB_E_ALIGN;
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
    L.sums = nullptr;
    L.counts = nullptr;
    L.partials = reinterpret_cast<double*>(workspace);
    // the last CTA of every launch leaves the counter at zero: the caller zeroes the workspace ONCE
    // (yb_kmeans_lloyd_init) and the loop needs no memset between iterations
    L.counter = reinterpret_cast<unsigned int*>(
        (char*)workspace + align_up((size_t)kKmGrid * 16 * (n_dim + 1) * sizeof(double), 256));
    L.state = reinterpret_cast<long long*>(state);
    L.packed = packed;
    L.centers_rw = (packed == nullptr) ? centers : nullptr;
    L.stop_dist = stop_dist;
    L.max_iter = max_iternum;
    for (int i = 0; i < YB_MAX_PEERS; ++i) L.mailbox[i] = nullptr;
    L.rank = rank;
    L.world = world;
    if (mailboxes != nullptr) {
        if (world < 1 || world > YB_MAX_PEERS || rank < 0 || rank >= world) return YB_E_PARAM;
        for (int i = 0; i < world; ++i) {
            if (mailboxes[i] == nullptr) return YB_E_NULL;
            L.mailbox[i] = reinterpret_cast<unsigned char*>(mailboxes[i]);
        }
    }
    switch (n_dim) {
        case 1: return launch_km_d<1>(L, stream);
        case 2: return launch_km_boxes(L, stream);
        case 3: return launch_km_d<3>(L, stream);
        default: return launch_km_d<4>(L, stream);
    }
}

extern "C" int yb_kmeans_lloyd_step(const double* data, int64_t n_points, int n_dim, double* centers, int k,
                                    int dist_kind, double stop_dist, int64_t max_iternum, int64_t* state,
                                    double* packed, int32_t* assign, void* workspace, size_t workspace_bytes,
                                    yb_stream_t stream) {
    return lloyd_step_impl(data, n_points, n_dim, centers, k, dist_kind, stop_dist, max_iternum, state, packed, assign,
                           workspace, workspace_bytes, nullptr, 0, 1, stream);
}

extern "C" int yb_kmeans_lloyd_step_peers(const double* data, int64_t n_points, int n_dim, double* centers, int k,
                                          int dist_kind, double stop_dist, int64_t max_iternum, int64_t* state,
                                          int32_t* assign, void* workspace, size_t workspace_bytes,
                                          void* const* mailboxes_host, int rank, int world, yb_stream_t stream) {
    if (mailboxes_host == nullptr) return YB_E_NULL;
    return lloyd_step_impl(data, n_points, n_dim, centers, k, dist_kind, stop_dist, max_iternum, state, nullptr, assign,
                           workspace, workspace_bytes, mailboxes_host, rank, world, stream);
}

extern "C" size_t yb_peer_mailbox_bytes(int world) {
    if (world < 1 || world > YB_MAX_PEERS) return 0;
    return align_up(2 * (size_t)world * kPeerSlotBytes, 256);
}

// The one place this library owns device memory: a mailbox must come from cudaMalloc to be exported
// to the other ranks' processes (cudaIpcGetMemHandle); the caller frees it with yb_peer_free.
extern "C" int yb_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle64) {
    if (dev_ptr == nullptr || ipc_handle64 == nullptr || bytes == 0) return YB_E_NULL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    YB_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
    YB_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    YB_CUDA_TRY(cudaIpcGetMemHandle(&h, *dev_ptr));
    memcpy(ipc_handle64, &h, sizeof(h));
    return YB_OK;
}
extern "C" int yb_peer_open(const void* ipc_handle64, void** dev_ptr) {
    if (dev_ptr == nullptr || ipc_handle64 == nullptr) return YB_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    YB_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return YB_OK;
}
extern "C" int yb_peer_close(void* dev_ptr) {
    if (dev_ptr == nullptr) return YB_E_NULL;
    YB_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return YB_OK;
}
extern "C" int yb_peer_free(void* dev_ptr) {
    if (dev_ptr == nullptr) return YB_E_NULL;
    YB_CUDA_TRY(cudaFree(dev_ptr));
    return YB_OK;
}

extern "C" int yb_kmeans_lloyd_init(int64_t* state, int k, int n_dim, void* workspace, size_t workspace_bytes,
                                    yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (state == nullptr || workspace == nullptr) return YB_E_NULL;
    const size_t sb = yb_kmeans_state_bytes(k, n_dim);
    if (sb == 0) return YB_E_SHAPE;
    if (workspace_bytes < yb_kmeans_workspace_bytes(0, k, n_dim) || ((uintptr_t)workspace & 255)) return YB_E_WORKSPACE;
    YB_CUDA_TRY(cudaMemsetAsync(state, 0, sb, stream));
    unsigned int* counter = reinterpret_cast<unsigned int*>(
        (char*)workspace + align_up((size_t)kKmGrid * 16 * (n_dim + 1) * sizeof(double), 256));
    YB_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    return YB_OK;
}

extern "C" int yb_kmeans_lloyd_update(const double* packed, double* centers, int k, int n_dim, int dist_kind,
                                      double stop_dist, int64_t max_iternum, int64_t* state, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (packed == nullptr || centers == nullptr || state == nullptr) return YB_E_NULL;
    if (k < 1 || k > 16 || n_dim < 1 || n_dim > 4) return YB_E_SHAPE;
    if (dist_kind != YB_DIST_IOU && dist_kind != YB_DIST_EUCLID) return YB_E_PARAM;
    kmeans_update_kernel<<<1, 32, 0, stream>>>(packed, centers, k, n_dim, dist_kind,
                                               reinterpret_cast<long long*>(state), stop_dist, max_iternum);
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
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
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}
