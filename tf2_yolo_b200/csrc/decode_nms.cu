// Head decode + per-class NMS of a batch in ONE launch, one CTA per image.
//
// Replaces the chain  yb_decode (scan + emit)  ->  yb_nms (classify, scatter, sweep, emit)  -
// six latency-bound launches and three memsets behind the counting pass - for the regime the
// train-and-evaluate step and per-image inference live in: at most a few thousand rows per image
// (utils/tools.py:370-438 decode, :687-733 nms with the IoU of :630-684).  The decode rows of an
// image never leave the SM: the counting pass (decode_count_kernel or the fused loss kernel) files
// the boxes with hits into per-image buckets; the image's CTA orders them (row-major cell, box),
// re-evaluates their class scores, builds the float64 rows [x, y, w, h, c, class, p] in shared
// memory in the reference's order, groups them by class, runs the greedy (D)IoU-NMS of every class
// on one warp each (visit order = confidence descending, equal confidences -> higher original
// index first; suppression on >=; division-free pair test with the pinned exact fallback, see
// nms_pair.cuh) and writes the survivors class-major, original order inside a class, at the
// image's offset in the compact output - found with a decoupled look-back over the earlier images
// (tickets make every predecessor of a CTA already running, so the chain cannot stall).
// Results are bit-identical to the six-launch chain.  An image with more rows than the caller's
// rows_per_img_cap is not processed: it contributes no rows and is counted in *n_overflow; the
// caller falls back to yb_decode + yb_nms (same contract as the row capacity of yb_decode).
#include <climits>
#include <cstring>

#include "common.cuh"
#include "decode_internal.cuh"
#include "nms_pair.cuh"

namespace yb {

constexpr int kFusedThreads = 512;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxClassWords = 8;   // class_num <= 256

constexpr unsigned long long kStFlagAgg = 1ull << 62, kStFlagIncl = 2ull << 62, kStFlagMask = 3ull << 62;
constexpr int kStOvfShift = 40;          // payload: survivors in bits 0..39, overflowed images in 40..61

struct FusedLaunch {
    DecodeLaunch D;            // float32 heads
    HotBuckets K;
    int row_cap;               // rows (and hot boxes) per image held in shared memory
    int cw;                    // ceil(C / 32)
    double nms_thr;
    double* out_rows;          // compact survivors (may be mapped host memory)
    long long out_cap;
    long long* out_offsets;    // [n_img + 1]
    unsigned int* n_overflow;  // [1] or null
    unsigned long long* status;   // [n_img] look-back words, zero on entry
    unsigned int* ticket;         // [1], zero on entry
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// shared-memory carve-up for R = row_cap rows and C classes (host and device agree through this)
struct FusedSmem {
    size_t rows, key, mem, ord, meta, cnt, off, mask, cls, member, vis, vrank, rank, outsrc, ccount, cstart, ckept, total;
    __host__ __device__ FusedSmem(int R, int C, int cw) {
        size_t o = 0;
        auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
        rows = take(sizeof(double) * 7 * R);
        key = take(4 * (size_t)R);
        mem = take(4 * (size_t)R);
        mask = take(4 * (size_t)R * cw);
        off = take(4 * ((size_t)R + 1));
        ord = take(2 * (size_t)R);
        meta = take(2 * (size_t)R);
        cnt = take(2 * (size_t)R);
        cls = take(2 * (size_t)R);
        member = take(2 * (size_t)R);
        vis = take(2 * (size_t)R);
        vrank = take(2 * (size_t)R);
        rank = take(2 * (size_t)R);
        outsrc = take(2 * (size_t)R);
        ccount = take(4 * (size_t)C);
        cstart = take(4 * ((size_t)C + 1));
        ckept = take(4 * ((size_t)C + 1));
        total = o;
    }
};

template <int MODE>
__global__ void __launch_bounds__(kFusedThreads)
decode_nms_image_kernel(const __grid_constant__ FusedLaunch F) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const DecodeLaunch& L = F.D;
    const int R = F.row_cap, C = L.C, CW = F.cw;
    const FusedSmem lay(R, C, CW);
    double* s_rows = reinterpret_cast<double*>(fsm + lay.rows);
    unsigned int* s_key = reinterpret_cast<unsigned int*>(fsm + lay.key);
    unsigned int* s_mem = reinterpret_cast<unsigned int*>(fsm + lay.mem);
    unsigned int* s_mask = reinterpret_cast<unsigned int*>(fsm + lay.mask);
    unsigned int* s_off = reinterpret_cast<unsigned int*>(fsm + lay.off);
    unsigned short* s_ord = reinterpret_cast<unsigned short*>(fsm + lay.ord);
    unsigned short* s_meta = reinterpret_cast<unsigned short*>(fsm + lay.meta);
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(fsm + lay.cnt);
    unsigned short* s_cls = reinterpret_cast<unsigned short*>(fsm + lay.cls);
    unsigned short* s_member = reinterpret_cast<unsigned short*>(fsm + lay.member);
    unsigned short* s_vis = reinterpret_cast<unsigned short*>(fsm + lay.vis);
    unsigned short* s_vrank = reinterpret_cast<unsigned short*>(fsm + lay.vrank);
    unsigned short* s_rank = reinterpret_cast<unsigned short*>(fsm + lay.rank);
    unsigned short* s_outsrc = reinterpret_cast<unsigned short*>(fsm + lay.outsrc);
    unsigned int* s_ccount = reinterpret_cast<unsigned int*>(fsm + lay.ccount);
    unsigned int* s_cstart = reinterpret_cast<unsigned int*>(fsm + lay.cstart);
    unsigned int* s_ckept = reinterpret_cast<unsigned int*>(fsm + lay.ckept);
    __shared__ unsigned int s_img, s_warp_tot[kFusedWarps], s_n_rows, s_kept;
    __shared__ long long s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (tid == 0) s_img = atomicAdd(F.ticket, 1u);   // images in ticket order: predecessors are running
    for (int c = tid; c < C; c += kFusedThreads) s_ccount[c] = 0u;
    __syncthreads();
    const long long img = s_img;
    const long long per_img = L.cell_base[L.n_scales];
    const float thr = (float)L.thr;

    // ---- 1. the image's boxes with hits -----------------------------------------------------------
    const unsigned n_filed = F.K.n[img];
    bool overflow = n_filed > (unsigned)min(R, F.K.cap);
    const int nh = overflow ? 0 : (int)n_filed;
    for (int i = tid; i < nh; i += kFusedThreads) {
        const HotBox hb = F.K.box[img * F.K.cap + i];
        const unsigned box = (hb.packed >> 4) & 63u;
        s_key[i] = (unsigned)(hb.out_idx - img * per_img) * 32u + box;   // (scale, y, x, box): the output order
        s_mem[i] = hb.mem_idx;
        s_meta[i] = (unsigned short)((hb.packed & 15u) | (box << 4));
    }
    __syncthreads();
    // ---- 2. order them (keys are unique: rank = number of smaller keys) ----------------------------
    for (int i = tid; i < nh; i += kFusedThreads) {
        const unsigned k = s_key[i];
        int r = 0;
        for (int j = 0; j < nh; ++j) r += (s_key[j] < k) ? 1 : 0;
        s_ord[r] = (unsigned short)i;
    }
    __syncthreads();
    // ---- 3. class scores of every box: hit masks and counts (utils/tools.py:411-412) ---------------
    auto box_ptrs = [&](int q, const float*& box, const float*& prob, int& s) {
        const int e = s_ord[q];
        s = s_meta[e] & 15;
        const int b = s_meta[e] >> 4;
        const float* cell = reinterpret_cast<const float*>(L.preds[s]) + (size_t)s_mem[e] * L.pcf[s];
        box = cell + b * ((L.version == 1) ? 5 : 5 + C);
        prob = (L.version == 1) ? cell + 5 * L.B[s] : box + 5;
    };
    // Several boxes per warp iteration, every global load of all of them issued before the first use: the
    // phase is a chain of HBM round trips (the head tensors were streamed long ago), so what counts
    // is how many loads are in flight, not how few instructions run.
    auto load_scores = [&](int q, float& c, float (&pv)[kFusedMaxClassWords]) {
        const float *box, *prob;
        int s;
        box_ptrs(q, box, prob, s);
#pragma unroll
        for (int w = 0; w < kFusedMaxClassWords; ++w) {
            const int k = w * 32 + lane;
            pv[w] = (w < CW && k < C) ? __ldg(prob + k) : 0.f;
        }
        c = __ldg(box + 4);
    };
    auto count_hits = [&](int q, float c, const float (&pv)[kFusedMaxClassWords]) {
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < kFusedMaxClassWords; ++w) {
            if (w < CW) {
                const int k = w * 32 + lane;
                const bool hit = (k < C) && (__fmul_rn(c, pv[w]) >= thr);
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_mask[q * CW + w] = m;
                cnt += __popc(m);
            }
        }
        if (lane == 0) s_cnt[q] = (unsigned short)min(cnt, 65535);
    };
    for (int q0 = warp; q0 < nh; q0 += 4 * kFusedWarps) {     // four boxes of a warp in flight
        const int q1 = q0 + kFusedWarps, q2 = q1 + kFusedWarps, q3 = q2 + kFusedWarps;
        float c0, c1 = 0.f, c2 = 0.f, c3 = 0.f;
        float pv0[kFusedMaxClassWords], pv1[kFusedMaxClassWords], pv2[kFusedMaxClassWords], pv3[kFusedMaxClassWords];
        load_scores(q0, c0, pv0);
        if (q1 < nh) load_scores(q1, c1, pv1);
        if (q2 < nh) load_scores(q2, c2, pv2);
        if (q3 < nh) load_scores(q3, c3, pv3);
        count_hits(q0, c0, pv0);
        if (q1 < nh) count_hits(q1, c1, pv1);
        if (q2 < nh) count_hits(q2, c2, pv2);
        if (q3 < nh) count_hits(q3, c3, pv3);
    }
    __syncthreads();
    // ---- exclusive scan of the counts -> first row of every box ------------------------------------
    {
        const int ipt = (nh + kFusedThreads - 1) / kFusedThreads;   // consecutive items per thread
        const int i0 = tid * ipt;
        unsigned v = 0;
        for (int i = i0; i < min(nh, i0 + ipt); ++i) v += s_cnt[i];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp_tot[warp] = inc;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp_tot[w];
        unsigned ex = woff + inc - v;
        for (int i = i0; i < min(nh, i0 + ipt); ++i) {
            s_off[i] = ex;
            ex += s_cnt[i];
        }
        if (tid == kFusedThreads - 1) s_n_rows = woff + inc;
        __syncthreads();
    }
    if (s_n_rows > (unsigned)R) overflow = true;
    const int n_rows = overflow ? 0 : (int)s_n_rows;

    // ---- 4. the rows, in the reference's order (utils/tools.py:414-436) ----------------------------
    if (!overflow) {
        for (int q = warp; q < nh; q += kFusedWarps) {
            const float *box, *prob;
            int s;
            box_ptrs(q, box, prob, s);
            float pv[kFusedMaxClassWords];
#pragma unroll
            for (int w = 0; w < kFusedMaxClassWords; ++w) {
                const int k = w * 32 + lane;
                pv[w] = (w < CW && k < C) ? __ldg(prob + k) : 0.f;   // L1 / L2 hits: read a moment ago
            }
            const float f0 = __ldg(box), f1 = __ldg(box + 1), f2 = __ldg(box + 2), f3 = __ldg(box + 3), c = __ldg(box + 4);
            const unsigned cell = s_mem[s_ord[q]] % (unsigned)L.cells[s];
            const int yi = (int)(cell / (unsigned)L.gw[s]), xi = (int)(cell - (unsigned)yi * (unsigned)L.gw[s]);
            const double bx = ((double)xi + (double)f0) / (double)L.gw[s];
            const double by = ((double)yi + (double)f1) / (double)L.gh[s];
            const double bw = (double)f2, bh = (double)f3, bc = (double)c;
            unsigned r = s_off[q];
#pragma unroll
            for (int w = 0; w < kFusedMaxClassWords; ++w) {
                if (w < CW) {
                    const unsigned m = s_mask[q * CW + w];
                    if ((m >> lane) & 1u) {
                        const int k = w * 32 + lane;
                        const unsigned at = r + __popc(m & lt_mask);
                        double* o = s_rows + (size_t)at * 7;
                        o[0] = bx; o[1] = by; o[2] = bw; o[3] = bh; o[4] = bc;
                        o[5] = (double)k;
                        o[6] = (double)pv[w];
                        s_cls[at] = (unsigned short)k;
                        atomicAdd(&s_ccount[k], 1u);
                    }
                    r += __popc(m);
                }
            }
        }
    }
    __syncthreads();
    // ---- 5. class segments: starts (exclusive scan over C), members in original order --------------
    for (int c = tid; c <= C; c += kFusedThreads) {
        unsigned sum = 0;
        for (int j = 0; j < c; ++j) sum += s_ccount[j];
        s_cstart[c] = sum;
        if (c < C) s_ckept[c] = 0u;
    }
    __syncthreads();
    const bool pos_thr = F.nms_thr > 0.0;
    const double nms_thr = F.nms_thr;
    // ---- 6. one warp per class: stable member list, visit order, greedy sweep ----------------------
    for (int c = warp; c < C; c += kFusedWarps) {
        const int n = (int)s_ccount[c];
        if (n == 0) continue;
        const int start = (int)s_cstart[c];
        {
            int pos = start;
            for (int r0 = 0; r0 < n_rows; r0 += 32) {
                const int r = r0 + lane;
                const bool is = r < n_rows && s_cls[r] == c;
                const unsigned m = __ballot_sync(0xffffffffu, is);
                if (is) s_member[pos + __popc(m & lt_mask)] = (unsigned short)r;
                pos += __popc(m);
            }
        }
        __syncwarp();
        // visit rank (np.argsort(conf)[::-1], utils/tools.py:716-717)
        for (int i = lane; i < n; i += 32) {
            const double* ri = s_rows + (size_t)s_member[start + i] * 7;
            const double ci = __dmul_rn(ri[4], ri[6]);
            int vis = 0;
            for (int j = 0; j < n; ++j) {
                const double* rj = s_rows + (size_t)s_member[start + j] * 7;
                vis += (j != i && visited_before(__dmul_rn(rj[4], rj[6]), j, ci, i)) ? 1 : 0;
            }
            s_vis[start + vis] = (unsigned short)i;
            s_vrank[start + i] = (unsigned short)vis;
        }
        __syncwarp();
        // greedy sweep: lane l owns members l, l+32, ...; dead bit t of a lane = member l + 32 t
        unsigned dead = 0;
        for (int v = 0; v < n; ++v) {
            const int iv = s_vis[start + v];
            const unsigned dv = __shfl_sync(0xffffffffu, dead, iv & 31);
            if ((dv >> (iv >> 5)) & 1u) continue;            // a suppressed box suppresses nothing (:723)
            const int rv = s_member[start + iv];
            const double* a = s_rows + (size_t)rv * 7;
            const BoxC av = make_box(a[0], a[1], a[2], a[3]);
            for (int i = lane, t = 0; i < n; i += 32, ++t) {
                if ((dead >> t) & 1u) continue;
                if ((int)s_vrank[start + i] <= v) continue;   // already visited (white list, :722)
                const int ri = s_member[start + i];
                const double* b = s_rows + (size_t)ri * 7;
                const BoxC bi = make_box(b[0], b[1], b[2], b[3]);
                if (suppresses<MODE>(av, bi, nms_thr, pos_thr, s_rows, rv, ri)) dead |= 1u << t;
            }
        }
        // survivors: position inside the class, original order
        int kept_before = 0;
        for (int i0 = 0, t = 0; i0 < n; i0 += 32, ++t) {
            const int i = i0 + lane;
            const bool alive = i < n && !((dead >> t) & 1u);
            const unsigned m = __ballot_sync(0xffffffffu, alive);
            if (i < n) s_rank[s_member[start + i]] = alive ? (unsigned short)(kept_before + __popc(m & lt_mask)) : 0xffffu;
            kept_before += __popc(m);
        }
        if (lane == 0) s_ckept[c] = (unsigned)kept_before;
    }
    __syncthreads();
    // ---- 7. survivors of the image: class-major positions; the image's offset by look-back ---------
    if (tid == 0) {
        unsigned sum = 0;
        for (int c = 0; c < C; ++c) {
            const unsigned k = s_ckept[c];
            s_ckept[c] = sum;          // exclusive prefix of the kept counts
            sum += k;
        }
        s_kept = sum;
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned long long agg = (unsigned long long)s_kept | ((unsigned long long)(overflow ? 1 : 0) << kStOvfShift);
        unsigned long long excl = 0;
        if (img > 0) {
            if (lane == 0) st_release_u64(&F.status[img], kStFlagAgg | agg);
            long long j = img - 1;
            for (;;) {
                const long long idx = j - lane;
                unsigned long long st = kStFlagIncl;   // before image 0: inclusive prefix 0
                if (idx >= 0) {
                    st = ld_acquire_u64(&F.status[idx]);
                    while ((st & kStFlagMask) == 0) {
                        __nanosleep(40);
                        st = ld_acquire_u64(&F.status[idx]);
                    }
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (st & kStFlagMask) == kStFlagIncl);
                const int stop = incl ? (__ffs(incl) - 1) : 31;   // nearest predecessor with an inclusive prefix
                unsigned long long part = (lane <= stop) ? (st & ~kStFlagMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                if (incl) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st_release_u64(&F.status[img], kStFlagIncl | (excl + agg));
            const long long base = (long long)(excl & ((1ull << kStOvfShift) - 1));
            s_base = base;
            F.out_offsets[img] = base;
            if (img == L.n_img - 1) {
                F.out_offsets[L.n_img] = base + s_kept;
                if (F.n_overflow != nullptr) *F.n_overflow = (unsigned)((excl + agg) >> kStOvfShift);
            }
        }
    }
    // source row of every output position, then a coalesced copy
    for (int r = tid; r < n_rows; r += kFusedThreads)
        if (s_rank[r] != 0xffffu) s_outsrc[s_ckept[s_cls[r]] + s_rank[r]] = (unsigned short)r;
    __syncthreads();
    if (F.out_rows == nullptr) return;
    const long long base = s_base;
    const int n_out = (int)s_kept;
    for (int f = tid; f < n_out * 7; f += kFusedThreads) {
        const int q = f / 7, k = f - q * 7;
        if (base + q < F.out_cap) F.out_rows[(base + q) * 7 + k] = s_rows[(size_t)s_outsrc[q] * 7 + k];
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct FusedWs {
    unsigned char* zero;          // [bucket counts | look-back status | ticket]: cleared by ONE memset
    size_t zero_bytes;
    HotBuckets K;
    unsigned long long* status;
    unsigned int* ticket;
};

static size_t fused_layout(long long n_img, int row_cap, FusedWs* W, char* base) {
    size_t off = 0;
    const size_t counts = align_up(sizeof(unsigned int) * (size_t)(n_img + 1), 256);
    const size_t status = align_up(sizeof(unsigned long long) * (size_t)(n_img + 1), 256);
    if (W) {
        W->zero = reinterpret_cast<unsigned char*>(base);
        W->K.n = reinterpret_cast<unsigned int*>(base);
        W->status = reinterpret_cast<unsigned long long*>(base + counts);
        W->ticket = reinterpret_cast<unsigned int*>(base + counts + status);
        W->zero_bytes = counts + status + 256;
        W->K.box = reinterpret_cast<HotBox*>(base + counts + status + 256);
        W->K.cap = row_cap;
    }
    off = counts + status + 256 + align_up(sizeof(HotBox) * (size_t)n_img * (size_t)row_cap, 256);
    return off;
}

static int fused_check(const yb_decode_params* p, int64_t n_img, int row_cap) {
    if (p == nullptr) return YB_E_NULL;
    if (p->is_f64) return YB_E_PARAM;                       // model outputs (float32) only
    if (n_img < 0 || n_img > (1 << 24)) return YB_E_SHAPE;
    if (p->class_num < 1 || p->class_num > 32 * kFusedMaxClassWords) return YB_E_SHAPE;
    if (row_cap < 32 || row_cap > YB_FUSED_MAX_ROWS) return YB_E_SHAPE;
    long long per_img = 0;
    for (int s = 0; s < p->n_scales && s < YB_MAX_SCALES; ++s) per_img += (long long)p->grid_h[s] * p->grid_w[s];
    if (per_img >= (1ll << 27)) return YB_E_SHAPE;          // (cell, box) key in 32 bits
    return YB_OK;
}

// the kernel launch alone (buckets already filled by a counting pass on the same stream)
static int fused_launch(const DecodeLaunch& D, const FusedWs& W, int row_cap, double nms_threshold, int iou_mode,
                        double* out_rows, long long out_cap, long long* out_offsets, unsigned int* n_overflow,
                        cudaStream_t stream) {
    FusedLaunch F;
    memset(&F, 0, sizeof(F));
    F.D = D;
    F.K = W.K;
    F.row_cap = row_cap;
    F.cw = (D.C + 31) / 32;
    F.nms_thr = nms_threshold;
    F.out_rows = out_rows;
    F.out_cap = out_cap;
    F.out_offsets = out_offsets;
    F.n_overflow = n_overflow;
    F.status = W.status;
    F.ticket = W.ticket;
    const FusedSmem lay(row_cap, D.C, F.cw);
    if (lay.total > 220 * 1024) return YB_E_SHAPE;
    if (iou_mode == 1) {
        static SmemRaised done;
        YB_CUDA_TRY(raise_dynamic_smem_once(decode_nms_image_kernel<1>, (int)lay.total, &done));
        decode_nms_image_kernel<1><<<(unsigned)D.n_img, kFusedThreads, lay.total, stream>>>(F);
    } else {
        static SmemRaised done;
        YB_CUDA_TRY(raise_dynamic_smem_once(decode_nms_image_kernel<2>, (int)lay.total, &done));
        decode_nms_image_kernel<2><<<(unsigned)D.n_img, kFusedThreads, lay.total, stream>>>(F);
    }
    YB_CUDA_TRY(cudaGetLastError());
    return YB_OK;
}

// Shared with loss.cu (yb_loss_decode_nms_fused): workspace carve-up + memset, then the launch.
int fused_prepare(const void* const* preds, int64_t n_img, const yb_decode_params* p, int row_cap, void* workspace,
                  size_t workspace_bytes, DecodeLaunch& D, HotBuckets& K, cudaStream_t stream) {
    int rc = fused_check(p, n_img, row_cap);
    if (rc != YB_OK) return rc;
    rc = decode_fill(preds, n_img, p, D);
    if (rc != YB_OK) return rc;
    if (workspace == nullptr) return YB_E_NULL;
    if (workspace_bytes < fused_layout(n_img, row_cap, nullptr, nullptr) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    for (int s = 0; s < D.n_scales; ++s)
        if (D.n_img * D.cells[s] > 0xffffffffll || D.B[s] > 32) return YB_E_SHAPE;
    FusedWs W;
    fused_layout(n_img, row_cap, &W, reinterpret_cast<char*>(workspace));
    YB_CUDA_TRY(cudaMemsetAsync(W.zero, 0, W.zero_bytes, stream));
    K = W.K;
    return YB_OK;
}

int fused_finish(const DecodeLaunch& D, int64_t n_img, int row_cap, void* workspace, double nms_threshold,
                 int iou_mode, double* out_rows, int64_t out_capacity, int64_t* out_offsets,
                 unsigned int* n_overflow, cudaStream_t stream) {
    FusedWs W;
    fused_layout(n_img, row_cap, &W, reinterpret_cast<char*>(workspace));
    return fused_launch(D, W, row_cap, nms_threshold, iou_mode, out_rows, out_capacity,
                        reinterpret_cast<long long*>(out_offsets), n_overflow, stream);
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_decode_nms_workspace_bytes(const yb_decode_params* p, int64_t n_img, int rows_per_img_cap) {
    if (fused_check(p, n_img, rows_per_img_cap) != YB_OK) return 0;
    return fused_layout(n_img, rows_per_img_cap, nullptr, nullptr);
}

extern "C" int yb_decode_nms(const void* const* preds, int64_t n_img, const yb_decode_params* p,
                             double nms_threshold, int iou_mode, int rows_per_img_cap, double* out_rows,
                             int64_t out_capacity, int64_t* out_offsets, unsigned int* n_overflow,
                             void* workspace, size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (out_offsets == nullptr) return YB_E_NULL;
    if (out_rows == nullptr && out_capacity > 0) return YB_E_NULL;
    if (out_capacity < 0) return YB_E_CAPACITY;
    if (iou_mode != 1 && iou_mode != 2) return YB_E_PARAM;
    if (n_img == 0) {
        int rc0 = fused_check(p, n_img, rows_per_img_cap);
        if (rc0 != YB_OK) return rc0;
        YB_CUDA_TRY(cudaMemsetAsync(out_offsets, 0, sizeof(int64_t), stream));
        if (n_overflow != nullptr) YB_CUDA_TRY(cudaMemsetAsync(n_overflow, 0, sizeof(unsigned int), stream));
        return YB_OK;
    }
    DecodeLaunch D;
    HotBuckets K;
    int rc = fused_prepare(preds, n_img, p, rows_per_img_cap, workspace, workspace_bytes, D, K, stream);
    if (rc != YB_OK) return rc;
    rc = decode_count(D, false, nullptr, nullptr, nullptr, K, stream);
    if (rc != YB_OK) return rc;
    return fused_finish(D, n_img, rows_per_img_cap, workspace, nms_threshold, iou_mode, out_rows, out_capacity,
                        out_offsets, n_overflow, stream);
}

// Second half of a split step: the per-image buckets are already in `workspace` (left there by
// yb_loss_decode_nms_fused called with out_offsets == NULL).
extern "C" int yb_decode_nms_finish(const void* const* preds, int64_t n_img, const yb_decode_params* p,
                                    double nms_threshold, int iou_mode, int rows_per_img_cap, double* out_rows,
                                    int64_t out_capacity, int64_t* out_offsets, unsigned int* n_overflow,
                                    void* workspace, size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (out_offsets == nullptr || workspace == nullptr) return YB_E_NULL;
    if (out_rows == nullptr && out_capacity > 0) return YB_E_NULL;
    if (out_capacity < 0) return YB_E_CAPACITY;
    if (iou_mode != 1 && iou_mode != 2) return YB_E_PARAM;
    int rc = fused_check(p, n_img, rows_per_img_cap);
    if (rc != YB_OK) return rc;
    if (workspace_bytes < fused_layout(n_img, rows_per_img_cap, nullptr, nullptr) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    DecodeLaunch D;
    rc = decode_fill(preds, n_img, p, D);
    if (rc != YB_OK) return rc;
    if (n_img == 0) return YB_OK;
    return fused_finish(D, n_img, rows_per_img_cap, workspace, nms_threshold, iou_mode, out_rows, out_capacity,
                        out_offsets, n_overflow, stream);
}
