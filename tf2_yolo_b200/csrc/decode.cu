// Head decode: threshold c*p_k per (cell, box, class) and emit float64 rows
// [x, y, w, h, c, class, p] in the reference's order (utils/tools.py:370-438),
// batched over images and scales.
//
//   K1 decode_count : one warp per cell reads the cell's scores once (coalesced),
//                     counts hits with warp ballots, stores the count at the
//                     cell's position in OUTPUT order (image, scale, y, x).
//   scan            : exclusive prefix of the per-cell counts -> row offset per cell.
//   K2 decode_emit  : warps skip empty cells (32 counts per load, one ballot) and
//                     re-evaluate only cells with hits, writing rows in (box, class)
//                     order at the scanned offset.
//
// The comparison runs in the INPUT dtype (fp32 product, threshold rounded to fp32
// for fp32 heads), the row arithmetic in fp64 with one rounding per operation
// (this file is compiled with -fmad=false).
#include <cstring>

#include "common.cuh"
#include "scan.cuh"

namespace yb {

struct DecodeLaunch {
    const void* preds[YB_MAX_SCALES];
    int gh[YB_MAX_SCALES], gw[YB_MAX_SCALES], B[YB_MAX_SCALES];
    int pcf[YB_MAX_SCALES];            // values per cell
    long long cells[YB_MAX_SCALES];    // gh*gw
    long long cell_base[YB_MAX_SCALES + 1];  // prefix of cells over scales (per image)
    long long scale_base[YB_MAX_SCALES + 1]; // prefix of n_img*cells over scales
    int n_scales, C, version;
    long long n_img;
    double thr;
};

template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }

// count (and optionally emit) the hits of one cell; whole warp cooperates.
template <typename T, bool kEmit>
__device__ __forceinline__ int cell_hits(const T* __restrict__ cell, int B, int C, int version, T thr,
                                         int lane, double* __restrict__ rows, long long row0,
                                         long long cap, int xi, int yi, int gw, int gh) {
    int n = 0;
    const int bstride = (version == 1) ? 5 : 5 + C;
    for (int b = 0; b < B; ++b) {
        const T* box = cell + b * bstride;
        const T c = box[4];
        const T* prob = (version == 1) ? cell + 5 * B : box + 5;
        for (int k0 = 0; k0 < C; k0 += 32) {
            const int k = k0 + lane;
            T p = 0;
            bool hit = false;
            if (k < C) {
                p = prob[k];
                hit = mul_rn<T>(c, p) >= thr;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (kEmit && hit) {
                const long long r = row0 + n + __popc(m & ((1u << lane) - 1u));
                if (r < cap) {
                    double* o = rows + r * 7;
                    o[0] = ((double)xi + (double)box[0]) / (double)gw;
                    o[1] = ((double)yi + (double)box[1]) / (double)gh;
                    o[2] = (double)box[2];
                    o[3] = (double)box[3];
                    o[4] = (double)c;
                    o[5] = (double)k;
                    o[6] = (double)p;
                }
            }
            n += __popc(m);
        }
    }
    return n;
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_count_kernel(const __grid_constant__ DecodeLaunch L, unsigned int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = L.scale_base[L.n_scales];
    const T thr = (T)L.thr;
    for (long long g = warp; g < total; g += n_warps) {
        int s = 0;
        while (g >= L.scale_base[s + 1]) ++s;
        const long long local = g - L.scale_base[s];  // img * cells + cell, memory order
        const long long img = local / L.cells[s];
        const long long cell = local - img * L.cells[s];
        const T* ptr = reinterpret_cast<const T*>(L.preds[s]) + local * L.pcf[s];
        const int n = cell_hits<T, false>(ptr, L.B[s], L.C, L.version, thr, lane, nullptr, 0, 0, 0, 0, 1, 1);
        if (lane == 0) counts[img * L.cell_base[L.n_scales] + L.cell_base[s] + cell] = (unsigned)n;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_emit_kernel(const __grid_constant__ DecodeLaunch L, const unsigned int* __restrict__ counts,
                   const long long* __restrict__ offsets, double* __restrict__ rows, long long cap,
                   long long* __restrict__ row_offsets) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long per_img = L.cell_base[L.n_scales];
    const long long total = L.n_img * per_img;
    const T thr = (T)L.thr;
    // per-image extents
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= L.n_img;
         i += (long long)gridDim.x * blockDim.x)
        row_offsets[i] = offsets[i * per_img];
    for (long long base = warp * 32; base < total; base += n_warps * 32) {
        const long long idx = base + lane;
        const unsigned cnt = (idx < total) ? counts[idx] : 0u;
        unsigned m = __ballot_sync(0xffffffffu, cnt != 0u);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const long long o = base + src;  // output-order cell index
            const long long img = o / per_img;
            const long long rem = o - img * per_img;
            int s = 0;
            while (rem >= L.cell_base[s + 1]) ++s;
            const long long cell = rem - L.cell_base[s];
            const int yi = (int)(cell / L.gw[s]), xi = (int)(cell - (long long)yi * L.gw[s]);
            const T* ptr = reinterpret_cast<const T*>(L.preds[s]) + (img * L.cells[s] + cell) * L.pcf[s];
            cell_hits<T, true>(ptr, L.B[s], L.C, L.version, thr, lane, rows, offsets[o], cap, xi, yi,
                               L.gw[s], L.gh[s]);
        }
    }
}

static int fill_decode(const void* const* preds, int64_t n_img, const yb_decode_params* p, DecodeLaunch& L) {
    if (preds == nullptr || p == nullptr) return YB_E_NULL;
    if (p->version < 1 || p->version > 4) return YB_E_PARAM;
    if (p->n_scales < 1 || p->n_scales > YB_MAX_SCALES || p->class_num <= 0 || n_img < 0) return YB_E_SHAPE;
    memset(&L, 0, sizeof(L));
    L.n_scales = p->n_scales;
    L.C = p->class_num;
    L.version = p->version;
    L.n_img = n_img;
    L.thr = p->threshold;
    const size_t esz = p->is_f64 ? 8 : 4;
    for (int s = 0; s < p->n_scales; ++s) {
        if (preds[s] == nullptr) return YB_E_NULL;
        if ((uintptr_t)preds[s] & (esz - 1)) return YB_E_ALIGN;
        if (p->grid_h[s] <= 0 || p->grid_w[s] <= 0 || p->bbox_num[s] <= 0) return YB_E_SHAPE;
        L.preds[s] = preds[s];
        L.gh[s] = p->grid_h[s];
        L.gw[s] = p->grid_w[s];
        L.B[s] = p->bbox_num[s];
        L.pcf[s] = (p->version == 1) ? 5 * p->bbox_num[s] + p->class_num
                                     : p->bbox_num[s] * (5 + p->class_num);
        L.cells[s] = (long long)p->grid_h[s] * p->grid_w[s];
        L.cell_base[s + 1] = L.cell_base[s] + L.cells[s];
        L.scale_base[s + 1] = L.scale_base[s] + n_img * L.cells[s];
    }
    return YB_OK;
}

}  // namespace yb

using namespace yb;

static size_t decode_counts_bytes(long long total_cells) {
    return align_up((size_t)(total_cells + 1) * sizeof(unsigned int), 256);
}
static size_t decode_offsets_bytes(long long total_cells) {
    return align_up((size_t)(total_cells + 2) * sizeof(long long), 256);
}

extern "C" size_t yb_decode_workspace_bytes(const yb_decode_params* p, int64_t n_img) {
    if (p == nullptr || n_img < 0) return 0;
    long long per_img = 0;
    for (int s = 0; s < p->n_scales && s < YB_MAX_SCALES; ++s)
        per_img += (long long)p->grid_h[s] * p->grid_w[s];
    const long long total = per_img * n_img;
    return decode_counts_bytes(total) + decode_offsets_bytes(total) + scan_workspace_bytes(total > 0 ? total : 1);
}

extern "C" int yb_decode(const void* const* preds, int64_t n_img, const yb_decode_params* p, double* rows,
                         int64_t row_capacity, int64_t* row_offsets, void* workspace,
                         size_t workspace_bytes, yb_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    DecodeLaunch L;
    int rc = fill_decode(preds, n_img, p, L);
    if (rc != YB_OK) return rc;
    if (row_offsets == nullptr || workspace == nullptr) return YB_E_NULL;
    if (rows == nullptr && row_capacity > 0) return YB_E_NULL;
    if (row_capacity < 0) return YB_E_CAPACITY;
    if (workspace_bytes < yb_decode_workspace_bytes(p, n_img) || ((uintptr_t)workspace & 255))
        return YB_E_WORKSPACE;
    const long long total = L.n_img * L.cell_base[L.n_scales];
    if (total == 0) {
        YB_CUDA_TRY(cudaMemsetAsync(row_offsets, 0, sizeof(int64_t) * (n_img + 1), stream));
        return YB_OK;
    }
    unsigned int* counts = reinterpret_cast<unsigned int*>(workspace);
    long long* offsets = reinterpret_cast<long long*>((char*)workspace + decode_counts_bytes(total));
    void* scan_ws = (char*)offsets + decode_offsets_bytes(total);

    const int threads = 256;
    const long long warps_needed = total;
    int blocks = (int)min((long long)kNumSMs * 8, (warps_needed * 32 + threads - 1) / threads);
    if (p->is_f64)
        decode_count_kernel<double><<<blocks, threads, 0, stream>>>(L, counts);
    else
        decode_count_kernel<float><<<blocks, threads, 0, stream>>>(L, counts);
    YB_CUDA_TRY(cudaGetLastError());
    rc = exclusive_scan_u32(counts, total, offsets, scan_ws, stream);
    if (rc != 0) return rc;
    int blocks2 = (int)min((long long)kNumSMs * 8, (total + threads - 1) / threads);
    if (blocks2 < 1) blocks2 = 1;
    if (p->is_f64)
        decode_emit_kernel<double><<<blocks2, threads, 0, stream>>>(
            L, counts, offsets, rows, row_capacity, reinterpret_cast<long long*>(row_offsets));
    else
        decode_emit_kernel<float><<<blocks2, threads, 0, stream>>>(
            L, counts, offsets, rows, row_capacity, reinterpret_cast<long long*>(row_offsets));
    return (int)cudaGetLastError();
}
