// Internal (not part of the C ABI): pieces of the decode pipeline shared with the fused
// loss+decode launch in loss.cu.
#pragma once
#include "common.cuh"

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
    // K1 tiling
    int tile_cells[YB_MAX_SCALES];
    int tile_base[YB_MAX_SCALES + 1];
    int bulk_ok[YB_MAX_SCALES];
    int stage_bytes;
};

// Work-list entry for a BOX with hits: everything the emit pass needs, no 64-bit divisions.
struct __align__(16) HotBox {
    long long out_idx;     // the cell's position in OUTPUT order (image, scale, y, x): index into offsets
    unsigned int mem_idx;  // the cell's position inside its scale tensor: img * cells + cell
    unsigned int packed;   // bits 0..3 scale, 4..9 box, 10..31 rows of the cell's earlier boxes
};
__host__ __device__ inline HotBox make_hot(long long o, unsigned mem_idx, int scale, int box, unsigned rows_before) {
    HotBox h;
    h.out_idx = o;
    h.mem_idx = mem_idx;
    h.packed = (unsigned)scale | ((unsigned)box << 4) | (rows_before << 10);
    return h;
}

// Per-image row buckets for the one-CTA-per-image decode + NMS kernel (decode_nms.cu): the
// counting pass files every (box, class) hit of image i - the values it has in registers and
// shared memory anyway - at row[i*cap + slot], slots handed out by an atomic on n[i], so the
// per-image kernel reads a few KB of compact rows instead of gathering the hot cells back from
// the head tensors.  n[i] > cap = overflow.  key orders the rows like the reference's decode:
// ((cell position in output order (scale, y, x)) * 32 + box) * 256 + class.
struct __align__(16) FusedRow {
    unsigned int key;
    float x, y, w, h, c, p;    // the head's values for the box (cell-relative x, y) and the class score
    unsigned int cell;         // scale << 28 | cell index inside the scale (y * grid_w + x)
};
struct HotBuckets {
    unsigned int* n = nullptr;   // [n_img], zero before the counting pass; nullptr = not collected
    FusedRow* row = nullptr;     // [n_img][cap]
    int cap = 0;
};
__device__ __forceinline__ unsigned fused_key(unsigned cellpos, int box, int cls) {
    return ((cellpos * 32u + (unsigned)box) << 8) | (unsigned)cls;
}

struct DecodeWs {          // carved out of the caller's decode workspace
    unsigned int* n_hot;   // number of boxes with hits
    unsigned int* counts;  // hits per cell, OUTPUT order (image, scale, y, x)
    long long* offsets;    // exclusive scan of counts (+ total)
    HotBox* hot;           // work list of boxes with hits
    void* scan_ws;
    size_t zero_bytes;     // n_hot + scan status: cleared by one memset before the counting pass
    long long total_cells;
    HotBuckets buckets;    // optional second destination of the counting pass
};

// fill L from the public parameter struct (no workspace); YB_OK / YB_E_*
int decode_fill(const void* const* preds, int64_t n_img, const yb_decode_params* p, DecodeLaunch& L);
// counting pass alone: per-cell counts + flat work list (either may be null) and / or per-image buckets
int decode_count(DecodeLaunch& L, bool is_f64, unsigned int* counts, unsigned int* n_hot, HotBox* hot,
                 const HotBuckets& buckets, cudaStream_t stream);
// validate + fill L and ws (no launches).  Returns YB_OK / YB_E_*.
int decode_setup(const void* const* preds, int64_t n_img, const yb_decode_params* p, void* workspace,
                 size_t workspace_bytes, DecodeLaunch& L, DecodeWs& ws);
// scan of the per-cell counts + emission of the rows (after counts / hot list are complete)
int decode_finish(const DecodeLaunch& L, const DecodeWs& ws, bool is_f64, double* rows, long long cap,
                  long long* row_offsets, cudaStream_t stream, bool status_zeroed = false);

// One-CTA-per-image decode + NMS (decode_nms.cu), split so that the fused loss kernel can do the
// counting pass: fused_prepare validates, carves the workspace, clears its control block (one
// memset) and returns the per-image buckets the counting pass must fill; fused_finish launches
// the per-image kernel.
int fused_prepare(const void* const* preds, int64_t n_img, const yb_decode_params* p, int row_cap, void* workspace,
                  size_t workspace_bytes, DecodeLaunch& D, HotBuckets& K, cudaStream_t stream, bool clear = true);
int fused_finish(const DecodeLaunch& D, int64_t n_img, int row_cap, void* workspace, double nms_threshold,
                 int iou_mode, double* out_rows, int64_t out_capacity, int64_t* out_offsets,
                 unsigned int* n_overflow, cudaStream_t stream);

}  // namespace yb
