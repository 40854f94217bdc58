// Pair-level building blocks of the NMS kernels (nms.cu, decode_nms.cu): the reference's visit
// order (utils/tools.py:717), its IoU / DIoU expression (utils/tools.py:649-682) and the
// division-free suppression test with the pinned exact fallback.  Callers are compiled with
// -fmad=false (NumPy rounds every multiply / add separately).
#pragma once
#include "common.cuh"

namespace yb {

constexpr double kIouEps = 1e-07;  // utils/tools.py:26

// Visit order of np.argsort(conf)[::-1] (tools.py:717): descending confidence, NaN first (NumPy
// sorts NaN last), equal confidences -> higher original index first (documented tie rule).
__device__ __forceinline__ bool visited_before(double ca, int a, double cb, int b) {
    const bool na = ca != ca, nb = cb != cb;
    if (na != nb) return na;
    if (!na && ca != cb) return ca > cb;
    return a > b;
}

// np.maximum / np.minimum: NaN propagates (fmax/fmin would drop it)
__device__ __forceinline__ double np_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double np_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
// the same for operands known not to be NaN: one compare + select
__device__ __forceinline__ double sel_max(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double sel_min(double a, double b) { return a < b ? a : b; }

// IoU (MODE 1) or DIoU (MODE 2) of "true" box t against "pred" box p, the
// reference's operation order (tools.py:649-682), NaN/Inf behaving as in NumPy.
template <int MODE>
__device__ __forceinline__ double pair_iou(double tx, double ty, double tw, double th, double px,
                                           double py, double pw, double ph) {
    const double thw = tw / 2.0, thh = th / 2.0, phw = pw / 2.0, phh = ph / 2.0;
    const double t0x = tx - thw, t0y = ty - thh, t1x = tx + thw, t1y = ty + thh;
    const double p0x = px - phw, p0y = py - phh, p1x = px + phw, p1y = py + phh;
    const double iw = np_max(np_min(p1x, t1x) - np_max(p0x, t0x), 0.0);
    const double ih = np_max(np_min(p1y, t1y) - np_max(p0y, t0y), 0.0);
    const double inter = iw * ih;
    const double uni = pw * ph + tw * th - inter;
    const double iou = inter / (uni + kIouEps);
    if (MODE == 1) return iou;
    const double ew = np_max(p1x, t1x) - np_min(p0x, t0x), eh = np_max(p1y, t1y) - np_min(p0y, t0y);
    const double c2 = ew * ew + eh * eh;
    const double dx = tx - px, dy = ty - py;
    const double rho2 = dx * dx + dy * dy;
    return iou - rho2 / c2;
}

// A box as the sweeps use it: corners and area computed ONCE per box with the reference's
// operations (x -+ w/2, y -+ h/2, w*h: tools.py:649-653,660), so every pair starts from the same
// bits the broadcast expression produces.  A box with a NaN among these values can neither
// suppress nor be suppressed (NumPy propagates the NaN into the IoU, which compares False): it is
// staged as an inverted box at infinity, whose overlap with anything is negative.
struct BoxC {
    double x0, x1, y0, y1, area, cx, cy;
};

__device__ __forceinline__ BoxC make_box(double x, double y, double w, double h) {
    BoxC b;
    const double hw = w / 2.0, hh = h / 2.0;
    b.x0 = x - hw; b.x1 = x + hw; b.y0 = y - hh; b.y1 = y + hh;
    b.area = w * h;
    b.cx = x; b.cy = y;
    const double probe = ((b.x0 + b.x1) + (b.y0 + b.y1)) + b.area;   // NaN iff any of them is (or inf - inf)
    if (probe != probe && (b.x0 != b.x0 || b.x1 != b.x1 || b.y0 != b.y0 || b.y1 != b.y1 || b.area != b.area)) {
        b.x0 = INFINITY; b.x1 = -INFINITY; b.y0 = INFINITY; b.y1 = -INFINITY;
    }
    return b;
}

// fp64 division the compiler may not move: the sweeps reach the exact expression for a vanishing
// fraction of the pairs, but a plain `/` lets the compiler hoist the division (and its slow path
// for a zero numerator) in front of the margin tests of EVERY pair.
__device__ __forceinline__ double div_pinned(double a, double b) {
    double q;
    asm volatile("div.rn.f64 %0, %1, %2;" : "=d"(q) : "d"(a), "d"(b));
    return q;
}

// Exact decision from the original rows: the reference's expression with IEEE divisions.
template <int MODE>
__device__ __forceinline__ bool suppresses_exact(const double* __restrict__ rows, long long ra, long long rb,
                                                 double thr) {
    const double* a = rows + ra * 7;
    const double* b = rows + rb * 7;
    const double tx = a[0], ty = a[1], tw = a[2], th = a[3], px = b[0], py = b[1], pw = b[2], ph = b[3];
    const double thw = tw / 2.0, thh = th / 2.0, phw = pw / 2.0, phh = ph / 2.0;
    const double t0x = tx - thw, t0y = ty - thh, t1x = tx + thw, t1y = ty + thh;
    const double p0x = px - phw, p0y = py - phh, p1x = px + phw, p1y = py + phh;
    const double iw = np_max(np_min(p1x, t1x) - np_max(p0x, t0x), 0.0);
    const double ih = np_max(np_min(p1y, t1y) - np_max(p0y, t0y), 0.0);
    const double inter = iw * ih;
    const double uni = pw * ph + tw * th - inter;
    const double iou = div_pinned(inter, uni + kIouEps);
    if (MODE == 1) return iou >= thr;
    const double ew = np_max(p1x, t1x) - np_min(p0x, t0x), eh = np_max(p1y, t1y) - np_min(p0y, t0y);
    const double c2 = ew * ew + eh * eh;
    const double dx = tx - px, dy = ty - py;
    const double rho2 = dx * dx + dy * dy;
    return (iou - div_pinned(rho2, c2)) >= thr;
}

// Does box a suppress box b, i.e. is fl(IoU) (MODE 1) / fl(fl(IoU) - fl(rho2/c2)) (MODE 2) >= thr ?
// Returns 1 / 0 when the answer is certain WITHOUT a division, -1 when the exact expression must be
// evaluated:
//   * thr > 0 and the boxes do not overlap: inter = 0, so IoU = 0 (or NaN) and DIoU <= 0 (or NaN):
//     never >= thr.  This also covers NaN boxes (staged with negative overlap).
//   * otherwise the cross-multiplied inequality is tested with a margin wider than every rounding
//     error involved (x = inter/den real: x >= thr => fl(x) >= thr; x < thr(1-2^-51) => fl(x) < thr).
// Second stage of the division-free test, for boxes that overlap (iw, ih > 0): 1 / 0 when
// certain, -1 when the exact expression must be evaluated.  (ew, eh) = extent of the enclosing
// box, (dx, dy) = centre distance (DIoU only).
template <int MODE>
__device__ __forceinline__ int decide_overlapping(double iw, double ih, double area_a, double area_b, double ew,
                                                  double eh, double dx, double dy, double thr) {
    const double inter = iw * ih;
    const double den = ((area_b + area_a) - inter) + kIouEps;
    if (!(den > 0.0)) return -1;
    if (MODE == 1) {
        const double P = thr * den;
        if (inter >= P * (1.0 + 4.5e-16)) return 1;
        if (inter <= P * (1.0 - 9.0e-16)) return 0;
        return -1;
    }
    const double c2 = ew * ew + eh * eh;
    const double rho2 = dx * dx + dy * dy;
    // y = inter/den - rho2/c2 (real); the computed value differs from y by < 4e-16 for
    // |terms| <= 1, so |y - thr| > 1e-15 decides.  y - thr = (A - B - thr*S) / S with
    // A = inter*c2, B = rho2*den, S = den*c2 > 0; every product carries 2^-53 relative error.
    const double A = inter * c2, B = rho2 * den, S = den * c2;
    const double lhs = A - B, rhs = thr * S;
    const double slack = 4.5e-16 * (A + B + fabs(rhs)) + 2.0e-15 * S;
    const bool sane = inter <= den && rho2 <= c2;
    if (lhs - rhs > slack && sane) return 1;
    if (rhs - lhs > slack && sane) return 0;
    return -1;
}

template <int MODE>
__device__ __forceinline__ int suppresses_fast(const BoxC& a, const BoxC& b, double thr, bool pos_thr) {
    const double iw = sel_min(a.x1, b.x1) - sel_max(a.x0, b.x0);
    const double ih = sel_min(a.y1, b.y1) - sel_max(a.y0, b.y0);
    if (!pos_thr) return -1;
    if (!(iw > 0.0 && ih > 0.0)) return 0;
    double ew = 0.0, eh = 0.0;
    if (MODE == 2) {
        ew = sel_max(a.x1, b.x1) - sel_min(a.x0, b.x0);
        eh = sel_max(a.y1, b.y1) - sel_min(a.y0, b.y0);
    }
    return decide_overlapping<MODE>(iw, ih, a.area, b.area, ew, eh, a.cx - b.cx, a.cy - b.cy, thr);
}

template <int MODE>
__device__ __forceinline__ bool suppresses(const BoxC& a, const BoxC& b, double thr, bool pos_thr,
                                           const double* __restrict__ rows, long long ra, long long rb) {
    const int r = suppresses_fast<MODE>(a, b, thr, pos_thr);
    if (r >= 0) return r != 0;
    return suppresses_exact<MODE>(rows, ra, rb, thr);
}

}  // namespace yb
