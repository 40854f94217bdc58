// TensorFlow custom ops wrapping the C ABI of libyolo_b200.so (include/yolo_b200.h).
//
//   YoloGridLoss        (y_true, y_pred; attrs) -> (loss [], dpred like y_pred)
//   YoloGridLossMetrics (y_true, y_pred; attrs) -> (loss [], dpred, metrics double[YB_LOSS_METRICS])
//   YoloGridLossFused   (N x y_true, N x y_pred; per-scale attr lists) -> (loss [N], N x dpred)
//
// One kernel class serves the four reference closures (attr `version`): yolov4/losses/loss.py:64-169,
// yolov3/losses/loss.py:40-164, yolov2/losses/loss.py:40-137, yolov1_5/losses/loss.py:40-118; the
// metrics variant adds yolov*/metrics/yolo_metrics.py:9-115 in the same pass; the fused variant
// runs the FPN outputs of one train step (yolov4/__init__.py:518-535) in ONE launch.
// The gradients are registered in yolo_loss_op.py: (None, upstream * dpred).
//
// Build (only where TensorFlow headers exist - not in the B200 build container, where
// tests/tf_stub/ stands in for the headers so that this file is at least compiled):
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC yolo_loss_op.cc -o libyolo_b200_tf.so $TF_CFLAGS $TF_LFLAGS
//       -I../../include -L.. -lyolo_b200 -Wl,-rpath,'$ORIGIN/..' -DGOOGLE_CUDA=1
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"
#include "yolo_b200.h"

namespace tf = tensorflow;

// The keyword arguments of the four wrap_yolo_loss signatures, as op attributes.
#define YB_LOSS_ATTRS                                   \
    .Attr("version: int")                               \
    .Attr("grid_h: int")                                \
    .Attr("grid_w: int")                                \
    .Attr("bbox_num: int")                              \
    .Attr("class_num: int")                             \
    .Attr("anchors: list(float) = []")                  \
    .Attr("binary_weight: float = 1.0")                 \
    .Attr("loss_weight: list(float)")                   \
    .Attr("wh_reg_weight: float = 0.01")                \
    .Attr("ignore_thresh: float = 0.6")                 \
    .Attr("truth_thresh: float = 1.0")                  \
    .Attr("label_smooth: float = 0.0")                  \
    .Attr("focal_loss_gamma: float = 2.0")              \
    .Attr("use_focal_loss: bool = false")               \
    .Attr("use_scale: bool = true")                     \
    .Attr("from_logits: bool = false")                  \
    .Attr("global_batch: int = 0")

REGISTER_OP("YoloGridLoss")
    .Input("y_true: float")
    .Input("y_pred: float")
    .Output("loss: float")
    .Output("dpred: float")
    YB_LOSS_ATTRS
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
        c->set_output(0, c->Scalar());
        c->set_output(1, c->input(1));
        return tf::OkStatus();
    });

REGISTER_OP("YoloGridLossMetrics")
    .Input("y_true: float")
    .Input("y_pred: float")
    .Output("loss: float")
    .Output("dpred: float")
    .Output("metrics: double")
    YB_LOSS_ATTRS
    .Attr("recall_iou_threshold: float = 0.5")
    .Attr("want_grad: bool = true")             // false: metrics only, dpred comes back empty
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
        c->set_output(0, c->Scalar());
        c->set_output(1, c->input(1));
        c->set_output(2, c->Vector(YB_LOSS_METRICS));
        return tf::OkStatus();
    });

REGISTER_OP("YoloGridLossFused")
    .Input("y_true: N * float")
    .Input("y_pred: N * float")
    .Output("loss: float")
    .Output("dpred: N * float")
    .Attr("N: int >= 1")
    .Attr("version: int")
    .Attr("grid_h: list(int)")
    .Attr("grid_w: list(int)")
    .Attr("bbox_num: int")
    .Attr("class_num: int")
    .Attr("anchors: list(float) = []")          // N * bbox_num (w, h) pairs, scale-major; [] = None
    .Attr("binary_weight: list(float)")         // one per scale
    .Attr("loss_weight: list(float)")
    .Attr("wh_reg_weight: float = 0.01")
    .Attr("ignore_thresh: float = 0.6")
    .Attr("truth_thresh: float = 1.0")
    .Attr("label_smooth: float = 0.0")
    .Attr("focal_loss_gamma: float = 2.0")
    .Attr("use_focal_loss: bool = false")
    .Attr("use_scale: bool = true")
    .Attr("from_logits: bool = false")
    .Attr("global_batch: int = 0")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
        int n = 0;
        TF_RETURN_IF_ERROR(c->GetAttr("N", &n));
        c->set_output(0, c->Vector(n));
        for (int i = 0; i < n; ++i) c->set_output(1 + i, c->input(n + i));
        return tf::OkStatus();
    });

namespace {

// Number of loss weights each reference signature takes (v4: box, conf, prob; v1-v3: xy, wh, conf, prob).
inline int expected_loss_weights(int version) { return version == 4 ? 3 : 4; }

// Attributes shared by the three ops -> yb_loss_params (everything but the grid, anchors and
// binary weight, which the fused op holds per scale).
#define YB_ATTR(NAME, PTR)                                   \
    do {                                                     \
        const tf::Status _st = ctx->GetAttr(NAME, PTR);      \
        if (!_st.ok()) {                                     \
            ctx->CtxFailure(_st);                            \
            return false;                                    \
        }                                                    \
    } while (0)
#define YB_ATTR_CHECK(COND, ...)                                        \
    do {                                                                \
        if (!(COND)) {                                                  \
            ctx->CtxFailure(tf::errors::InvalidArgument(__VA_ARGS__));  \
            return false;                                               \
        }                                                               \
    } while (0)

bool read_common_attrs(tf::OpKernelConstruction* ctx, yb_loss_params* p, int* global_batch) {
    int v = 0;
    std::vector<float> lw;
    bool focal = false, scale = true, logits = false;
    *p = yb_loss_params{};
    YB_ATTR("version", &v);
    YB_ATTR_CHECK(v >= 1 && v <= 4, "version must be 1, 2, 3 or 4, got ", v);
    p->version = v;
    YB_ATTR("bbox_num", &p->bbox_num);
    YB_ATTR("class_num", &p->class_num);
    YB_ATTR_CHECK(p->bbox_num >= 1 && p->bbox_num <= YB_MAX_BOXES && p->class_num >= 1,
                  "bbox_num must be in [1, ", YB_MAX_BOXES, "], class_num >= 1");
    YB_ATTR("loss_weight", &lw);
    // the reference indexes loss_weight[0..2] (v4) / [0..3] (v1-v3): a shorter list raises there
    YB_ATTR_CHECK((int)lw.size() == expected_loss_weights(v), "loss_weight needs ", expected_loss_weights(v),
                  " entries for version ", v, ", got ", lw.size());
    for (size_t i = 0; i < 4; ++i) p->loss_weight[i] = i < lw.size() ? lw[i] : 0.f;
    YB_ATTR("wh_reg_weight", &p->wh_reg_weight);
    YB_ATTR("ignore_thresh", &p->ignore_thresh);
    YB_ATTR("truth_thresh", &p->truth_thresh);
    YB_ATTR("label_smooth", &p->label_smooth);
    YB_ATTR("focal_loss_gamma", &p->focal_gamma);
    YB_ATTR("use_focal_loss", &focal);
    YB_ATTR("use_scale", &scale);
    YB_ATTR("from_logits", &logits);
    YB_ATTR_CHECK(!logits || v == 3 || v == 4, "from_logits exists for the v3/v4 heads only");
    p->use_focal = focal;
    p->use_scale = scale;
    p->from_logits = logits;
    YB_ATTR("global_batch", global_batch);
    return true;
}
#undef YB_ATTR
#undef YB_ATTR_CHECK

inline int64_t values_per_cell(const yb_loss_params& p) {
    return p.version == 1 ? 5 * (int64_t)p.bbox_num + p.class_num : (int64_t)p.bbox_num * (5 + p.class_num);
}

inline yb_stream_t cuda_stream_of(tf::OpKernelContext* ctx) {
    auto* stream = ctx->op_device_context()->stream();
    return reinterpret_cast<yb_stream_t>(stream->platform_specific_handle().stream);
}

// 256-byte aligned scratch of `bytes` bytes inside a temp tensor.
inline char* aligned_workspace(tf::OpKernelContext* ctx, size_t bytes, tf::Tensor* holder) {
    if (!ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(int64_t)bytes + 256}), holder).ok()) return nullptr;
    char* p = reinterpret_cast<char*>(holder->flat<tf::uint8>().data());
    return p + (256 - reinterpret_cast<uintptr_t>(p) % 256) % 256;
}

}  // namespace

// YoloGridLoss and YoloGridLossMetrics (kMetrics): one scale per op instance, as Keras calls the
// closures of model.compile(loss=[f0, f1, f2]) one output at a time.
template <bool kMetrics>
class YoloGridLossOp : public tf::OpKernel {
 public:
    explicit YoloGridLossOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
        std::vector<float> anchors;
        if (!read_common_attrs(ctx, &p_, &global_batch_)) return;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_h", &p_.grid_h));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_w", &p_.grid_w));
        OP_REQUIRES(ctx, p_.grid_h > 0 && p_.grid_w > 0, tf::errors::InvalidArgument("grid_shape must be positive"));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("anchors", &anchors));
        OP_REQUIRES(ctx, anchors.empty() || (int)anchors.size() == 2 * p_.bbox_num,
                    tf::errors::InvalidArgument("anchors must hold bbox_num (w,h) pairs"));
        OP_REQUIRES(ctx, !(p_.version == 2 && anchors.empty()),
                    tf::errors::InvalidArgument("yolov2's wrap_yolo_loss has no default for anchors"));
        OP_REQUIRES(ctx, !(p_.from_logits && anchors.empty()),
                    tf::errors::InvalidArgument("from_logits needs the anchors of the scale"));
        p_.has_anchors = (anchors.empty() || p_.version == 1) ? 0 : 1;
        for (size_t i = 0; i < anchors.size(); ++i) p_.anchors[i] = anchors[i];
        OP_REQUIRES_OK(ctx, ctx->GetAttr("binary_weight", &p_.binary_weight));
        if (kMetrics) {
            OP_REQUIRES_OK(ctx, ctx->GetAttr("recall_iou_threshold", &recall_thr_));
            OP_REQUIRES_OK(ctx, ctx->GetAttr("want_grad", &want_grad_));
        }
    }

    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& y_true = ctx->input(0);
        const tf::Tensor& y_pred = ctx->input(1);
        const int64_t cells_per_img = (int64_t)p_.grid_h * p_.grid_w;
        const int64_t pcf = values_per_cell(p_);
        OP_REQUIRES(ctx, y_pred.NumElements() % (cells_per_img * pcf) == 0,
                    tf::errors::InvalidArgument("y_pred does not reshape to (-1, grid_h, grid_w, info)"));
        const int64_t n_img = y_pred.NumElements() / (cells_per_img * pcf);
        OP_REQUIRES(ctx, y_true.NumElements() == n_img * cells_per_img * (5 + p_.class_num),
                    tf::errors::InvalidArgument("y_true does not match y_pred"));
        tf::Tensor* loss = nullptr;
        tf::Tensor* dpred = nullptr;
        tf::Tensor* metrics = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({}), &loss));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, want_grad_ ? y_pred.shape() : tf::TensorShape({0}), &dpred));
        if (kMetrics) OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({YB_LOSS_METRICS}), &metrics));
        tf::Tensor ws;
        const size_t ws_bytes = yb_loss_workspace_bytes(1);
        char* ws_ptr = aligned_workspace(ctx, ws_bytes, &ws);
        OP_REQUIRES(ctx, ws_ptr != nullptr, tf::errors::ResourceExhausted("yolo_b200: no workspace"));

        yb_loss_scale sc;
        sc.y_true = y_true.flat<float>().data();
        sc.y_pred = y_pred.flat<float>().data();
        sc.dpred = want_grad_ ? dpred->flat<float>().data() : nullptr;   // NULL = forward only
        sc.n_cells = n_img * cells_per_img;
        sc.p = p_;
        sc.p.inv_batch = 1.0 / (double)(global_batch_ > 0 ? global_batch_ : std::max<int64_t>(n_img, 1));
        int rc;
        if (kMetrics)
            rc = yb_loss_fwd_bwd_metrics(&sc, 1, loss->flat<float>().data(), nullptr, metrics->flat<double>().data(),
                                         (double)recall_thr_, ws_ptr, ws_bytes, cuda_stream_of(ctx));
        else
            rc = yb_loss_fwd_bwd(&sc, 1, loss->flat<float>().data(), nullptr, ws_ptr, ws_bytes, cuda_stream_of(ctx));
        OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("yolo_b200: ", yb_status_string(rc)));
    }

 private:
    yb_loss_params p_;
    int global_batch_ = 0;
    float recall_thr_ = 0.5f;
    bool want_grad_ = true;
};

// YoloGridLossFused: the N FPN outputs of one train step in ONE launch (yb_loss_fwd_bwd with
// n_scales = N), e.g. from a custom train_step; loss[i] is what closure i of Yolo.loss() returns.
class YoloGridLossFusedOp : public tf::OpKernel {
 public:
    explicit YoloGridLossFusedOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
        std::vector<float> anchors, bw;
        std::vector<int> gh, gw;
        yb_loss_params common;
        if (!read_common_attrs(ctx, &common, &global_batch_)) return;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("N", &n_));
        OP_REQUIRES(ctx, n_ >= 1 && n_ <= YB_MAX_SCALES,
                    tf::errors::InvalidArgument("between 1 and ", YB_MAX_SCALES, " scales"));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_h", &gh));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_w", &gw));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("anchors", &anchors));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("binary_weight", &bw));
        OP_REQUIRES(ctx, (int)gh.size() == n_ && (int)gw.size() == n_ && (int)bw.size() == n_,
                    tf::errors::InvalidArgument("grid_h, grid_w and binary_weight need one entry per scale"));
        OP_REQUIRES(ctx, anchors.empty() || (int)anchors.size() == 2 * common.bbox_num * n_,
                    tf::errors::InvalidArgument("anchors must hold N * bbox_num (w,h) pairs"));
        OP_REQUIRES(ctx, !((common.version == 2 || common.from_logits) && anchors.empty()),
                    tf::errors::InvalidArgument("anchors are required (yolov2 / from_logits)"));
        for (int s = 0; s < n_; ++s) {
            p_[s] = common;
            p_[s].grid_h = gh[s];
            p_[s].grid_w = gw[s];
            OP_REQUIRES(ctx, gh[s] > 0 && gw[s] > 0, tf::errors::InvalidArgument("grid_shape must be positive"));
            p_[s].binary_weight = bw[s];
            p_[s].has_anchors = (anchors.empty() || common.version == 1) ? 0 : 1;
            for (int i = 0; i < 2 * common.bbox_num && !anchors.empty(); ++i)
                p_[s].anchors[i] = anchors[(size_t)s * 2 * common.bbox_num + i];
        }
    }

    void Compute(tf::OpKernelContext* ctx) override {
        yb_loss_scale sc[YB_MAX_SCALES];
        int64_t n_img = -1;
        for (int s = 0; s < n_; ++s) {
            const tf::Tensor& y_true = ctx->input(s);
            const tf::Tensor& y_pred = ctx->input(n_ + s);
            const int64_t cells_per_img = (int64_t)p_[s].grid_h * p_[s].grid_w;
            const int64_t pcf = values_per_cell(p_[s]);
            OP_REQUIRES(ctx, y_pred.NumElements() % (cells_per_img * pcf) == 0,
                        tf::errors::InvalidArgument("y_pred[", s, "] does not reshape to (-1, grid_h, grid_w, info)"));
            const int64_t n = y_pred.NumElements() / (cells_per_img * pcf);
            OP_REQUIRES(ctx, n_img < 0 || n == n_img, tf::errors::InvalidArgument("scales hold different batches"));
            n_img = n;
            OP_REQUIRES(ctx, y_true.NumElements() == n * cells_per_img * (5 + p_[s].class_num),
                        tf::errors::InvalidArgument("y_true[", s, "] does not match y_pred[", s, "]"));
            tf::Tensor* dpred = nullptr;
            OP_REQUIRES_OK(ctx, ctx->allocate_output(1 + s, y_pred.shape(), &dpred));
            sc[s].y_true = y_true.flat<float>().data();
            sc[s].y_pred = y_pred.flat<float>().data();
            sc[s].dpred = dpred->flat<float>().data();
            sc[s].n_cells = n * cells_per_img;
            sc[s].p = p_[s];
        }
        for (int s = 0; s < n_; ++s)
            sc[s].p.inv_batch = 1.0 / (double)(global_batch_ > 0 ? global_batch_ : std::max<int64_t>(n_img, 1));
        tf::Tensor* loss = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n_}), &loss));
        tf::Tensor ws;
        const size_t ws_bytes = yb_loss_workspace_bytes(n_);
        char* ws_ptr = aligned_workspace(ctx, ws_bytes, &ws);
        OP_REQUIRES(ctx, ws_ptr != nullptr, tf::errors::ResourceExhausted("yolo_b200: no workspace"));
        const int rc = yb_loss_fwd_bwd(sc, n_, loss->flat<float>().data(), nullptr, ws_ptr, ws_bytes, cuda_stream_of(ctx));
        OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("yolo_b200: ", yb_status_string(rc)));
    }

 private:
    yb_loss_params p_[YB_MAX_SCALES];
    int n_ = 0;
    int global_batch_ = 0;
};

// GPU only: there is deliberately no CPU kernel.
REGISTER_KERNEL_BUILDER(Name("YoloGridLoss").Device(tf::DEVICE_GPU), YoloGridLossOp<false>);
REGISTER_KERNEL_BUILDER(Name("YoloGridLossMetrics").Device(tf::DEVICE_GPU), YoloGridLossOp<true>);
REGISTER_KERNEL_BUILDER(Name("YoloGridLossFused").Device(tf::DEVICE_GPU), YoloGridLossFusedOp);
