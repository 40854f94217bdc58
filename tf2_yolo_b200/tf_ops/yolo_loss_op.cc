// TensorFlow custom op wrapping the C ABI of libyolo_b200.so (include/yolo_b200.h).
//
//   YoloGridLoss(y_true: float, y_pred: float; attrs ...) -> (loss: float [], dpred: float like y_pred)
//
// One op serves the four reference closures (attr `version`): yolov4/losses/loss.py:64-169,
// yolov3/losses/loss.py:40-164, yolov2/losses/loss.py:40-137, yolov1_5/losses/loss.py:40-118.
// The gradient is registered in yolo_loss_op.py: (None, upstream * dpred).
//
// Build (only where TensorFlow headers exist - not in the B200 build container):
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC yolo_loss_op.cc -o libyolo_b200_tf.so $TF_CFLAGS $TF_LFLAGS \
//       -I../../include -L.. -lyolo_b200 -Wl,-rpath,'$ORIGIN/..' -DGOOGLE_CUDA=1
#include <vector>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"
#include "yolo_b200.h"

namespace tf = tensorflow;

REGISTER_OP("YoloGridLoss")
    .Input("y_true: float")
    .Input("y_pred: float")
    .Output("loss: float")
    .Output("dpred: float")
    .Attr("version: int")
    .Attr("grid_h: int")
    .Attr("grid_w: int")
    .Attr("bbox_num: int")
    .Attr("class_num: int")
    .Attr("anchors: list(float) = []")
    .Attr("binary_weight: float = 1.0")
    .Attr("loss_weight: list(float)")
    .Attr("wh_reg_weight: float = 0.01")
    .Attr("ignore_thresh: float = 0.6")
    .Attr("truth_thresh: float = 1.0")
    .Attr("label_smooth: float = 0.0")
    .Attr("focal_loss_gamma: float = 2.0")
    .Attr("use_focal_loss: bool = false")
    .Attr("use_scale: bool = true")
    .Attr("global_batch: int = 0")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
        c->set_output(0, c->Scalar());
        c->set_output(1, c->input(1));
        return tf::OkStatus();
    });

class YoloGridLossOp : public tf::OpKernel {
 public:
    explicit YoloGridLossOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
        int v;
        std::vector<float> anchors, lw;
        bool focal, scale;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("version", &v));
        p_ = yb_loss_params{};
        p_.version = v;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_h", &p_.grid_h));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("grid_w", &p_.grid_w));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("bbox_num", &p_.bbox_num));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("class_num", &p_.class_num));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("anchors", &anchors));
        OP_REQUIRES(ctx, anchors.empty() || (int)anchors.size() == 2 * p_.bbox_num,
                    tf::errors::InvalidArgument("anchors must hold bbox_num (w,h) pairs"));
        OP_REQUIRES(ctx, p_.bbox_num <= YB_MAX_BOXES, tf::errors::InvalidArgument("bbox_num too large"));
        p_.has_anchors = anchors.empty() ? 0 : 1;
        for (size_t i = 0; i < anchors.size(); ++i) p_.anchors[i] = anchors[i];
        OP_REQUIRES_OK(ctx, ctx->GetAttr("binary_weight", &p_.binary_weight));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("loss_weight", &lw));
        for (size_t i = 0; i < 4; ++i) p_.loss_weight[i] = i < lw.size() ? lw[i] : 0.f;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("wh_reg_weight", &p_.wh_reg_weight));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("ignore_thresh", &p_.ignore_thresh));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("truth_thresh", &p_.truth_thresh));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("label_smooth", &p_.label_smooth));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("focal_loss_gamma", &p_.focal_gamma));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("use_focal_loss", &focal));
        OP_REQUIRES_OK(ctx, ctx->GetAttr("use_scale", &scale));
        p_.use_focal = focal;
        p_.use_scale = scale;
        OP_REQUIRES_OK(ctx, ctx->GetAttr("global_batch", &global_batch_));
    }

    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& y_true = ctx->input(0);
        const tf::Tensor& y_pred = ctx->input(1);
        const int64_t cells_per_img = (int64_t)p_.grid_h * p_.grid_w;
        const int64_t pcf = p_.version == 1 ? 5 * p_.bbox_num + p_.class_num : p_.bbox_num * (5 + p_.class_num);
        OP_REQUIRES(ctx, y_pred.NumElements() % (cells_per_img * pcf) == 0,
                    tf::errors::InvalidArgument("y_pred does not reshape to (-1, grid_h, grid_w, info)"));
        const int64_t n_img = y_pred.NumElements() / (cells_per_img * pcf);
        OP_REQUIRES(ctx, y_true.NumElements() == n_img * cells_per_img * (5 + p_.class_num),
                    tf::errors::InvalidArgument("y_true does not match y_pred"));
        tf::Tensor* loss = nullptr;
        tf::Tensor* dpred = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({}), &loss));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, y_pred.shape(), &dpred));
        tf::Tensor ws;
        const size_t ws_bytes = yb_loss_workspace_bytes(1);
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(int64_t)ws_bytes + 256}), &ws));
        char* ws_ptr = reinterpret_cast<char*>(ws.flat<tf::uint8>().data());
        ws_ptr += (256 - reinterpret_cast<uintptr_t>(ws_ptr) % 256) % 256;

        yb_loss_params p = p_;
        p.inv_batch = 1.0 / (double)(global_batch_ > 0 ? global_batch_ : std::max<int64_t>(n_img, 1));
        auto* stream = ctx->op_device_context()->stream();
        yb_stream_t cu_stream = reinterpret_cast<yb_stream_t>(
            stream->platform_specific_handle().stream);
        int rc;
        const float* yt = y_true.flat<float>().data();
        const float* yp = y_pred.flat<float>().data();
        float* lo = loss->flat<float>().data();
        float* dp = dpred->flat<float>().data();
        const int64_t n_cells = n_img * cells_per_img;
        switch (p.version) {
            case 1: rc = yb_loss_v1_fwd_bwd(yt, yp, n_cells, lo, dp, &p, ws_ptr, ws_bytes, cu_stream); break;
            case 2: rc = yb_loss_v2_fwd_bwd(yt, yp, n_cells, lo, dp, &p, ws_ptr, ws_bytes, cu_stream); break;
            case 3: rc = yb_loss_v3_fwd_bwd(yt, yp, n_cells, lo, dp, &p, ws_ptr, ws_bytes, cu_stream); break;
            default: rc = yb_loss_v4_fwd_bwd(yt, yp, n_cells, lo, dp, &p, ws_ptr, ws_bytes, cu_stream); break;
        }
        OP_REQUIRES(ctx, rc == 0, tf::errors::Internal("yolo_b200: ", yb_status_string(rc)));
    }

 private:
    yb_loss_params p_;
    int global_batch_ = 0;
};

// GPU only: there is deliberately no CPU kernel.
REGISTER_KERNEL_BUILDER(Name("YoloGridLoss").Device(tf::DEVICE_GPU), YoloGridLossOp);
