/*
 * dlb200_spec.h — plain-C descriptors shared by the C-ABI (include/dlb200.h), the CPU oracle
 * (oracle/dl_oracle.c) and the reference driver (oracle/ref_driver.cpp).
 *
 * Every enumerator keeps the numeric value the reference uses, so that ids can be passed through
 * unchanged from reference client code:
 *   layer types        DeepLearning/NeuralNet/LayerTypeId.h:27-32
 *   activation ids     DeepLearning/Math/ActivationFunctionId.h:26-33
 *   cost ids           DeepLearning/Math/CostFunctionId.h:26-32
 *   pool ids           DeepLearning/Math/PoolOperator.h:28-33
 */
#ifndef DLB200_SPEC_H
#define DLB200_SPEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DLB_LAYER_UNKNOWN = 0, DLB_LAYER_FULL = 1, DLB_LAYER_CONVOLUTION = 2, DLB_LAYER_PULL = 3 };
enum { DLB_ACT_UNKNOWN = 0, DLB_ACT_SIGMOID = 1, DLB_ACT_TANH = 2, DLB_ACT_RELU = 3, DLB_ACT_SOFTMAX = 4, DLB_ACT_LINEAR = 5 };
enum { DLB_COST_UNKNOWN = 0, DLB_COST_SQUARED_ERROR = 1, DLB_COST_CROSS_ENTROPY = 2, DLB_COST_LINEAR = 3 };
enum { DLB_POOL_UNKNOWN = 0, DLB_POOL_MAX = 1, DLB_POOL_MIN = 2, DLB_POOL_AVERAGE = 3 };

/*
 * One layer of a feed-forward net (NLayer / CLayer / PLayer). Tensors are the reference's
 * Index3d {x = layer(channel), y = row, z = column}, row-major CHW (Tensor.cpp:216-219).
 *
 *  FULL         (NLayer.inl:27-48):  in = in_c*in_h*in_w (flattened), out_c = output dimension.
 *  CONVOLUTION  (CLayer.inl:29-48):  filter window {in_c, win_h, win_w}, out_c filters, zero paddings
 *               {0, pad_h, pad_w}, strides {1, stride_h, stride_w}; the bias tensor is UNTIED and has the
 *               shape of the output {out_c, out_h, out_w} (CLayer.inl:43-44).
 *  PULL         (PLayer.inl:27-33):  window {1, win_h, win_w}, stride == window, no padding; pool_type.
 */
typedef struct dlb_layer_spec {
    int32_t type;        /* DLB_LAYER_* */
    int32_t activation;  /* DLB_ACT_*   (ignored for PULL) */
    int32_t in_c, in_h, in_w;
    int32_t win_h, win_w;
    int32_t out_c;
    int32_t pad_h, pad_w;
    int32_t stride_h, stride_w;
    int32_t pool_type;   /* DLB_POOL_*  (PULL only) */
    int32_t reserved[3];
} dlb_layer_spec;

#ifdef __cplusplus
}
#endif
#endif /* DLB200_SPEC_H */
