// conv28.h — host interface of the large-batch kernels of the first LeNet stage (conv28.cu): CLayer 5x5 on one 28x28
// channel (stride 1, no padding, even filter count) + untied bias + activation + PLayer 2x2 max|min, Real = float,
// forward and backward (CLayer.inl:108-197, PLayer.inl:115-170, Tensor.cpp:466-546,673-734).  BASELINE.json configs[1], [2]
// and [4] start with this stage; at mini-batches of a thousand samples and more it was 2/3 of the step.
// Compiled in its own translation unit; dlb200.cu only sees this.
#pragma once

#include <cstddef>
#include <cstdint>

#ifndef DLB_CUDA_EMU
#include <cuda_runtime.h>
#endif

namespace dlb {
namespace conv28 {

#ifndef DLB_CUDA_EMU
struct Stage {
    int F;            // filters (even, <= 64)
    int act;          // DLB_ACT_* of the CLayer (element-wise)
    int is_max;       // PLayer operator: 1 max, 0 min
};
// true when the layer pair has the shape above
bool eligible(int in_c, int in_h, int in_w, int win_h, int win_w, int pad_h, int pad_w, int stride_h, int stride_w, int filters, int activation,
              int pool_h, int pool_w, int pool_type);

// pooled [B][F][144]; sel [B][F][144] bytes: position 0..3 of the window's extremum (row-major inside the 2x2 window), 255 = none
// (all candidates NaN).  x rows are addressed through gather_base[*cursor + b] when gather_base is given (device memory).
void forward(cudaStream_t st, const Stage& s, const float* x, const int* gather_base, const int* cursor, const float* bias, const float* filt,
             float* pooled, unsigned char* sel, int B);

// partial [rows][F * 576 + F * 25]: per sample-group sums of the untied-bias and filter gradients in the layer's flat parameter
// layout; returns the number of rows written (<= max_rows; every element of those rows is written).
int backward(cudaStream_t st, const Stage& s, const float* x, const int* gather_base, const int* cursor, const float* dpool, const unsigned char* sel,
             const float* pooled, float* partial, int max_rows, int B);
#endif

}  // namespace conv28
}  // namespace dlb
