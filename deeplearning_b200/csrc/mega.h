// mega.h — host interface of the LeNet step megakernel (mega.cu): ONE persistent cooperative kernel per training step for
// nets of the shape  CLayer 5x5 (1 channel, 28x28, stride 1, no padding, even filter count) -> PLayer 2x2 max|min ->
// NLayer (<= 128 outputs, element-wise activation) -> NLayer head (<= 16 outputs)  with Real = float - BASELINE.json
// configs[1] and [4].  Everything Net<D>::learn does for one mini-batch (Net.inl:226-317) runs inside it; phases are
// separated by grid barriers instead of kernel launches.  Compiled in its own translation unit; dlb200.cu only sees this.
#pragma once

#include <cstddef>
#include <cstdint>

#ifndef DLB_CUDA_EMU
#include <cuda_runtime.h>
#endif

namespace dlb {
namespace mega {

struct Shape { int F, H1, n_out; };

#ifndef DLB_CUDA_EMU
// What the net (dlb200.cu) hands over per step; everything the kernel needs beyond it (scratch, barrier, decomposition) is
// owned by the Runner.  Kept stable so that work on the kernel does not recompile the rest of the library.
struct StepIn {
    cudaStream_t stream;
    const float* x; const float* refs; const int* gather_base; const int* cursor; int* cursor_rw; int advance;
    int B, cap;                        // samples of this rank; batch capacity of pooled / dpool
    int act0, is_max, act1, act2, cost_id;
    float* params; float* grads; long long off_b0, off_k0, off_b1, off_w1, off_b2, off_w2, n_params;
    float* pooled; float* dpool;       // the net's PLayer output and PLayer output-gradient buffers
    double lr, reg_factor; long long batch_items; double* cost_out;
    int n_steps;                       // consecutive steps of identical shape to run (the cursor advances on the device)
    int world, rank; const void* const* peer_regions; size_t publish_bytes; unsigned long long* step_counter;   // data parallelism over peer memory
    int split_phases;                  // 1: one launch per phase (per-phase timing), 0: one cooperative launch
    long long reserved[8];
};
class Runner;
Runner* runner_create(const Shape& shape, int device);
void runner_destroy(Runner* r);
int runner_step(Runner* r, const StepIn& in);   // returns the number of kernel launches
// development: copies one of the runner's scratch buffers (ids 101..) to the host; returns the bytes copied
size_t runner_debug_read(Runner* r, int what, void* dst, size_t bytes, cudaStream_t stream);
#endif

}  // namespace mega
}  // namespace dlb
