/*
 * dlb200.h — C ABI of libdlb200.so: the B200 (sm_100a) implementation of the mini-batch training hot path of
 * dragunovdenis/DeepLearning (Net<D>::learn through CLayer / PLayer / NLayer with the GpuDC data context).
 *
 * This header is the drop-in boundary.  The reference has no FFI of its own — its GPU plug-in seam is the
 * compile-time data-context policy `GpuDC` (DeepLearningCuda/NeuralNet/DataContextCuda.h:28-77) whose tensor
 * types (CudaTensor / CudaMatrix / CudaVector, DeepLearningCuda/Math/*.cuh) call CUDA kernels directly.  Here the
 * same operations are reached through plain `extern "C"` entry points (pointers + sizes + POD structs, `int`
 * status, no C++ or torch types), and the host side that mirrors the reference's classes lives above this line:
 * C++ (deeplearning_b200/host/) for reference client code, Python/ctypes (deeplearning_b200/) for tests and bench.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success; on failure a non-zero code, and dlb_last_error() (thread-local) holds
 *     the message — shape/argument violations and CUDA failures alike (the reference throws std::exception with
 *     the message, CudaUtils.cpp:24-30; the C++ wrappers re-throw);
 *   - `_f32` / `_f64` suffixes select the reference's `Real` (defs.h:25-29: double by default, float under
 *     USE_SINGLE_PRECISION);
 *   - device pointers are raw CUDA device pointers; all work is enqueued on the context's stream and is
 *     asynchronous unless stated otherwise;
 *   - a mini-batch is packed sample-major: [B][C][H][W], each sample in the reference's CHW order
 *     (Tensor.cpp:216-219); matrices are row-major [out][in] (Matrix.cpp:125-128, NLayer.h:41-44);
 *   - flat parameter / gradient layout = the reference's container order (NLayer.inl:180-193,
 *     CLayer.inl:199-213): per layer [biases, weights] or [biases (F x OH x OW, untied), filter_0 .. filter_{F-1}].
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device and fails loudly without one.
 */
#ifndef DLB200_H
#define DLB200_H

#include <stddef.h>
#include <stdint.h>
#include "dlb200_spec.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DLB_API __attribute__((visibility("default")))

typedef struct dlb_context dlb_context; /* one device + one stream + scratch; replaces CudaSetup/cudaStreamPerThread (CudaSetup.cpp:29-76, CudaUtils.h:43-108) */
typedef struct dlb_net dlb_net;         /* batched Net<GpuDC>: parameters, traces, gradient buffers, resident data set */

DLB_API int dlb_version(void);
DLB_API const char* dlb_last_error(void);

/* ---- context & memory (replaces CudaUtils::cuda_allocate/cuda_free/cuda_copy_*, CudaUtils.h:43-108) ---------- */
/* external_stream: a cudaStream_t to enqueue on (e.g. torch's current stream) or NULL to create a private one. */
DLB_API int dlb_context_create(int device, void* external_stream, dlb_context** out);
DLB_API int dlb_context_destroy(dlb_context* ctx);
DLB_API int dlb_context_synchronize(dlb_context* ctx);
DLB_API void* dlb_context_stream(dlb_context* ctx);
DLB_API int dlb_context_device(dlb_context* ctx);
DLB_API int dlb_malloc(dlb_context* ctx, size_t bytes, void** out);
DLB_API int dlb_free(dlb_context* ctx, void* ptr);
DLB_API int dlb_memcpy_h2d(dlb_context* ctx, void* dst, const void* src, size_t bytes);
DLB_API int dlb_memcpy_d2h(dlb_context* ctx, void* dst, const void* src, size_t bytes); /* synchronises the stream */
DLB_API int dlb_memcpy_d2d(dlb_context* ctx, void* dst, const void* src, size_t bytes);
DLB_API int dlb_memset(dlb_context* ctx, void* dst, int value, size_t bytes);
DLB_API int dlb_host_alloc_pinned(size_t bytes, void** out);
DLB_API int dlb_host_free_pinned(void* ptr);
/* number of kernels this library has launched on the context since creation (bench.py's gpu_launches) */
DLB_API int64_t dlb_context_launch_count(dlb_context* ctx);

/* ---- batched layer operators on raw device pointers ---------------------------------------------------------
 * `gather` (nullable, device int32[B]) maps batch slot -> row of `x`, so layer 0 can read a resident data set
 * through a slice of the epoch permutation (Net.inl:260-262) without a copy.
 * `scale` is the reference's gradient_scale_factor: 0 overwrites the gradient, otherwise grad = scale*grad + new
 * (NLayer.inl:158-164, CLayer.inl:148-156, Tensor.cpp:476-479).  `g` is scratch of B*out elements that receives
 * the activation gradient dC/dz.  `dx` may be NULL (first layer, Net.inl:272).  `deriv`/`index_map` may be NULL in
 * forward calls (inference, ALayer::act with trace_data == nullptr).
 */
#define DLB_DECLARE_OPS(SFX, REAL)                                                                                              \
    /* CLayer<D>::act, CLayer.inl:108-137 (replaces CudaTensor::convolve, CudaTensor.cu:395-416 + `+= biases` + activation) */ \
    DLB_API int dlb_conv_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x,                  \
                                       const int32_t* gather, const REAL* bias, const REAL* filters, REAL* y, REAL* deriv);     \
    /* CLayer<D>::backpropagate, CLayer.inl:139-197 (replaces CudaTensor::convolution_gradient, CudaTensor.cu:492-518) */      \
    DLB_API int dlb_conv_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x,                 \
                                        const int32_t* gather, const REAL* delta, const REAL* deriv, const REAL* filters,       \
                                        REAL* g, REAL* dbias, REAL* dfilters, REAL* dx, double scale);                          \
    /* PLayer<D>::act, PLayer.inl:115-138 (replaces CudaTensor::min_max_pool / average_pool, CudaTensor.cu:565-746) */         \
    DLB_API int dlb_pool_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x, REAL* y,         \
                                       int32_t* index_map);                                                                     \
    /* PLayer<D>::backpropagate, PLayer.inl:140-168 (replaces min_max_pool_input_gradient / average_pool_input_gradient) */    \
    DLB_API int dlb_pool_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* delta,             \
                                        const int32_t* index_map, REAL* dx);                                                    \
    /* NLayer<D>::act, NLayer.inl:131-142 (replaces CudaMatrix::mul_add, CudaMatrix.cu:355-365 + activation) */                \
    DLB_API int dlb_full_forward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x,                  \
                                       const int32_t* gather, const REAL* bias, const REAL* weights, REAL* y, REAL* deriv);     \
    /* NLayer<D>::backpropagate, NLayer.inl:144-171 (replaces vector_col_times_vector_row + transpose_mul, CudaMatrix.cu) */   \
    DLB_API int dlb_full_backward_##SFX(dlb_context* ctx, const dlb_layer_spec* spec, int batch, const REAL* x,                 \
                                        const int32_t* gather, const REAL* delta, const REAL* deriv, const REAL* weights,       \
                                        REAL* g, REAL* dbias, REAL* dweights, REAL* dx, double scale);                          \
    /* CostFunction<T>::deriv_in_place, CostFunction.inl:47-51 (replaces CostFunctionHelperCuda::evaluate_gradient) */         \
    DLB_API int dlb_cost_gradient_##SFX(dlb_context* ctx, int cost_id, int batch, int n, const REAL* out, const REAL* ref,      \
                                        const int32_t* gather, REAL* grad);                                                     \
    /* ALayer<D>::update on one container, ALayer.inl:97-122; l_rate is signed as there (Net passes -learning_rate) */      \
    DLB_API int dlb_sgd_update_##SFX(dlb_context* ctx, REAL* params, const REAL* grads, int64_t n, int regularised,             \
                                     double l_rate, double reg_factor, double grad_scale_factor);

DLB_DECLARE_OPS(f32, float)
DLB_DECLARE_OPS(f64, double)

/* ---- the batched Net<GpuDC> --------------------------------------------------------------------------------- */
/* real_bytes: 4 (float) or 8 (double).  Layers are validated like Net<D>::try_load_from_script (Net.inl:83-88). */
DLB_API int dlb_net_create(dlb_context* ctx, const dlb_layer_spec* specs, int n_layers, int real_bytes, dlb_net** out);
DLB_API int dlb_net_destroy(dlb_net* net);
DLB_API int64_t dlb_net_param_count(const dlb_net* net);
DLB_API int dlb_net_layer_count(const dlb_net* net);
DLB_API int64_t dlb_net_layer_in_size(const dlb_net* net, int layer);
DLB_API int64_t dlb_net_layer_out_size(const dlb_net* net, int layer);
DLB_API int64_t dlb_net_layer_param_offset(const dlb_net* net, int layer);
/* host <-> device parameter transfer in the flat layout (GpuDC::from_host / to_host, DataContextCuda.h:40-70) */
DLB_API int dlb_net_set_params(dlb_net* net, const void* host_params);
DLB_API int dlb_net_get_params(dlb_net* net, void* host_params);
DLB_API void* dlb_net_device_params(dlb_net* net);
DLB_API void* dlb_net_device_grads(dlb_net* net);

/* Training data: n items [n][in_size] and references [n][out_size] in the net's Real; copied once into HBM and
 * kept resident (the reference passes std::vector<tensor_t>, Net.inl:193-194). on_device != 0: the pointers are
 * device pointers already and are adopted without a copy (they must outlive the net's use of them). */
DLB_API int dlb_net_set_dataset(dlb_net* net, const void* inputs, const void* refs, int64_t n_items, int on_device);

/* Net<D>::reset_random_generator(seed), Net.inl:599-603: seeds the generator that draws the epoch permutations. */
DLB_API int dlb_net_reset_random_generator(dlb_net* net, unsigned seed);

/* Dropout, ALayer.inl:25-55 (the keep_rate argument of the reference's CLayer / PLayer / NLayer constructors, CLayer.h:145-147,
 * PLayer.h:104-105, NLayer.h:130-132; ALayer::set_keep_rate): a layer with keep_rate < 1 multiplies its input by a 0/1 keep
 * mask (exactly floor(in_size * keep_rate) ones) that is redrawn for every mini-batch and shared by the batch's samples,
 * multiplies its input gradient by the same mask, and scales its input by keep_rate at inference (dlb_net_eval_host,
 * dlb_net_evaluate_host).  The masks are drawn like the reference draws them (Vector::fill_with_random_selection_map on
 * ALayer's generator, which dlb_net_reset_random_generator seeds together with the permutation generator, Net.inl:620-624).
 * A net with dropout runs the layer-by-layer operators (dlb_net_describe_plan says so). */
DLB_API int dlb_net_set_keep_rate(dlb_net* net, int layer, double keep_rate);
/* Elements of one mini-batch's masks: the input sizes of the layers with keep_rate < 1, in layer order. */
DLB_API int64_t dlb_net_mask_size(const dlb_net* net);
/* Explicit keep masks (host, the net's Real, dlb_net_mask_size elements) for dlb_net_forward_backward_host, which - like
 * Net::act_bpg - applies whatever masks are currently set up; the training entry points draw their own per mini-batch. */
DLB_API int dlb_net_set_dropout_masks(dlb_net* net, const void* masks_host);

typedef void (*dlb_epoch_callback)(int64_t epoch, double scaled_l2_reg_factor, void* user);

/* Net<D>::learn, Net.inl:193-324, on the resident data set: per epoch one std::ranges::shuffle of the index map,
 * mini-batches of `batch_size` with the reference's tail rule (fewer than 1.5*batch_size remaining items form one
 * batch), gradient = SUM over the batch, update with -lr/B and reg_factor = lambda / n_items.  Under data
 * parallelism (dlb_net_set_distributed) every rank takes its contiguous slice of each batch and the summed
 * gradients are all-reduced before the (identical) update.  The callback may be NULL. */
DLB_API int dlb_net_learn(dlb_net* net, int64_t batch_size, int64_t epochs, double learning_rate, int cost_id, double lambda,
                          dlb_epoch_callback callback, void* user);

/* Net<D>::learn made resumable: runs the next n_steps mini-batches of the same epoch sequence (a new permutation
 * is drawn whenever an epoch is exhausted; tail rule as in learn) and reports the items consumed over all ranks.
 * learn(batch, epochs) == train_steps until epochs * steps_per_epoch steps have run. */
DLB_API int dlb_net_train_steps(dlb_net* net, int64_t batch_size, int64_t n_steps, double learning_rate, int cost_id, double lambda,
                                int64_t* items_out);

/* Which execution plan learn / train_steps / step_* use for this net: "fused: ..." (fused.cuh) or "unfused: ...". */
DLB_API int dlb_net_describe_plan(dlb_net* net, char* buf, int64_t cap);
/* CUDA-graph replay of the step on the resident-data path (on by default); off = every step launches its kernels. */
DLB_API int dlb_net_enable_graphs(dlb_net* net, int on);

/* fp32 contractions - the NLayer forward / weight-gradient / input-gradient products and, in the layer-by-layer path, the
 * convolutions of multi-channel CLayers as implicit GEMM (forward, filter gradient, input gradient of stride-1 layers):
 * mode 1 (default) = tcgen05 tensor cores with 3xTF32 operand splitting for mini-batches that fill at least one 128-row
 * tensor-core tile (128 samples per GPU) and the packed-FMA (FFMA2) kernels below that; 0 = packed-FMA always; 2 = tcgen05
 * always.  Real = double always uses the FMA path (tcgen05 has no fp64 kind). */
DLB_API int dlb_net_enable_tensor_cores(dlb_net* net, int mode);

/* The whole mini-batch step of Net<D>::learn (Net.inl:226-317) as ONE persistent cooperative kernel (csrc/mega.cu) for float nets of
 * LeNet's shape (CLayer 5x5 over one 28x28 channel -> PLayer 2x2 max|min -> NLayer -> NLayer head); dlb_net_train_steps / dlb_net_learn
 * then run all consecutive full mini-batches of an epoch inside a single launch.  mode 0 (default) = the multi-kernel CUDA-graph step,
 * 1 = the step kernel where the net qualifies.  Same results within the parity bar either way; measured on B200 the two are within
 * 5 % of each other at mini-batch 256, so the graph path stays the default. */
DLB_API int dlb_net_enable_step_kernel(dlb_net* net, int mode);

/* Dense fp32 contraction C[M][N] = sum_k A(m,k) * B(n,k) (replaces CudaMatrix::mul_add / transpose_mul /
 * vector_col_times_vector_row style products, CudaMatrix.cu:227-498, for whole batches).  a_transposed / b_transposed:
 * the operand is stored [K][M] / [K][N] with leading dimension lda / ldb.  use_tensor_cores selects tcgen05 (3xTF32). */
DLB_API int dlb_matmul_f32(dlb_context* ctx, int M, int N, int K, const float* A, int64_t lda, int a_transposed, const float* B,
                           int64_t ldb, int b_transposed, float* C, int use_tensor_cores);
/* The same for Real = double (DFMA micro-kernels; tcgen05 has no fp64 kind). */
DLB_API int dlb_matmul_f64(dlb_context* ctx, int M, int N, int K, const double* A, int64_t lda, int a_transposed, const double* B,
                           int64_t ldb, int b_transposed, double* C);

/* One mini-batch step on explicit item indices of the resident data set (host int64 indices, n of them). */
DLB_API int dlb_net_step_indices(dlb_net* net, const int64_t* indices, int64_t n, double learning_rate, int cost_id, double reg_factor);

/* One mini-batch step on HOST buffers (inputs [n][in], refs [n][out], pinned or pageable): copies them to the
 * device, runs forward + backward + update and, if `cost_out` is not NULL, copies back the summed cost of the batch
 * evaluated before the update (a device->host read that synchronises). This is the end-to-end entry point. */
DLB_API int dlb_net_step_host(dlb_net* net, const void* inputs, const void* refs, int64_t n, double learning_rate, int cost_id,
                              double reg_factor, double* cost_out);

/* The same step without the final synchronisation: the host->device copy of this batch (on a copy stream, into the other
 * half of a staging double buffer), the step and the device->host copy of its cost into *cost_out (pinned memory) are
 * left in flight, so that consecutive calls overlap copy and compute.  The host buffers must stay valid and unchanged until
 * dlb_context_synchronize(); *cost_out is defined after it. */
DLB_API int dlb_net_step_host_async(dlb_net* net, const void* inputs, const void* refs, int64_t n, double learning_rate, int cost_id,
                                    double reg_factor, double* cost_out);

/* Forward + cost derivative + backward on a host batch WITHOUT the update; the traces stay on the device for
 * dlb_net_read.  `scale` is applied to the previous content of the gradient buffer (gradient_scale_factor). */
DLB_API int dlb_net_forward_backward_host(dlb_net* net, const void* inputs, const void* refs, int64_t n, int cost_id, double scale);

enum {
    DLB_BUF_OUTPUT = 1,      /* [n][out_size(layer)]  output of `layer` (post-activation)            */
    DLB_BUF_DERIVATIVES = 2, /* [n][out_size(layer)]  Trace.Derivatives (softmax: exp(x - max))       */
    DLB_BUF_INDEX_MAP = 3,   /* [n][out_size(layer)]  int32 Trace.IndexMapping of a MIN/MAX PLayer    */
    DLB_BUF_INPUT_GRAD = 4,  /* [n][in_size(layer)]   gradient w.r.t. the input of `layer` (layer>=1) */
    DLB_BUF_COST_GRAD = 5,   /* [n][out_size(last)]   cost derivative fed to the last layer           */
    DLB_BUF_GRADIENT = 6     /* [param_count]         flat summed gradient                            */
};
/* Copies a trace buffer of the last forward/backward (rows = its batch size) to the host; synchronises. */
DLB_API int dlb_net_read(dlb_net* net, int what, int layer, void* host_dst, int64_t dst_bytes);
DLB_API int dlb_net_write_gradient(dlb_net* net, const void* host_src);

/* Net<D>::act + max_element_id (Net.inl:134-155, 432-433) over n host items; classes and/or outputs may be NULL. */
DLB_API int dlb_net_eval_host(dlb_net* net, const void* inputs, int64_t n, int32_t* classes, void* outputs);

/* Net<D>::evaluate_cost_function_and_correct_answers (Net.inl:411-439) over n host items: *cost_out = mean cost +
 * l2_reg_factor * sum of squared weights (NLayer weights and CLayer filters), *correct_out = fraction of items whose
 * arg-max matches the reference's. */
DLB_API int dlb_net_evaluate_host(dlb_net* net, const void* inputs, const void* refs, int64_t n, int cost_id, double l2_reg_factor,
                                  double* cost_out, double* correct_out);

/* Data parallelism: one process per GPU.  nccl_unique_id: the 128-byte ncclUniqueId created by rank 0
 * (dlb_nccl_get_unique_id) and distributed by the launcher's own plumbing (torch.distributed, MPI, a file, ...). */
DLB_API int dlb_nccl_get_unique_id(void* id_out_128_bytes);
DLB_API int dlb_net_set_distributed(dlb_net* net, int rank, int world_size, const void* nccl_unique_id);

/* The contiguous share [first, first + mine) of a global mini-batch of `count` items that `rank` of `world` takes
 * (host-only helper; the data-parallel partitioning rule used by learn / train_steps / step_indices). */
DLB_API int dlb_rank_slice(int64_t count, int rank, int world, int64_t* first, int64_t* mine);

/* CUDA events on the context's stream (device-side timing for callers that do not link the CUDA runtime). */
DLB_API int dlb_event_create(void** event_out);
DLB_API int dlb_event_destroy(void* event);
DLB_API int dlb_event_record(dlb_context* ctx, void* event);
DLB_API int dlb_event_elapsed_ms(void* start, void* stop, double* ms_out); /* synchronises on `stop` */

/* Per-kernel device timing (measurement aid for bench.py's roofline pass; off by default).  The report is text:
 * one "<layer>:<operator>:<kernel>\t<launches>\t<total ms>" line per kernel; reading it clears the records. */
DLB_API int dlb_timing_enable(int on);
DLB_API int dlb_timing_report(char* buf, int64_t cap);

/* Data parallelism over NVLink peer memory instead of NCCL: every rank exports the CUDA IPC handle (64 bytes) of its
 * gradient-exchange region, the launcher gathers the `world` handles (index = rank), every rank connects.  From then on the
 * step ends with ONE kernel per rank that barriers with its peers through flags in peer memory, sums all ranks' gradients
 * through peer pointers in rank order (bit-identical on every rank) and applies the update in the same pass.  All ranks
 * must be on one node with peer access between their devices; call after every rank has called dlb_net_p2p_export. */
DLB_API int dlb_net_p2p_export(dlb_net* net, void* handle_out_64_bytes);
DLB_API int dlb_net_set_distributed_p2p(dlb_net* net, int rank, int world_size, const void* handles_world_x_64_bytes);

/* The epoch permutations learn() would draw after reset_random_generator(seed): iota + one shuffle per epoch on
 * the same std::mt19937 (Net.inl:158-171, 223).  Host-only helper for tests and for reference-side callers. */
DLB_API int dlb_shuffle_sequence(unsigned seed, int64_t n, int64_t epochs, int64_t* out);

/* The keep masks the training entry points draw after dlb_net_reset_random_generator(seed): per mini-batch, per layer with
 * keep_rate < 1 in layer order, Vector::fill_with_random_selection_map over the layer's input (ALayer.inl:26-35,
 * Vector.cpp:208-233).  out: n_batches blocks of sum(in_sizes[l] for keep_rates[l] < 1) elements of float (real_bytes 4)
 * or double (8).  Host-only helper for tests and for reference-side callers. */
DLB_API int dlb_dropout_masks(unsigned seed, const int64_t* in_sizes, const double* keep_rates, int n_layers, int64_t n_batches,
                              int real_bytes, void* out);

/* ---- the recurrent stack: MNet<D> of RMLayer<D> (csrc/mnet.cu) -------------------------------------
 * Replaces, batched over the mini-batch on the device: RMLayer::act / backpropagate / update (NeuralNet/RMLayer.inl:165-260: Elman
 * recurrence h_t = f(W_in x_t + b + W_rec h_{t-1}), full back-propagation through time, gradient-norm clipping in the update) and
 * MNet::act / calc_gradient_sum / learn (NeuralNet/MNet.inl:79-180).  Sequences are [n][depth][dim] arrays of the net's Real on the
 * host (one sequence after the other).  Flat parameter (and gradient) layout, per layer:  bias [out] | in-weights [out][in] |
 * recurrent weights [out][out], row-major - the order of RMLayer's gradient containers (RMLayer.h:120-122).
 * Element-wise activations only (DLB_ACT_SIGMOID / TANH / RELU / LINEAR). */
typedef struct dlb_mnet dlb_mnet;
typedef struct dlb_rmlayer_spec {
    int32_t in_dim, out_dim;      /* items per time step in / out (RMLayer(rec_depth, in_sub_dim, out_sub_dim, ...), RMLayer.h:131-133) */
    int32_t activation;           /* DLB_ACT_* */
    int32_t reserved;
    double clip_threshold;        /* gradient-norm clipping threshold of the update, <= 0 disables it (RMLayer.inl:238-245) */
} dlb_rmlayer_spec;
DLB_API const char* dlb_mnet_last_error(void);
DLB_API int dlb_mnet_create(dlb_context* ctx, int depth, const dlb_rmlayer_spec* layers, int n_layers, int real_bytes, dlb_mnet** out);
DLB_API int dlb_mnet_destroy(dlb_mnet* net);
DLB_API int64_t dlb_mnet_param_count(const dlb_mnet* net);
DLB_API int64_t dlb_mnet_layer_param_offset(const dlb_mnet* net, int layer);
DLB_API int dlb_mnet_set_params(dlb_mnet* net, const void* host_params);
DLB_API int dlb_mnet_get_params(dlb_mnet* net, void* host_params);
/* MNet::act for n sequences (MNet.inl:79-98): outputs [n][depth][out] */
DLB_API int dlb_mnet_act_host(dlb_mnet* net, const void* inputs, int64_t n, void* outputs);
/* MNet::calc_gradient_sum (MNet.inl:100-121): the gradient SUM over the n sequences, flat layout */
DLB_API int dlb_mnet_gradient_sum_host(dlb_mnet* net, const void* inputs, const void* refs, int64_t n, int cost_id, void* host_grads);
/* MNet::learn on one mini-batch (MNet.inl:171-180): gradient sum, every layer's update with -learning_rate / n */
DLB_API int dlb_mnet_learn_host(dlb_mnet* net, const void* inputs, const void* refs, int64_t n, int cost_id, double learning_rate, double reg_factor);
/* the same with the mini-batch resident on the device: upload once, then n_steps x (forward, backward, update) without host traffic */
DLB_API int dlb_mnet_set_batch(dlb_mnet* net, const void* inputs, const void* refs, int64_t n);
DLB_API int dlb_mnet_learn_resident(dlb_mnet* net, int cost_id, double learning_rate, double reg_factor, int64_t n_steps);

#ifdef __cplusplus
}
#endif
#endif /* DLB200_H */
