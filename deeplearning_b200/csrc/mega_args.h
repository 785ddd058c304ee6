// mega_args.h — kernel-side argument block and workspace layout of the step megakernel (included by mega.cu only, so
// that work on the kernel does not recompile the rest of the library).
#pragma once

#include "mega.h"

namespace dlb {
namespace mega {

constexpr int kMaxPeers = 16;

// Phases (bit mask): a launch runs the phases in its mask in order, with a grid barrier between consecutive ones.
enum Phase : int {
    kConvFwd = 1,     // CLayer::act + PLayer::act                                       (filter-pair major)
    kFc1Fwd = 2,      // hidden NLayer forward as per-filter partial pre-activations     (filter major, tcgen05)
    kHead = 4,        // partial sum + bias + activation, head forward, cost derivative, head backward
    kFc1Bwd = 8,      // hidden NLayer input gradient and weight-gradient partials       (filter major, tcgen05)
    kConvBwd = 16,    // PLayer + CLayer backward: filter and untied-bias gradient partials
    kUpdate = 32,     // ordered sum of the partials (+ the other ranks' sums) and the SGD update
    kAll = 63
};

struct Args {
    // ---- the mini-batch ----
    const float* x;             // [rows][784] items the gather indexes into
    const float* refs;          // [rows][n_out]
    const int* gather_base;     // nullable: batch slot -> row (epoch permutation slice of this rank)
    const int* cursor;          // nullable: device-side offset added to gather_base
    int* cursor_rw;             // nullable: advanced by `advance` at the end of the step
    int advance;
    int B;                      // samples of this rank
    // ---- the net ----
    int F, H1, n_out;           // filters, hidden NLayer outputs, head outputs
    int act0, is_max, act1, act2, cost_id;
    float* params;              // flat parameter buffer (the reference's container order)
    float* grads;               // flat gradient buffer (receives the batch sum)
    long long off_b0, off_k0, off_b1, off_w1, off_b2, off_w2, n_params;
    // ---- workspace (per net, sized for `cap` samples) ----
    float* pooled;              // [cap][F * 144]   PLayer output
    unsigned char* sel;         // [cap][F * 144]   arg-max position inside the 2 x 2 window (0..3, 255: none)
    float* z1part;              // [F][cap][H1]     per-filter partial pre-activations of the hidden NLayer
    float* g1;                  // [cap][128]       dC/dz of the hidden NLayer (columns >= H1 stay zero)
    float* dpool;               // [cap][F * 144]   gradient w.r.t. the PLayer output
    float* dw1part;             // [groups_f][H1][F * 144]
    float* db1part;             // [groups_f][128]
    float* convpart;            // [groups_c][F * 576 + F * 25]
    float* headpart;            // [grid][n_out * (H1 + 1)]   (bias gradient first, then the weight gradient)
    double* costpart;           // [grid]
    double* cost_out;           // nullable: receives the batch cost
    unsigned long long* barrier;   // grid-barrier counter (monotonic)
    // ---- decomposition ----
    int grid, groups_c, groups_f;  // CTAs; sample groups of the convolution phases (x F/2 filter pairs) and of the NLayer phases (x F filters)
    // ---- update ----
    float param_scale, grad_scale; // w <- w * param_scale + g * grad_scale (ALayer.inl:108-121)
    int nontrivial_reg;
    int do_update;
    // ---- data parallelism over peer memory (world > 1) ----
    int world, rank;
    const unsigned char* peer_region[kMaxPeers];   // [publish 0 | publish 1 | flags]
    size_t publish_bytes;
    const unsigned long long* step_counter;        // device: which half is published this step (= barrier ticket - 1)
    unsigned long long* step_counter_rw;
    int phase_mask;
    int flags;                  // development switches: bit 0 = no W1 prefetch during the convolution forward
    int n_steps;                // steps of identical shape run by this launch (the cursor advances on the device)
    long long* prof;            // nullable (development): [grid][16] clock64 stamps of CTA thread 0 at the phase boundaries
};

// Workspace sizes (elements) for a batch capacity and a launch decomposition.
struct Layout {
    int grid, groups_c, groups_f;
    size_t pooled, sel, z1part, g1, dpool, dw1part, db1part, convpart, headpart, costpart;
};
Layout make_layout(const Shape& s, int cap, int n_sms);

int launch(const Args& args, cudaStream_t stream);
size_t smem_bytes(int H1);

}  // namespace mega
}  // namespace dlb
