// tools/flag_probe.cu — two GPUs of one box: round-trip time of a flag exchanged through peer memory the way p2p_reduce_update_kernel
// does it (volatile store of a ticket into the peer's memory after __threadfence_system, volatile spin on the local flag), with one
// polling block and with 296 polling blocks of 8 polling threads each.  Ping-pong of `iters` tickets; prints microseconds per one-way hop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/flag_probe tools/flag_probe.cu && tools/_build/flag_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

// side 0 sends odd tickets and waits for even ones, side 1 the other way round; only block 0 sends, every block polls
__global__ void pingpong(volatile unsigned long long* mine, volatile unsigned long long* theirs, int side, int iters, int pollers, unsigned long long* ns_out) {
    unsigned long long t0 = 0, t1 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < iters; ++i) {
        const unsigned long long send = 2ull * i + 1 + side, expect = 2ull * i + 2 - side;   // side 0: sends 1,3,.. waits 2,4,..; side 1: waits 1,3,.. sends 2,4,..
        if (side == 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *theirs = send; }
            if ((int)threadIdx.x < pollers) { while (*mine < expect) {} __threadfence_system(); }
        } else {
            if ((int)threadIdx.x < pollers) { while (*mine < send - 1) {} __threadfence_system(); }
            __syncthreads();
            if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *theirs = send; }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); *ns_out = t1 - t0; }
}

int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    unsigned long long *f[2], *ns[2];
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&f[d], 256)); CK(cudaMemset(f[d], 0, 256)); CK(cudaMalloc(&ns[d], 8)); CK(cudaStreamCreate(&st[d]));
    }
    const int iters = 2000;
    for (int blocks : {1, 296}) {
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(f[d], 0, 256)); CK(cudaDeviceSynchronize()); }
        // side 1's first wait is for ticket 1 - 1 = 0... shift: side 1 waits for `send - 1` = odd ticket of the same round
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); pingpong<<<blocks, 256, 0, st[d]>>>(f[d], f[1 - d], d, iters, 8, ns[d]); }
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
        unsigned long long h = 0; CK(cudaSetDevice(0)); CK(cudaMemcpy(&h, ns[0], 8, cudaMemcpyDeviceToHost));
        printf("%3d polling blocks x 8 threads: %.2f us per one-way flag hop (%d round trips)\n", blocks, 1e-3 * (double)h / (2.0 * iters), iters);
    }
    return 0;
}
