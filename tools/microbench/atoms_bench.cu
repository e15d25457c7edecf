// Microbenchmark: throughput of conflict-free shared-memory atomic adds (RED, no return value) per SM,
// next to a plain LDS+IADD+STS read-modify-write, to decide between a scatter and a gather pileup.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned *out, int iters) {
    __shared__ unsigned h[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned row = threadIdx.x >> 5;
    for (int i = 0; i < iters; ++i) {
        row = (row * 5 + 3) & 63;
        unsigned *p = &h[row * 32 + lane];
        if (MODE == 0) atomicAdd(p, 1u << ((i & 3) * 8));
        else if (MODE == 1) *p += 1u << ((i & 3) * 8);
        else { atomicAdd(p, 1u << ((i & 3) * 8)); atomicAdd(&h[((row + 7) & 63) * 32 + lane], (unsigned)i); }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) s += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    unsigned *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters); else k<2><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double warp_ops = 148.0 * 8 * 8 * iters * (mode == 2 ? 2 : 1);
        printf("mode %d (%s): %.3f ms, %.2f warp-ops/cycle/SM at 1.965 GHz\n", mode, mode == 0 ? "RED.ADD" : mode == 1 ? "LDS+ADD+STS" : "2x RED.ADD",
               ms, warp_ops / 148.0 / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
