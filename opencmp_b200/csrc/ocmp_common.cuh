// Shared helpers of the opencmp_b200 CUDA sources.
#pragma once
#include <cuda_runtime.h>

int ocmp_fail(int code, const char* msg);
int ocmp_check(const char* where);
int ocmp_sm_count();

__device__ __forceinline__ double warp_reduce_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; result valid in thread 0
__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_reduce_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) v = warp_reduce_sum(v);
    return v;
}
