// Shared helpers of the opencmp_b200 CUDA sources.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

int ocmp_fail(int code, const char* msg);
int ocmp_check(const char* where);
int ocmp_sm_count();
int ocmp_patch_invert_registers(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                                const double* fm, double* inv, int* flag_dev, const int* pos, cudaStream_t st);
// smoother application: y[p, :] = A_p^-1 r[dofs_p] (bulk-copy streamed, FP64 / FP32 / bfloat16 stored inverses), then
// the deterministic per-dof gather z (+)= scale * w * m * sum of the patch-local entries
int ocmp_patch_apply_y(int npatch, int bs, const int* pd, const double* inv, const double* r, double* y,
                       cudaStream_t st);
int ocmp_patch_apply_y_f32(int npatch, int bs, const int* pd, const float* inv, const double* r, double* y,
                           cudaStream_t st);
int ocmp_patch_apply_y_bf16(int npatch, int bs, const int* pd, const __nv_bfloat16* inv, const double* r, double* y,
                            cudaStream_t st);
int ocmp_patch_gather(long long n, const int* inc_ptr, const int* inc_idx, const double* y, const double* w,
                      const double* m, double scale, double* z, int accumulate, cudaStream_t st);
int ocmp_patch_invert_registers_f32(int npatch, int bs, const int* pd, const int* rp, const int* ci,
                                    const double* vals, const double* fm, float* inv, int* flag_dev, const int* pos,
                                    cudaStream_t st);
int ocmp_patch_invert_registers_bf16(int npatch, int bs, const int* pd, const int* rp, const int* ci,
                                     const double* vals, const double* fm, __nv_bfloat16* inv, int* flag_dev,
                                     const int* pos, cudaStream_t st);
// conversions between FP64 arithmetic and the storage type of the preconditioner data
template <typename T> __device__ __forceinline__ T ocmp_store(double v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 ocmp_store<__nv_bfloat16>(double v) { return __double2bfloat16(v); }
template <typename T> __device__ __forceinline__ double ocmp_load(const T* p) { return (double)__ldg(p); }
template <> __device__ __forceinline__ double ocmp_load<__nv_bfloat16>(const __nv_bfloat16* p) {
    return (double)__bfloat162float(__ldg(p));
}
// optional per-category device timing (CUDA events on the launching stream) and launch counting
enum { PROF_SPMV = 0, PROF_ASM_APPLY, PROF_COEF, PROF_CONTRACT, PROF_LIN, PROF_MDOT, PROF_MAXPY, PROF_VEC,
       PROF_SETUP, PROF_SPMV_MG, PROF_HALO, PROF_NCAT };
void ocmp_prof_begin(int cat, cudaStream_t st);
void ocmp_prof_end(int cat, cudaStream_t st);
void ocmp_prof_bytes(int cat, double bytes);
struct ProfScope {
    int cat; cudaStream_t st;
    ProfScope(int c, cudaStream_t s) : cat(c), st(s) { ocmp_prof_begin(c, s); }
    ~ProfScope() { ocmp_prof_end(cat, st); }
};

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
