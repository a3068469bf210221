// Batched dense patch kernels for the additive-Schwarz smoother (sm_100a).
//
//   k_patch_invert<T>   one CTA (16 x 16 threads) per patch: the (16T x 16T padded) patch matrix lives entirely in
//                       registers, cyclically distributed (thread (ty,tx) owns rows ty+16a, columns tx+16b); each
//                       Gauss-Jordan step broadcasts the pivot row / column through shared memory and every thread
//                       does T*T FP64 FMAs on its tile. No pivoting: patch dofs are listed in ascending global order,
//                       so velocity dofs precede pressure dofs and the saddle-point blocks are quasi-definite; a
//                       vanishing pivot raises a flag and the host falls back to the pivoted shared-memory kernel.
//   k_patch_apply       one CTA per patch: z[dofs] += A_p^-1 r[dofs], columns split over the warps, rows over lanes.
//
// Stand in for the block inversions / block solves of NGSolve's block-Jacobi and multigrid smoothers behind
// ngs.Preconditioner(...).Update() and its application inside the Krylov loop
// (reference opencmp/models/base_model.py:365-383, opencmp/solvers/base_solver.py:711-719).
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/opencmp_b200.h"
#include "ocmp_common.cuh"

template <int T>
__global__ void __launch_bounds__(256, 1) k_patch_invert(int npatch, int bs, const int* __restrict__ pdofs,
                                                         const int* __restrict__ rowptr,
                                                         const int* __restrict__ colidx,
                                                         const double* __restrict__ vals,
                                                         const double* __restrict__ fm, double* __restrict__ inv,
                                                         int* __restrict__ flag) {
    constexpr int NP = 16 * T;
    __shared__ double rowk[2][NP];
    __shared__ double colk[2][NP];
    __shared__ int sd[NP];
    __shared__ unsigned char sfree[NP];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < NP; i += 256) {
            const int d = i < bs ? __ldg(pdofs + (long long)p * bs + i) : -1;
            sd[i] = d;
            sfree[i] = (d >= 0 && (!fm || __ldg(fm + d) > 0.0)) ? 1 : 0;
        }
        __syncthreads();
        double M[T][T];
#pragma unroll
        for (int a = 0; a < T; ++a) {
            const int i = ty + 16 * a;
            const int di = sd[i];
            const bool fi = sfree[i];
            int lo0 = 0, hi0 = -1;
            if (fi) { lo0 = __ldg(rowptr + di); hi0 = __ldg(rowptr + di + 1) - 1; }
#pragma unroll
            for (int b = 0; b < T; ++b) {
                const int j = tx + 16 * b;
                double v = (i == j) ? 1.0 : 0.0;
                if (fi && sfree[j]) {
                    const int dj = sd[j];
                    v = 0.0;
                    int lo = lo0, hi = hi0;
                    while (lo <= hi) {
                        const int mid = (lo + hi) >> 1;
                        const int c = __ldg(colidx + mid);
                        if (c == dj) { v = __ldg(vals + mid); break; }
                        if (c < dj) lo = mid + 1; else hi = mid - 1;
                    }
                }
                M[a][b] = v;
            }
        }
        bool bad = false;
#pragma unroll
        for (int ka = 0; ka < T; ++ka) {
            for (int kr = 0; kr < 16; ++kr) {
                const int k = kr + 16 * ka;
                const int buf = kr & 1;
                if (ty == kr) {
#pragma unroll
                    for (int b = 0; b < T; ++b) rowk[buf][tx + 16 * b] = M[ka][b];
                }
                if (tx == kr) {
#pragma unroll
                    for (int a = 0; a < T; ++a) colk[buf][ty + 16 * a] = M[a][ka];
                }
                __syncthreads();
                const double d = rowk[buf][k];
                if (!(fabs(d) > 1e-280)) bad = true;
                const double ip = 1.0 / d;
                double ci[T], rj[T];
#pragma unroll
                for (int a = 0; a < T; ++a) ci[a] = colk[buf][ty + 16 * a] * ip;
#pragma unroll
                for (int b = 0; b < T; ++b) rj[b] = rowk[buf][tx + 16 * b];
                if (ty == kr) ci[ka] = 0.0;              // pivot row and column are rewritten below
                if (tx == kr) rj[ka] = 0.0;
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) M[a][b] = fma(-ci[a], rj[b], M[a][b]);
                if (ty == kr) {
#pragma unroll
                    for (int b = 0; b < T; ++b) M[ka][b] = rowk[buf][tx + 16 * b] * ip;
                }
                if (tx == kr) {
#pragma unroll
                    for (int a = 0; a < T; ++a) M[a][ka] = -colk[buf][ty + 16 * a] * ip;
                }
                if (ty == kr && tx == kr) M[ka][ka] = ip;
                // double-buffered row/column staging: one barrier per step is enough
            }
        }
        if (bad) *flag = 1;
        double* out = inv + (long long)p * bs * bs;
#pragma unroll
        for (int a = 0; a < T; ++a) {
            const int i = ty + 16 * a;
#pragma unroll
            for (int b = 0; b < T; ++b) {
                const int j = tx + 16 * b;
                if (i < bs && j < bs) out[(long long)j * bs + i] = M[a][b];
            }
        }
    }
}

template <int T>
static void launch_invert(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                          const double* fm, double* inv, int* flag, cudaStream_t st) {
    const int cap = ocmp_sm_count();
    k_patch_invert<T><<<npatch < cap ? npatch : cap, 256, 0, st>>>(npatch, bs, pd, rp, ci, vals, fm, inv, flag);
}

// returns 1 if handled (flag_dev is set to 1 on a vanishing pivot), 0 if the patch is too large for this kernel
int ocmp_patch_invert_registers(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                                const double* fm, double* inv, int* flag_dev, cudaStream_t st) {
    const int T = (bs + 15) / 16;
    switch (T) {
        case 1: launch_invert<1>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 2: launch_invert<2>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 3: launch_invert<3>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 4: launch_invert<4>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 5: launch_invert<5>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 6: launch_invert<6>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 7: launch_invert<7>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 8: launch_invert<8>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 9: launch_invert<9>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        case 10: launch_invert<10>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, st); return 1;
        default: return 0;
    }
}

// ---- application ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_patch_apply(int npatch, int bs, const int* __restrict__ pdofs,
                                                     const double* __restrict__ inv, const double* __restrict__ r,
                                                     double* __restrict__ z) {
    extern __shared__ double sm[];         // r_loc[bs], partial[8][bs]
    double* rl = sm;
    double* part = sm + bs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int* d = pdofs + (long long)p * bs;
        __syncthreads();
        for (int j = threadIdx.x; j < bs; j += 256) {
            const int dj = __ldg(d + j);
            rl[j] = dj >= 0 ? __ldg(r + dj) : 0.0;
        }
        __syncthreads();
        const double* A = inv + (long long)p * bs * bs;
        double s[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) s[m] = 0.0;
        for (int j = warp; j < bs; j += 8) {
            const double rj = rl[j];
            const double* col = A + (long long)j * bs;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = lane + 32 * m;
                if (i < bs) s[m] = fma(__ldg(col + i), rj, s[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = lane + 32 * m;
            if (i < bs) part[warp * bs + i] = s[m];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < bs; i += 256) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += part[w * bs + i];
            const int di = __ldg(d + i);
            if (di >= 0) atomicAdd(z + di, t);
        }
    }
}

int ocmp_patch_apply_cta(int npatch, int bs, const int* pd, const double* inv, const double* r, double* z,
                         cudaStream_t st) {
    if (bs > 256) return 0;
    const size_t smem = sizeof(double) * 9 * bs;
    const int cap = ocmp_sm_count() * 8;
    k_patch_apply<<<npatch < cap ? npatch : cap, 256, smem, st>>>(npatch, bs, pd, inv, r, z);
    return 1;
}
