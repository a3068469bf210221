// Sparse direct solve behind `a.mat.Inverse(freedofs)` — the reference's default `linear_solver = direct`
// (reference opencmp/models/base_model.py:908-922: UMFPACK / PARDISO through NGSolve).
//
// The free-free part of the CSR matrix is permuted by a bandwidth-reducing ordering (reverse Cuthill-McKee, computed
// once per pattern on the host) and factorised as a BAND matrix with partial pivoting, LAPACK dgbtrf's storage and
// semantics: column j of the band array holds A(j-kl-ku .. j+kl, j), the top kl rows take the fill that row
// interchanges create, L keeps its multipliers unswapped and the interchanges are replayed on the right-hand side.
//
//   k_band_fill    CSR -> band scatter through the permutation (16 lanes per row)
//   k_band_lu      ONE persistent cooperative kernel for the whole factorisation: per column one grid barrier;
//                  CTA 0 runs one column ahead (updates column j+1, finds its pivot, scales it) while the other CTAs
//                  apply the rank-1 update of column j to the rest of the window, one warp per column
//   k_band_solve   forward + backward substitution in one CTA with the active window of the right-hand side in
//                  shared memory (circular buffer), the band columns prefetched into registers one step ahead
//
// Work: n * kl * (kl + ku) FMAs, L2-resident window of (kl + ku) columns — fine for the sizes a direct solve is asked
// for (examples, parity cases: 1e3 .. 3e5 DOFs); larger systems go through the Krylov path (backend.solve_free).
#include <cooperative_groups.h>
#include <cstdio>
#include "ocmp_common.cuh"
#include "../../include/opencmp_b200.h"

namespace cg = cooperative_groups;

__global__ void k_band_fill(int nrows, const int* __restrict__ rp, const int* __restrict__ ci,
                            const double* __restrict__ vals, const int* __restrict__ perm, int kl, int ku,
                            double* __restrict__ ab) {
    const int lane = threadIdx.x & 15;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (row >= nrows) return;
    const int pi = perm[row];
    if (pi < 0) return;
    const int kv = kl + ku;
    const long long ld = 2LL * kl + ku + 1;
    for (int k = rp[row] + lane; k < rp[row + 1]; k += 16) {
        const int pc = perm[ci[k]];
        if (pc >= 0) ab[(long long)pc * ld + kv + pi - pc] = vals[k];
    }
}

__global__ void k_band_gather(int nrows, const int* __restrict__ perm, const double* __restrict__ r,
                              double* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nrows && perm[i] >= 0) b[perm[i]] = r[i];
}

// out[i] = (accumulate ? out[i] : 0) + x[perm[i]] on the free entries; constrained entries: untouched / 0
__global__ void k_band_scatter(int nrows, const int* __restrict__ perm, const double* __restrict__ x,
                               double* __restrict__ out, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const int p = perm[i];
    if (p >= 0) out[i] = (accumulate ? out[i] : 0.0) + x[p];
    else if (!accumulate) out[i] = 0.0;
}

// ---- factorisation ---------------------------------------------------------------------------------------------
// Column j is "final" once its pivot row is known (ipiv[j]), the pivot sits on the diagonal and the sub-diagonal
// holds the multipliers. info[0]: 1-based index of the first zero pivot (0 = none); info[1]: widest U row seen.
__device__ void band_finalize_column(int j, int n, int kl, int kv, long long ld, double* ab, int* ipiv, int* info) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    double* col = ab + (long long)j * ld + kv;
    const int km = min(kl, n - 1 - j);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5;
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = tid; i <= km; i += nt) {
        const double a = fabs(col[i]);
        if (a > best) { best = a; bi = i; }          // a thread sees its rows in increasing order: first maximum kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_val[w] = best; s_idx[w] = bi; }
    __syncthreads();
    if (w == 0) {
        const int nw = (nt + 31) >> 5;
        best = lane < nw ? s_val[lane] : -1.0;
        bi = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            if (!(best > 0.0)) {                      // exactly singular (or NaN): LAPACK's info > 0
                bi = 0;
                if (info[0] == 0) info[0] = j + 1;
            }
            ipiv[j] = j + bi;
            const double p = col[bi];
            col[bi] = col[0];
            col[0] = p;
        }
    }
    __syncthreads();
    const double piv = col[0];
    if (piv != 0.0)
        for (int i = 1 + tid; i <= km; i += nt) col[i] /= piv;
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_band_lu(int n, int kl, int ku, double* ab, int* ipiv, int* info) {
    extern __shared__ double sl[];                    // multipliers of the current column
    cg::grid_group grid = cg::this_grid();
    const int kv = kl + ku;
    const long long ld = 2LL * kl + ku + 1;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int gw = (blockIdx.x * nt + tid) >> 5, ngw = (gridDim.x * nt) >> 5;
    if (blockIdx.x == 0) band_finalize_column(0, n, kl, kv, ld, ab, ipiv, info);
    grid.sync();
    int ju = 0, umax = 0;
    for (int j = 0; j < n - 1; ++j) {
        const int km = min(kl, n - 1 - j);
        const double* colj = ab + (long long)j * ld + kv;
        const int jp = ipiv[j] - j;
        const bool regular = colj[0] != 0.0;
        if (regular) ju = max(ju, min(j + ku + jp, n - 1));
        umax = max(umax, ju - j);
        if (regular) {
            for (int i = tid; i < km; i += nt) sl[i] = colj[1 + i];
            __syncthreads();
        }
        if (blockIdx.x == 0) {
            // look-ahead: the next pivot column first, with the whole CTA
            if (regular && j + 1 <= ju) {
                double* cc = ab + (long long)(j + 1) * ld + kv - 1;        // cc[i] = A(j + i, j + 1)
                const double t = cc[0], s = cc[jp];
                __syncthreads();
                for (int i = 1 + tid; i <= km; i += nt) cc[i] = (i == jp ? t : cc[i]) - sl[i - 1] * s;
                if (tid == 0) cc[0] = s;
                __syncthreads();
            }
            band_finalize_column(j + 1, n, kl, kv, ld, ab, ipiv, info);
        }
        if (regular) {
            for (int c = j + 2 + gw; c <= ju; c += ngw) {
                double* cc = ab + (long long)c * ld + kv + j - c;           // cc[i] = A(j + i, c)
                const double t = cc[0], s = cc[jp];
                __syncwarp();
                for (int i = 1 + lane; i <= km; i += 32) cc[i] = (i == jp ? t : cc[i]) - sl[i - 1] * s;
                if (lane == 0) cc[0] = s;
            }
        }
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0) info[1] = max(umax, min(ku, n - 1));
}

// ---- substitution ------------------------------------------------------------------------------------------------
// One CTA. The right-hand side lives in global memory (in place); a circular window of W slots in shared memory holds
// the entries the next `AHEAD` steps touch. PF = band entries per thread kept in registers for the step in flight.
#define BAND_AHEAD 32
template <int PF>
__global__ void __launch_bounds__(1024) k_band_solve(int n, int kl, int ku, int ubw, const double* __restrict__ ab,
                                                     const int* __restrict__ ipiv, double* __restrict__ b) {
    extern __shared__ double w[];
    const int kv = kl + ku;
    const long long ld = 2LL * kl + ku + 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    double lr[PF];
    // ---- forward: L y = P b -----------------------------------------------------------------------------------
    if (kl > 0) {
        const int W = kl + 1 + 2 * BAND_AHEAD;
        for (int i = tid; i < min(n, kl + 1 + BAND_AHEAD); i += nt) w[i % W] = b[i];
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const int km = min(kl, n - 1 - j);
            const double* col = ab + (long long)j * ld + kv;
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = 1 + tid + k * nt;
                lr[k] = i <= km ? __ldg(col + i) : 0.0;
            }
            const int p = ipiv[j];
            const double a = w[j % W], bp = w[p % W];
            if (p != j) __syncthreads();               // every thread holds a / bp before slot p is overwritten
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = 1 + tid + k * nt;
                if (i <= km) {
                    const int s = (j + i) % W;
                    w[s] = ((j + i) == p ? a : w[s]) - lr[k] * bp;
                }
            }
            if (tid == 0) b[j] = bp;
            if ((j % BAND_AHEAD) == 0 && tid >= nt - BAND_AHEAD) {
                // refill: entries first touched BAND_AHEAD steps from now go into slots retired long ago
                const int idx = j + kl + 1 + BAND_AHEAD + (tid - (nt - BAND_AHEAD));
                if (idx < n) w[idx % W] = b[idx];
            }
            __syncthreads();
        }
        __syncthreads();
    }
    // ---- backward: U x = y, U has at most `ubw` super-diagonals -------------------------------------------------
    {
        const int W = ubw + 1 + 2 * BAND_AHEAD;
        __syncthreads();
        for (int i = tid; i < min(n, ubw + 1 + BAND_AHEAD); i += nt) w[(n - 1 - i) % W] = b[n - 1 - i];
        __syncthreads();
        for (int j = n - 1; j >= 0; --j) {
            const int km = min(ubw, j);
            const double* col = ab + (long long)j * ld + kv;     // col[-i] = U(j - i, j)
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = 1 + tid + k * nt;
                lr[k] = i <= km ? __ldg(col - i) : 0.0;
            }
            const double xj = w[j % W] / __ldg(col);
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = 1 + tid + k * nt;
                if (i <= km) w[(j - i) % W] -= lr[k] * xj;
            }
            if (tid == 0) b[j] = xj;
            const int step = n - 1 - j;
            if ((step % BAND_AHEAD) == 0 && tid >= nt - BAND_AHEAD) {
                const int idx = j - ubw - 1 - BAND_AHEAD - (tid - (nt - BAND_AHEAD));
                if (idx >= 0) w[idx % W] = b[idx];
            }
            __syncthreads();
        }
    }
}

extern "C" {

long long ocmp_band_len(int n, int kl, int ku) { return (long long)n * (2LL * kl + ku + 1); }

int ocmp_band_fill(int nrows, const int* rowptr, const int* colidx, const double* vals, const int* perm, int n,
                   int kl, int ku, double* ab, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    cudaMemsetAsync(ab, 0, sizeof(double) * (size_t)ocmp_band_len(n, kl, ku), st);
    const long long threads = (long long)nrows * 16;
    k_band_fill<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(nrows, rowptr, colidx, vals, perm, kl, ku, ab);
    return ocmp_check("ocmp_band_fill");
}

int ocmp_band_factor(int n, int kl, int ku, double* ab, int* ipiv, int* info_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return 0;
    ProfScope ps(PROF_SETUP, st);
    const size_t smem = sizeof(double) * (size_t)max(kl, 1);
    if (smem > 200 * 1024) return ocmp_fail(-30, "ocmp_band_factor: lower bandwidth too large for shared memory");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_band_lu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_band_lu, 256, smem);
    if (per_sm < 1) return ocmp_fail(-31, "ocmp_band_factor: kernel does not fit");
    // one warp per window column; no more CTAs than the window can feed (a smaller grid has a cheaper barrier)
    const int want = (kl + ku + 1 + 7) / 8 + 1;
    int grid = min(ocmp_sm_count() * min(per_sm, 2), want);
    if (grid < 1) grid = 1;
    cudaMemsetAsync(info_dev, 0, 2 * sizeof(int), st);
    void* args[] = {&n, &kl, &ku, &ab, &ipiv, &info_dev};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_band_lu, dim3(grid), dim3(256), args, smem, st);
    if (e != cudaSuccess) return ocmp_fail(-32, cudaGetErrorString(e));
    return ocmp_check("ocmp_band_factor");
}

int ocmp_band_solve(int n, int kl, int ku, int ubw, const double* ab, const int* ipiv, double* b, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return 0;
    ProfScope ps(PROF_SPMV, st);
    if (ubw < 0 || ubw > kl + ku) ubw = kl + ku;
    const int wide = max(kl, ubw);
    int nt = ((wide + 31) / 32) * 32;
    nt = max(64, min(1024, nt));
    const int pf = (wide + nt - 1) / nt;
    const size_t smem = sizeof(double) * (size_t)(wide + 1 + 2 * BAND_AHEAD);
    if (smem > 220 * 1024 || pf > 16) return ocmp_fail(-33, "ocmp_band_solve: bandwidth too large");
#define LAUNCH(P)                                                                                                  \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            cudaFuncSetAttribute(k_band_solve<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        k_band_solve<P><<<1, nt, smem, st>>>(n, kl, ku, ubw, ab, ipiv, b);                                         \
    } while (0)
    if (pf <= 1) LAUNCH(1);
    else if (pf <= 2) LAUNCH(2);
    else if (pf <= 4) LAUNCH(4);
    else if (pf <= 8) LAUNCH(8);
    else LAUNCH(16);
#undef LAUNCH
    return ocmp_check("ocmp_band_solve");
}

int ocmp_band_gather(int nrows, const int* perm, const double* r, double* b, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nrows <= 0) return 0;
    k_band_gather<<<(nrows + 255) / 256, 256, 0, st>>>(nrows, perm, r, b);
    return ocmp_check("ocmp_band_gather");
}

int ocmp_band_scatter(int nrows, const int* perm, const double* x, double* out, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nrows <= 0) return 0;
    k_band_scatter<<<(nrows + 255) / 256, 256, 0, st>>>(nrows, perm, x, out, accumulate);
    return ocmp_check("ocmp_band_scatter");
}

}  // extern "C"
