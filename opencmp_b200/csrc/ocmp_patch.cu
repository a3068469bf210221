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

template <int T, typename OutT = double>
__global__ void __launch_bounds__(256, 1) k_patch_invert(int npatch, int bs, const int* __restrict__ pdofs,
                                                         const int* __restrict__ rowptr,
                                                         const int* __restrict__ colidx,
                                                         const double* __restrict__ vals,
                                                         const double* __restrict__ fm, OutT* __restrict__ inv,
                                                         int* __restrict__ flag, const int* __restrict__ pos) {
    constexpr int NP = 16 * T;
    __shared__ double rowk[2][NP];
    __shared__ double colk[2][NP];
    __shared__ double pivr[2];            // reciprocal of the pivot, computed once by its owner
    __shared__ int sd[NP];
    __shared__ unsigned char sfree[NP];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        __syncthreads();
        int mine = 0;
        if (threadIdx.x < NP) {                   // NP <= 160 < 256: one entry per thread
            const int i = threadIdx.x;
            const int d = i < bs ? __ldg(pdofs + (long long)p * bs + i) : -1;
            sd[i] = d;
            sfree[i] = (d >= 0 && (!fm || __ldg(fm + d) > 0.0)) ? 1 : 0;
            mine = d >= 0;
        }
        const int nvalid = __syncthreads_count(mine);   // padding (-1) sits at the end of the sorted dof list
        // gather through the cached patch -> nnz positions (built once per pattern by k_patch_positions): T*T
        // independent loads per thread, fully pipelined
        double M[T][T];
        {
            const int* pp = pos + (long long)p * NP * NP;
            int q[T][T];
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) q[a][b] = __ldg(pp + (ty + 16 * a) * NP + tx + 16 * b);
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    const int i = ty + 16 * a, j = tx + 16 * b;
                    double v = (i == j) ? 1.0 : 0.0;
                    if (sfree[i] && sfree[j]) v = q[a][b] >= 0 ? __ldg(vals + q[a][b]) : 0.0;
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
                    if (ty == kr) {
                        const double d = M[ka][ka];
                        if (!(fabs(d) > 1e-280)) bad = true;
                        pivr[buf] = 1.0 / d;
                    }
                }
                __syncthreads();
                const double ip = pivr[buf];
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
        OutT* out = inv + (long long)p * bs * bs;
#pragma unroll
        for (int a = 0; a < T; ++a) {
            const int i = ty + 16 * a;
#pragma unroll
            for (int b = 0; b < T; ++b) {
                const int j = tx + 16 * b;
                if (i < bs && j < bs) out[(long long)j * bs + i] = ocmp_store<OutT>(M[a][b]);
            }
        }
    }
}

template <int T, typename OutT>
static void launch_invert(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                          const double* fm, OutT* inv, int* flag, const int* pos, cudaStream_t st) {
    const int cap = ocmp_sm_count();
    k_patch_invert<T, OutT><<<npatch < cap ? npatch : cap, 256, 0, st>>>(npatch, bs, pd, rp, ci, vals, fm, inv, flag,
                                                                         pos);
}

// returns 1 if handled (flag_dev is set to 1 on a vanishing pivot), 0 if the patch is too large for this kernel
// positions of the (16T x 16T padded) patch entries in the CSR value array, -1 where the pair is not in the pattern
__global__ void __launch_bounds__(256) k_patch_positions(int npatch, int bs, int NP, const int* __restrict__ pdofs,
                                                         const int* __restrict__ rowptr,
                                                         const int* __restrict__ colidx, int* __restrict__ pos) {
    const long long total = (long long)npatch * NP * NP;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / (NP * NP)), rem = (int)(idx % (NP * NP)), i = rem / NP, j = rem % NP;
        int out = -1;
        if (i < bs && j < bs) {
            const int di = __ldg(pdofs + (long long)p * bs + i), dj = __ldg(pdofs + (long long)p * bs + j);
            if (di >= 0 && dj >= 0) {
                int lo = __ldg(rowptr + di), hi = __ldg(rowptr + di + 1) - 1;
                while (lo <= hi) {
                    const int mid = lo + ((hi - lo) >> 1);     // lo + hi overflows int32 once nnz > 2^30
                    const int c = __ldg(colidx + mid);
                    if (c == dj) { out = mid; break; }
                    if (c < dj) lo = mid + 1; else hi = mid - 1;
                }
            }
        }
        pos[idx] = out;
    }
}

extern "C" int ocmp_patch_positions(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                                    int* pos, void* stream) {
    const int NP = 16 * ((bs + 15) / 16);
    const long long total = (long long)npatch * NP * NP;
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ocmp_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    k_patch_positions<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(npatch, bs, NP, patch_dofs, rowptr, colidx, pos);
    return ocmp_check("ocmp_patch_positions");
}

template <typename OutT>
static int invert_registers(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                            const double* fm, OutT* inv, int* flag_dev, const int* pos, cudaStream_t st) {
    const int T = (bs + 15) / 16;
    switch (T) {
        case 1: launch_invert<1, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 2: launch_invert<2, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 3: launch_invert<3, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 4: launch_invert<4, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 5: launch_invert<5, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 6: launch_invert<6, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 7: launch_invert<7, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 8: launch_invert<8, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 9: launch_invert<9, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        case 10: launch_invert<10, OutT>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st); return 1;
        default: return 0;
    }
}

int ocmp_patch_invert_registers(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                                const double* fm, double* inv, int* flag_dev, const int* pos, cudaStream_t st) {
    return invert_registers<double>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st);
}

// the same inversion (FP64 arithmetic in registers), inverse stored in FP32: the smoother only preconditions, and its
// application is bound by streaming the stored inverses from HBM — half the bytes, same GMRES iteration counts
int ocmp_patch_invert_registers_f32(int npatch, int bs, const int* pd, const int* rp, const int* ci,
                                    const double* vals, const double* fm, float* inv, int* flag_dev, const int* pos,
                                    cudaStream_t st) {
    return invert_registers<float>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st);
}

int ocmp_patch_invert_registers_bf16(int npatch, int bs, const int* pd, const int* rp, const int* ci,
                                     const double* vals, const double* fm, __nv_bfloat16* inv, int* flag_dev,
                                     const int* pos, cudaStream_t st) {
    return invert_registers<__nv_bfloat16>(npatch, bs, pd, rp, ci, vals, fm, inv, flag_dev, pos, st);
}

// ---- application ------------------------------------------------------------------------------------------------
// One CTA per patch. Columns of the (transposed-stored) inverse are dealt round-robin to the 8 warps; within a column
// every lane owns the row pairs (2*lane, 2*lane+1) + 64*m and streams them with 16-byte loads, two columns per
// iteration, so a warp keeps up to 2*MR independent 512-byte requests in flight.
template <int MR, int NW>
__global__ void __launch_bounds__(NW * 32) k_patch_apply(int npatch, int bs, const int* __restrict__ pdofs,
                                                     const double* __restrict__ inv, const double* __restrict__ r,
                                                     double* __restrict__ z) {
    extern __shared__ double sm[];         // r_loc[bs], partial[NW][bs]
    double* rl = sm;
    double* part = sm + bs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool even = (bs & 1) == 0;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int* d = pdofs + (long long)p * bs;
        __syncthreads();
        for (int j = threadIdx.x; j < bs; j += NW * 32) {
            const int dj = __ldg(d + j);
            rl[j] = dj >= 0 ? __ldg(r + dj) : 0.0;
        }
        __syncthreads();
        const double* A = inv + (long long)p * bs * bs;
        double s0[MR], s1[MR];
#pragma unroll
        for (int m = 0; m < MR; ++m) { s0[m] = 0.0; s1[m] = 0.0; }
        if (even) {
            int j = warp;
            for (; j + NW < bs; j += 2 * NW) {
                const double ra = rl[j], rb = rl[j + NW];
                const double2* ca = reinterpret_cast<const double2*>(A + (long long)j * bs);
                const double2* cb = reinterpret_cast<const double2*>(A + (long long)(j + NW) * bs);
                double2 va[MR], vb[MR];
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i2 = lane + 32 * m;
                    const bool ok = 2 * i2 < bs;
                    va[m] = ok ? __ldg(ca + i2) : make_double2(0.0, 0.0);
                    vb[m] = ok ? __ldg(cb + i2) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    s0[m] = fma(va[m].x, ra, s0[m]); s1[m] = fma(va[m].y, ra, s1[m]);
                    s0[m] = fma(vb[m].x, rb, s0[m]); s1[m] = fma(vb[m].y, rb, s1[m]);
                }
            }
            for (; j < bs; j += NW) {
                const double ra = rl[j];
                const double2* ca = reinterpret_cast<const double2*>(A + (long long)j * bs);
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i2 = lane + 32 * m;
                    if (2 * i2 < bs) {
                        const double2 v = __ldg(ca + i2);
                        s0[m] = fma(v.x, ra, s0[m]); s1[m] = fma(v.y, ra, s1[m]);
                    }
                }
            }
        } else {
            for (int j = warp; j < bs; j += NW) {
                const double ra = rl[j];
                const double* ca = A + (long long)j * bs;
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i = 2 * (lane + 32 * m);
                    if (i < bs) s0[m] = fma(__ldg(ca + i), ra, s0[m]);
                    if (i + 1 < bs) s1[m] = fma(__ldg(ca + i + 1), ra, s1[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            const int i = 2 * (lane + 32 * m);
            if (i < bs) part[warp * bs + i] = s0[m];
            if (i + 1 < bs) part[warp * bs + i + 1] = s1[m];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < bs; i += NW * 32) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += part[w * bs + i];
            const int di = __ldg(d + i);
            if (di >= 0) atomicAdd(z + di, t);
        }
    }
}

int ocmp_patch_apply_cta(int npatch, int bs, const int* pd, const double* inv, const double* r, double* z,
                         cudaStream_t st) {
    if (bs > 256) return 0;
    constexpr int NW = 4;
    const size_t smem = sizeof(double) * (NW + 1) * bs;
    const int cap = ocmp_sm_count() * 16;
    const int grid = npatch < cap ? npatch : cap;
    if (bs <= 64) k_patch_apply<1, NW><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    else if (bs <= 128) k_patch_apply<2, NW><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    else if (bs <= 192) k_patch_apply<3, NW><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    else k_patch_apply<4, NW><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    return 1;
}

// FP32-stored inverses (bs a multiple of 4 so that every column starts 16-byte aligned): the same column-per-warp
// scheme, every lane owns the rows 4*lane .. 4*lane+3 (+ 128*m) and streams them with 16-byte loads, UC columns per
// iteration in flight; products and sums in FP64.
template <int MR, int NW, int UC>
__global__ void __launch_bounds__(NW * 32) k_patch_apply_f32(int npatch, int bs, const int* __restrict__ pdofs,
                                                             const float* __restrict__ inv,
                                                             const double* __restrict__ r, double* __restrict__ z) {
    extern __shared__ double sm[];         // r_loc[bs], partial[NW][bs]
    double* rl = sm;
    double* part = sm + bs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int* d = pdofs + (long long)p * bs;
        __syncthreads();
        for (int j = threadIdx.x; j < bs; j += NW * 32) {
            const int dj = __ldg(d + j);
            rl[j] = dj >= 0 ? __ldg(r + dj) : 0.0;
        }
        __syncthreads();
        const float* A = inv + (long long)p * bs * bs;
        double s[MR][4];
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[m][c] = 0.0;
        int j = warp;
        for (; j + (UC - 1) * NW < bs; j += UC * NW) {
            float4 v[UC][MR];
            double rj[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) {
                rj[u] = rl[j + u * NW];
                const float4* col = reinterpret_cast<const float4*>(A + (long long)(j + u * NW) * bs);
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i4 = lane + 32 * m;
                    v[u][m] = (4 * i4 < bs) ? __ldg(col + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UC; ++u)
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    s[m][0] = fma((double)v[u][m].x, rj[u], s[m][0]);
                    s[m][1] = fma((double)v[u][m].y, rj[u], s[m][1]);
                    s[m][2] = fma((double)v[u][m].z, rj[u], s[m][2]);
                    s[m][3] = fma((double)v[u][m].w, rj[u], s[m][3]);
                }
        }
        for (; j < bs; j += NW) {
            const double rj = rl[j];
            const float4* col = reinterpret_cast<const float4*>(A + (long long)j * bs);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const int i4 = lane + 32 * m;
                if (4 * i4 < bs) {
                    const float4 v = __ldg(col + i4);
                    s[m][0] = fma((double)v.x, rj, s[m][0]);
                    s[m][1] = fma((double)v.y, rj, s[m][1]);
                    s[m][2] = fma((double)v.z, rj, s[m][2]);
                    s[m][3] = fma((double)v.w, rj, s[m][3]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            const int i = 4 * (lane + 32 * m);
            if (i < bs) {                      // bs % 4 == 0: all four rows exist
#pragma unroll
                for (int c = 0; c < 4; ++c) part[warp * bs + i + c] = s[m][c];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < bs; i += NW * 32) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += part[w * bs + i];
            const int di = __ldg(d + i);
            if (di >= 0) atomicAdd(z + di, t);
        }
    }
}

int ocmp_patch_apply_cta_f32(int npatch, int bs, const int* pd, const float* inv, const double* r, double* z,
                             cudaStream_t st) {
    if (bs > 256 || (bs & 3)) return 0;
    constexpr int NW = 4;
    const size_t smem = sizeof(double) * (NW + 1) * bs;
    const int cap = ocmp_sm_count() * 16;
    const int grid = npatch < cap ? npatch : cap;
    if (bs <= 128) k_patch_apply_f32<1, NW, 4><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    else k_patch_apply_f32<2, NW, 2><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    return 1;
}

// bfloat16-stored inverses (bs a multiple of 8): every HALF warp streams one column, a lane owns the rows
// 8*(lane % 16) .. +7 (+ 128*m) and loads them as one 16-byte word; UC columns per half warp in flight. A bf16 is the
// upper half of an FP32, so the conversion is a shift; products and sums in FP64.
__device__ __forceinline__ void bf16x2_fma(unsigned w, double rj, double& s0, double& s1) {
    s0 = fma((double)__uint_as_float(w << 16), rj, s0);
    s1 = fma((double)__uint_as_float(w & 0xffff0000u), rj, s1);
}

template <int MR, int NW, int UC>
__global__ void __launch_bounds__(NW * 32) k_patch_apply_bf16(int npatch, int bs, const int* __restrict__ pdofs,
                                                              const __nv_bfloat16* __restrict__ inv,
                                                              const double* __restrict__ r, double* __restrict__ z) {
    extern __shared__ double sm[];         // r_loc[bs], partial[2 * NW][bs]
    constexpr int G = 2 * NW;              // column groups of the CTA
    double* rl = sm;
    double* part = sm + bs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l16 = lane & 15;
    const int g = 2 * warp + (lane >> 4);
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const int* d = pdofs + (long long)p * bs;
        __syncthreads();
        for (int j = threadIdx.x; j < bs; j += NW * 32) {
            const int dj = __ldg(d + j);
            rl[j] = dj >= 0 ? __ldg(r + dj) : 0.0;
        }
        __syncthreads();
        const __nv_bfloat16* A = inv + (long long)p * bs * bs;
        double s[MR][8];
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int c = 0; c < 8; ++c) s[m][c] = 0.0;
        int j = g;
        for (; j + (UC - 1) * G < bs; j += UC * G) {
            uint4 v[UC][MR];
            double rj[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) {
                rj[u] = rl[j + u * G];
                const uint4* col = reinterpret_cast<const uint4*>(A + (long long)(j + u * G) * bs);
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    const int i8 = l16 + 16 * m;
                    v[u][m] = (8 * i8 < bs) ? __ldg(col + i8) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int u = 0; u < UC; ++u)
#pragma unroll
                for (int m = 0; m < MR; ++m) {
                    bf16x2_fma(v[u][m].x, rj[u], s[m][0], s[m][1]);
                    bf16x2_fma(v[u][m].y, rj[u], s[m][2], s[m][3]);
                    bf16x2_fma(v[u][m].z, rj[u], s[m][4], s[m][5]);
                    bf16x2_fma(v[u][m].w, rj[u], s[m][6], s[m][7]);
                }
        }
        for (; j < bs; j += G) {
            const double rj = rl[j];
            const uint4* col = reinterpret_cast<const uint4*>(A + (long long)j * bs);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const int i8 = l16 + 16 * m;
                if (8 * i8 < bs) {
                    const uint4 v = __ldg(col + i8);
                    bf16x2_fma(v.x, rj, s[m][0], s[m][1]);
                    bf16x2_fma(v.y, rj, s[m][2], s[m][3]);
                    bf16x2_fma(v.z, rj, s[m][4], s[m][5]);
                    bf16x2_fma(v.w, rj, s[m][6], s[m][7]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            const int i = 8 * (l16 + 16 * m);
            if (i < bs) {                      // bs % 8 == 0: all eight rows exist
#pragma unroll
                for (int c = 0; c < 8; ++c) part[g * bs + i + c] = s[m][c];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < bs; i += NW * 32) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < G; ++w) t += part[w * bs + i];
            const int di = __ldg(d + i);
            if (di >= 0) atomicAdd(z + di, t);
        }
    }
}

int ocmp_patch_apply_cta_bf16(int npatch, int bs, const int* pd, const __nv_bfloat16* inv, const double* r, double* z,
                              cudaStream_t st) {
    if (bs > 256 || (bs & 7)) return 0;
    constexpr int NW = 4;
    const size_t smem = sizeof(double) * (2 * NW + 1) * bs;
    const int cap = ocmp_sm_count() * 16;
    const int grid = npatch < cap ? npatch : cap;
    if (bs <= 128) k_patch_apply_bf16<1, NW, 4><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    else k_patch_apply_bf16<2, NW, 4><<<grid, NW * 32, smem, st>>>(npatch, bs, pd, inv, r, z);
    return 1;
}
