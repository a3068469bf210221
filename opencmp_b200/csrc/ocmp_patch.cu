// Batched dense patch kernels for the additive-Schwarz smoother (sm_100a).
//
//   k_patch_invert<T>   one CTA (16 x 16 threads) per patch: the (16T x 16T padded) patch matrix lives entirely in
//                       registers, cyclically distributed (thread (ty,tx) owns rows ty+16a, columns tx+16b); each
//                       Gauss-Jordan step broadcasts the pivot row / column through shared memory and every thread
//                       does T*T FP64 FMAs on its tile. No pivoting: patch dofs are listed in ascending global order,
//                       so velocity dofs precede pressure dofs and the saddle-point blocks are quasi-definite; a
//                       vanishing pivot raises a flag and the host falls back to the pivoted shared-memory kernel.
//                       (A 512-thread variant — row groups split over two halves of the CTA, 16 warps per SM, <= 128
//                       registers — was measured in round 2: 12.9 ms against 9.9 ms for 16 641 patches of 132
//                       dofs; the extra barrier participants cost more than the added latency hiding gains.)
//   k_patch_apply       one CTA per patch: z[dofs] += A_p^-1 r[dofs], columns split over the warps, rows over lanes.
//
// Stand in for the block inversions / block solves of NGSolve's block-Jacobi and multigrid smoothers behind
// ngs.Preconditioner(...).Update() and its application inside the Krylov loop
// (reference opencmp/models/base_model.py:365-383, opencmp/solvers/base_solver.py:711-719).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
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
                        // no pivoting, and no RELATIVE pivot test either: the reference's regularisation (-1e-10 p q,
                        // phi clamped at 1e-10) makes legitimate pressure pivots 1e-20 times smaller than the
                        // velocity ones (a 1e-13 relative threshold was tried in round 2: it sent every patch of the
                        // INS systems to the pivoted fallback). Only an exactly vanishing / non-finite pivot does.
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
// y[p, :] = A_p^-1 r[dofs_p] for every patch — the smoother streams its stored inverses once per application, so the
// kernel is a pure HBM stream and is built like one: persistent CTAs, each walking its patches as ONE sequence of
// chunks (a chunk = CC whole columns of the transposed-stored inverse = one contiguous byte range), copied into a
// ring of NSTAGE shared-memory buffers by the bulk-copy engine (cp.async.bulk + mbarrier transaction counts, issued
// by one thread), so the bytes in flight per SM are set by the ring, not by registers or occupancy. A thread owns
// VEC = 16 / sizeof(T) consecutive rows (one 128-bit shared-memory load per column: 2 FP64, 4 FP32 or 8 bfloat16
// entries) and every VEC-th column of the chunk; the VEC column groups are summed through shared memory in a fixed
// order when the patch ends. One load instruction per 16 bytes instead of one per entry is what keeps the narrow
// storage types on the HBM roofline instead of the issue rate (FP32 with one row per thread: 0.77 of the peak).
// Accumulation is FP64 FMA for every storage type. The right-hand side of the NEXT patch is gathered into a register
// while the current one streams. Results go to the patch-local output y (coalesced stores, no atomics);
// k_patch_gather then sums, for every dof, the entries of its patches in a fixed order — the smoother is
// deterministic.
namespace {
constexpr int NSTAGE = 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
}  // namespace

// One 128-bit shared-memory load = VEC consecutive rows of one column, widened to FP64. The FP32 -> FP64 widening stays
// on F2F (XU pipe, 60 % busy at this rate): doing it, or half of it, with integer instructions (re-biased exponent,
// mantissa spread over two words) was measured slower (0.241 / 0.220 ms against 0.209 ms on the fine level).
template <typename T> struct RowVec;
template <> struct RowVec<double> {
    static constexpr int N = 2;
    static __device__ __forceinline__ void load(const unsigned char* p, double (&a)[2]) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        a[0] = v.x; a[1] = v.y;
    }
};
template <> struct RowVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const unsigned char* p, double (&a)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        a[0] = (double)v.x; a[1] = (double)v.y; a[2] = (double)v.z; a[3] = (double)v.w;
    }
};
template <> struct RowVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const unsigned char* p, double (&a)[8]) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {           // a bfloat16 is the upper half of the FP32 with the same value
            a[2 * k] = (double)__uint_as_float(w[k] << 16);
            a[2 * k + 1] = (double)__uint_as_float(w[k] & 0xffff0000u);
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(256) k_patch_apply_stream(int npatch, int bs, int cc, int nchunk, int stage_bytes,
                                                            const int* __restrict__ pdofs, const T* __restrict__ inv,
                                                            const double* __restrict__ r, double* __restrict__ y) {
    constexpr int VEC = RowVec<T>::N;
    extern __shared__ __align__(128) unsigned char smraw[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw);            // NSTAGE barriers
    double* rl = reinterpret_cast<double*>(smraw + 128);                                // bs doubles
    const int rl_bytes = ((bs * 8 + 127) / 128) * 128;
    double* part = reinterpret_cast<double*>(smraw + 128 + rl_bytes);                   // VEC x bs partial sums
    unsigned char* stages = smraw + 128 + (1 + VEC) * rl_bytes;
    const int tid = threadIdx.x;
    if ((int)blockIdx.x >= npatch) return;
    const int mine = (npatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long total = (long long)mine * nchunk;
    const long long colbytes = (long long)bs * sizeof(T);
    const long long patch_elems = (long long)bs * bs;
    const int nq = bs / VEC;                 // row groups; bs is a multiple of VEC (16-byte columns)
    const int q = tid % nq, grp = tid / nq;  // this thread: rows VEC q .. VEC q + VEC - 1, columns grp, grp + VEC, ...
    const bool worker = grp < VEC;

    auto issue = [&](long long g) {          // thread 0: start the copy of chunk g of this CTA's sequence
        const long long k = g / nchunk;
        const int c = (int)(g - k * nchunk);
        const long long p = blockIdx.x + k * gridDim.x;
        const int cols = (c == nchunk - 1) ? bs - c * cc : cc;
        const unsigned bytes = (unsigned)(cols * colbytes);
        const int s = (int)(g % NSTAGE);
        const unsigned bar = smem_u32(bars + s);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(stages + (size_t)s * stage_bytes), inv + p * patch_elems + (long long)c * cc * bs, bytes, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (long long g = 0; g < NSTAGE && g < total; ++g) issue(g);
    long long p = blockIdx.x;
    if (tid < bs) {
        const int d = __ldg(pdofs + p * bs + tid);
        rl[tid] = d >= 0 ? __ldg(r + d) : 0.0;
    }
    __syncthreads();
    double acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
    double rnext = 0.0;
    int c = 0;
    for (long long g = 0; g < total; ++g) {
        const int s = (int)(g % NSTAGE);
        if (c == 0 && tid < bs && p + gridDim.x < npatch) {
            // right-hand side of the next patch: two dependent loads whose latency the chunks of this patch hide
            const int d = __ldg(pdofs + (p + gridDim.x) * bs + tid);
            rnext = d >= 0 ? __ldg(r + d) : 0.0;
        }
        mbar_wait(smem_u32(bars + s), (unsigned)((g / NSTAGE) & 1));
        const int cols = (c == nchunk - 1) ? bs - c * cc : cc;
        if (worker) {
            const unsigned char* A = stages + (size_t)s * stage_bytes + (size_t)q * 16;
            const double* rr = rl + c * cc;
#pragma unroll 2
            for (int j = grp; j < cols; j += VEC) {
                double a[VEC];
                RowVec<T>::load(A + (size_t)j * colbytes, a);
                const double rj = rr[j];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = fma(a[v], rj, acc[v]);
            }
            if (c == nchunk - 1) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) part[grp * bs + q * VEC + v] = acc[v];
            }
        }
        __syncthreads();                      // stage s (and, on the last chunk, rl) may be overwritten now
        if (tid == 0 && g + NSTAGE < total) issue(g + NSTAGE);
        if (++c == nchunk) {
            if (tid < bs) {
                double sum = part[tid];
#pragma unroll
                for (int v = 1; v < VEC; ++v) sum += part[v * bs + tid];
                y[p * bs + tid] = sum;
                rl[tid] = rnext;
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
            c = 0;
            p += gridDim.x;
            __syncthreads();
        }
    }
}

// z[d] (+)= scale * w[d] * m[d] * sum over the (patch, slot) entries of dof d, in the fixed order of the incidence list
__global__ void __launch_bounds__(256) k_patch_gather(long long n, const int* __restrict__ inc_ptr,
                                                      const int* __restrict__ inc_idx, const double* __restrict__ y,
                                                      const double* __restrict__ w, const double* __restrict__ m,
                                                      double scale, double* __restrict__ z, int accumulate) {
    for (long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x; d < n; d += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        const int a = __ldg(inc_ptr + d), b = __ldg(inc_ptr + d + 1);
        for (int k = a; k < b; ++k) s += __ldg(y + __ldg(inc_idx + k));
        if (w) s *= __ldg(w + d);
        if (m) s *= __ldg(m + d);
        s *= scale;
        z[d] = accumulate ? z[d] + s : s;
    }
}

template <typename T>
static int patch_apply_stream(int npatch, int bs, const int* pd, const T* inv, const double* r, double* y,
                              cudaStream_t st) {
    if (npatch <= 0) return 0;
    if (bs > 256) return ocmp_fail(-20, "patch apply: more than 256 dofs per patch");
    if ((bs * sizeof(T)) % 16) return ocmp_fail(-21, "patch apply: the patch stride must keep columns 16-byte aligned");
    const long long colbytes = (long long)bs * sizeof(T);
    // chunk: whole columns, about 8 KB. Measured on the fine level of the 128 x 128 Taylor-Hood problem (FP32 storage,
    // ms per application): 4 KB 0.319, 6 KB 0.218, 8 KB 0.208, 12 KB 0.226, 16 KB 0.281, 24 KB 0.400 — smaller chunks
    // pay the per-chunk barrier, larger ones leave too few CTAs per SM to cover the FP64 accumulation latency.
    int cc = (int)(8192 / colbytes);
    if (cc < 1) cc = 1;
    if (cc > bs) cc = bs;
    const int nchunk = (bs + cc - 1) / cc;
    const int stage_bytes = (int)(((cc * colbytes) + 127) / 128 * 128);
    const int rl_bytes = ((bs * 8 + 127) / 128) * 128;
    const size_t smem = 128 + (size_t)(1 + RowVec<T>::N) * rl_bytes + (size_t)NSTAGE * stage_bytes;
    static size_t configured[3] = {0, 0, 0};
    const int slot = sizeof(T) == 8 ? 0 : sizeof(T) == 4 ? 1 : 2;
    if (smem > 48 * 1024 && smem > configured[slot]) {
        cudaFuncSetAttribute(k_patch_apply_stream<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[slot] = smem;
    }
    const int threads = ((bs + 31) / 32) * 32;
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    const int cap = ocmp_sm_count() * per_sm;
    const int grid = npatch < cap ? npatch : cap;
    k_patch_apply_stream<T><<<grid, threads, smem, st>>>(npatch, bs, cc, nchunk, stage_bytes, pd, inv, r, y);
    return 0;
}

int ocmp_patch_apply_y(int npatch, int bs, const int* pd, const double* inv, const double* r, double* y,
                       cudaStream_t st) {
    return patch_apply_stream<double>(npatch, bs, pd, inv, r, y, st);
}
int ocmp_patch_apply_y_f32(int npatch, int bs, const int* pd, const float* inv, const double* r, double* y,
                           cudaStream_t st) {
    return patch_apply_stream<float>(npatch, bs, pd, inv, r, y, st);
}
int ocmp_patch_apply_y_bf16(int npatch, int bs, const int* pd, const __nv_bfloat16* inv, const double* r, double* y,
                            cudaStream_t st) {
    return patch_apply_stream<__nv_bfloat16>(npatch, bs, pd, inv, r, y, st);
}
int ocmp_patch_gather(long long n, const int* inc_ptr, const int* inc_idx, const double* y, const double* w,
                      const double* m, double scale, double* z, int accumulate, cudaStream_t st) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)ocmp_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_patch_gather<<<(unsigned)blocks, 256, 0, st>>>(n, inc_ptr, inc_idx, y, w, m, scale, z, accumulate);
    return 0;
}
