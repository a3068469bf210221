// Sparse / dense vector kernels, preconditioners and Krylov drivers of the opencmp_b200 backend (sm_100a).
//
// Stand in for NGSolve's SparseMatrix<double>::Mult, BaseVector arithmetic, Preconditioner(a,'local'|...) and the
// pure-Python ngsolve.solvers CG / GMRes / PreconditionedRichardson that Model.linear_solve dispatches to
// (reference opencmp/models/base_model.py:886-947).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <unordered_map>
#include <vector>
#include "../../include/opencmp_b200.h"
#include "ocmp_common.cuh"

// ---- error plumbing ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int ocmp_fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int ocmp_check(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
        return -100;
    }
    return 0;
}
int ocmp_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
// ---- profiling / launch counting ---------------------------------------------------------------------------------
namespace {
struct ProfCat {
    std::vector<cudaEvent_t> ev;     // begin/end pairs not yet folded into ms
    double ms = 0.0;
    long long count = 0;
    double bytes = 0.0;              // algorithmic bytes moved by the launches of this category
};
ProfCat g_cat[PROF_NCAT];
std::vector<cudaEvent_t> g_pool;
bool g_prof = false;
long long g_launches = 0;
cudaEvent_t prof_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_fold(ProfCat& c) {
    for (size_t i = 0; i + 1 < c.ev.size(); i += 2) {
        float ms = 0.f;
        cudaEventSynchronize(c.ev[i + 1]);
        cudaEventElapsedTime(&ms, c.ev[i], c.ev[i + 1]);
        c.ms += ms;
        g_pool.push_back(c.ev[i]);
        g_pool.push_back(c.ev[i + 1]);
    }
    c.ev.clear();
}
}  // namespace
void ocmp_prof_begin(int cat, cudaStream_t st) {
    ++g_launches;
    if (!g_prof) return;
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_cat[cat].ev.push_back(e);
}
void ocmp_prof_end(int cat, cudaStream_t st) {
    if (!g_prof) return;
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_cat[cat].ev.push_back(e);
    g_cat[cat].count++;
    if (g_cat[cat].ev.size() > 4096) prof_fold(g_cat[cat]);
}
extern "C" void ocmp_profile_enable(int on) { g_prof = on != 0; }
extern "C" void ocmp_profile_reset(void) {
    for (auto& c : g_cat) { prof_fold(c); c.ms = 0.0; c.count = 0; c.bytes = 0.0; }
    g_launches = 0;
}
extern "C" int ocmp_profile_read(int cat, long long* count, double* ms) {
    if (cat < 0 || cat >= PROF_NCAT) return -1;
    prof_fold(g_cat[cat]);
    if (count) *count = g_cat[cat].count;
    if (ms) *ms = g_cat[cat].ms;
    return 0;
}
void ocmp_prof_bytes(int cat, double bytes) {
    if (g_prof) g_cat[cat].bytes += bytes;
}
extern "C" double ocmp_profile_bytes(int cat) { return (cat < 0 || cat >= PROF_NCAT) ? 0.0 : g_cat[cat].bytes; }
extern "C" long long ocmp_launch_count(void) { return g_launches; }
extern "C" const char* ocmp_last_error(void) { return g_err; }
extern "C" int ocmp_version(void) { return 100; }

// ---- SpMV: CSR, LPR lanes cooperate on one row; the epilogue is fused into the row store ---------------------------
//   EP_PLAIN  y = A x                     EP_RESID  y = m .* m2 .* (b - A x)   (residual, restriction input)
//   EP_MASK   y = m .* (A x)              EP_ADD    y += m .* (A x)            (prolongation + correction)
// VT = double | float: the matrix values may be an FP32 copy (operator applications INSIDE the multigrid cycle when
// ocmp_system.vals32 is set — 8 instead of 12 bytes per non-zero); products and sums are FP64.
enum { EP_PLAIN = 0, EP_RESID = 1, EP_MASK = 2, EP_ADD = 3 };
// Column-index compression for vector-valued spaces ("component runs"): the components of a VectorH1 block are
// numbered one after the other (u_x dofs, then u_y, then u_z, each `shift` apart), and the pattern keeps every
// component pair, so a row's columns start with nc runs of equal length L holding the same scalar columns shifted by
// 0, shift, 2 shift. Such rows (runlen[row] = L > 0, verified once by k_spmv_runs) read only the FIRST run's indices:
// 4 index bytes per nc values instead of 4 per value. The rest of the row (pressure columns) is read as usual.
struct SpmvRuns {
    const int* runlen;       // per row: L, or 0 for a plain row; NULL = no compression
    int shift, nc;
    int grouped;             // > 0: the nc component rows of every node have identical column lists (k_spmv_vec);
                             // with a row list: the number of listed rows below `shift`
};

template <int LPR, typename VT, int EP>
__global__ void __launch_bounds__(256) k_spmv(int nrows, const int* __restrict__ rowptr, const int* __restrict__ col,
                                              const VT* __restrict__ val, const double* __restrict__ x,
                                              double* __restrict__ y, const double* __restrict__ b,
                                              const double* __restrict__ m, const double* __restrict__ m2,
                                              const int* __restrict__ rows, SpmvRuns runs) {
    const int lane = threadIdx.x % LPR;
    const long long row0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const long long stride = (long long)gridDim.x * blockDim.x / LPR;
    for (long long rr = row0; rr < nrows; rr += stride) {
        const long long row = rows ? __ldg(rows + rr) : rr;      // `nrows` counts the listed rows when a list is given
        int a = __ldg(rowptr + row);
        const int e = __ldg(rowptr + row + 1);
        double s = 0.0;
        const int L = runs.runlen ? __ldg(runs.runlen + row) : 0;
        if (L > 0) {
            const VT* v0 = val + a;
            const VT* v1 = v0 + L;
            const VT* v2 = v1 + L;
            if (runs.nc == 3) {
                for (int k = lane; k < L; k += LPR) {
                    const double* xp = x + __ldg(col + a + k);
                    s = fma((double)__ldg(v0 + k), __ldg(xp), s);
                    s = fma((double)__ldg(v1 + k), __ldg(xp + runs.shift), s);
                    s = fma((double)__ldg(v2 + k), __ldg(xp + 2 * (long long)runs.shift), s);
                }
            } else {
                for (int k = lane; k < L; k += LPR) {
                    const double* xp = x + __ldg(col + a + k);
                    s = fma((double)__ldg(v0 + k), __ldg(xp), s);
                    s = fma((double)__ldg(v1 + k), __ldg(xp + runs.shift), s);
                }
            }
            a += runs.nc * L;
        }
        for (int k = a + lane; k < e; k += LPR) s = fma((double)__ldg(val + k), __ldg(x + __ldg(col + k)), s);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
        if (lane == 0) {
            if (EP == EP_PLAIN) y[row] = s;
            else {
                double v = (EP == EP_RESID) ? __ldg(b + row) - s : s;
                if (m) v *= __ldg(m + row);
                if (EP == EP_RESID && m2) v *= __ldg(m2 + row);
                y[row] = (EP == EP_ADD) ? y[row] + v : v;
            }
        }
    }
}

// The nc component rows of a node (row, row + shift, ...) have the SAME column list, so one lane group computes all of
// them: the x values and the column indices are fetched once per nc x nc values — the gather traffic through L2
// (one 32-byte sector per 8-byte x value for a scattered CG numbering) drops by nc, which is what bounds the plain
// CSR product on the 3-D Taylor-Hood matrices (210 non-zeros per row). Groups [0, ngroups) are nodes, the remaining
// `nsingle` rows (pressure) are handled like in k_spmv.
template <int LPR, typename VT, int EP, int NC>
__global__ void __launch_bounds__(256) k_spmv_vec(int ngroups, int nsingle, const int* __restrict__ rowptr,
                                                  const int* __restrict__ col, const VT* __restrict__ val,
                                                  const double* __restrict__ x, double* __restrict__ y,
                                                  const double* __restrict__ b, const double* __restrict__ m,
                                                  const double* __restrict__ m2, const int* __restrict__ rows,
                                                  SpmvRuns runs) {
    const int lane = threadIdx.x % LPR;
    const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const long long stride = (long long)gridDim.x * blockDim.x / LPR;
    const long long shift = runs.shift;
    for (long long g = g0; g < (long long)ngroups + nsingle; g += stride) {
        if (g < ngroups) {
            const long long row0 = rows ? __ldg(rows + g) : g;
            const int a0 = __ldg(rowptr + row0), len = __ldg(rowptr + row0 + 1) - a0;
            const int L = __ldg(runs.runlen + row0);
            const VT* v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = val + __ldg(rowptr + row0 + c * shift);
            double s[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) s[c] = 0.0;
            for (int k = lane; k < L; k += LPR) {
                const double* xp = x + __ldg(col + a0 + k);
                double xv[NC];
#pragma unroll
                for (int c2 = 0; c2 < NC; ++c2) xv[c2] = __ldg(xp + c2 * shift);
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int c2 = 0; c2 < NC; ++c2) s[c] = fma((double)__ldg(v[c] + c2 * L + k), xv[c2], s[c]);
            }
            for (int k = NC * L + lane; k < len; k += LPR) {
                const double xv = __ldg(x + __ldg(col + a0 + k));
#pragma unroll
                for (int c = 0; c < NC; ++c) s[c] = fma((double)__ldg(v[c] + k), xv, s[c]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) s[c] += __shfl_down_sync(0xffffffffu, s[c], o, LPR);
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const long long row = row0 + c * shift;
                    if (EP == EP_PLAIN) y[row] = s[c];
                    else {
                        double vv = (EP == EP_RESID) ? __ldg(b + row) - s[c] : s[c];
                        if (m) vv *= __ldg(m + row);
                        if (EP == EP_RESID && m2) vv *= __ldg(m2 + row);
                        y[row] = (EP == EP_ADD) ? y[row] + vv : vv;
                    }
                }
            }
        } else {
            const long long si = g - ngroups;
            const long long row = rows ? __ldg(rows + (long long)NC * ngroups + si) : NC * shift + si;
            int a = __ldg(rowptr + row);
            const int e = __ldg(rowptr + row + 1);
            double s = 0.0;
            const int L = __ldg(runs.runlen + row);
            if (L > 0) {
                for (int k = lane; k < L; k += LPR) {
                    const double* xp = x + __ldg(col + a + k);
#pragma unroll
                    for (int c2 = 0; c2 < NC; ++c2) s = fma((double)__ldg(val + a + c2 * L + k), __ldg(xp + c2 * shift), s);
                }
                a += NC * L;
            }
            for (int k = a + lane; k < e; k += LPR) s = fma((double)__ldg(val + k), __ldg(x + __ldg(col + k)), s);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPR);
            if (lane == 0) {
                if (EP == EP_PLAIN) y[row] = s;
                else {
                    double vv = (EP == EP_RESID) ? __ldg(b + row) - s : s;
                    if (m) vv *= __ldg(m + row);
                    if (EP == EP_RESID && m2) vv *= __ldg(m2 + row);
                    y[row] = (EP == EP_ADD) ? y[row] + vv : vv;
                }
            }
        }
    }
}

// group check of k_spmv_vec: rows row + c * shift (row < shift, c < nc) have the length, run length and columns of
// row; *bad is raised otherwise
__global__ void __launch_bounds__(256) k_spmv_groups(int shift, int nc, const int* __restrict__ rowptr,
                                                     const int* __restrict__ col, const int* __restrict__ runlen,
                                                     int* __restrict__ bad) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= shift) return;
    const int a = rowptr[row], len = rowptr[row + 1] - a;
    bool ok = runlen[row] > 0;
    for (int c = 1; ok && c < nc; ++c) {
        const int r2 = row + c * shift, a2 = rowptr[r2];
        ok = ok && rowptr[r2 + 1] - a2 == len && runlen[r2] == runlen[row];
        for (int k = 0; ok && k < len; ++k) ok = ok && col[a2 + k] == col[a + k];
    }
    if (!ok) *bad = 1;
}

// runlen[row] = L if the row starts with nc runs of length L whose columns are those of the first run shifted by
// k * shift (k < nc) and the first run lies in [0, shift); else 0
__global__ void __launch_bounds__(256) k_spmv_runs(int nrows, const int* __restrict__ rowptr,
                                                   const int* __restrict__ col, int shift, int nc,
                                                   int* __restrict__ runlen) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int a = rowptr[row], e = rowptr[row + 1];
    int L = 0;
    while (a + L < e && col[a + L] < shift) ++L;
    bool ok = L > 0 && a + (long long)nc * L <= e;
    for (int k = 0; ok && k < L; ++k)
        for (int c = 1; c < nc; ++c) ok = ok && col[a + c * L + k] == col[a + k] + c * shift;
    runlen[row] = ok ? L : 0;
}

__global__ void __launch_bounds__(256) k_to_f32(long long n, const double* __restrict__ src, float* __restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

// Lanes per row from the mean row length (16 for the 100+ non-zeros per row of the high-order / DG systems, 8 or 4
// for low-order H1 matrices and the multigrid transfer operators: with 16 lanes a 12-entry row of Poisson P2 left
// half the lanes idle and the product ran at 1.4 TB/s). The row count of a pattern is read back once and remembered.
static int spmv_nnz(const int* rowptr, int nrows_total) {
    static std::unordered_map<const int*, int> cache;
    auto it = cache.find(rowptr);
    if (it == cache.end()) {
        int nnz = 0;
        if (cudaMemcpy(&nnz, rowptr + nrows_total, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) nnz = 0;
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(rowptr, nnz).first;
    }
    return it->second;
}
static int spmv_lanes(const int* rowptr, int nrows_total) {
    const double mean = nrows_total > 0 ? (double)spmv_nnz(rowptr, nrows_total) / nrows_total : 0.0;
    return mean >= 40.0 ? 16 : mean >= 14.0 ? 8 : 4;
}

// one launch: category `cat`, values from `vals32` when given, else `vals`
static int spmv_ep(int cat, int ep, int nrows, const int* rowptr, const int* colidx, const double* vals,
                   const float* vals32, const double* x, double* y, const double* b, const double* m,
                   const double* m2, cudaStream_t st, const int* rows = nullptr,
                   SpmvRuns runs = SpmvRuns{nullptr, 0, 0, 0}, int nrows_total = -1) {
    if (nrows <= 0) return 0;
    const int threads = 256;
    // a row list (owned rows) does not say where rowptr ends: the callers that pass one also pass the full row count
    const int lanes = (runs.runlen && runs.grouped > 0) ? 16 : spmv_lanes(rowptr, nrows_total >= 0 ? nrows_total : nrows);
    const long long want = ((long long)nrows * lanes + threads - 1) / threads;
    const long long cap = (long long)ocmp_sm_count() * 64;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    {   // algorithmic bytes of this product by the CSR count of SURVEY 8(d): nnz (value + 4) + rows (4 + 8) + cols 8
        const int ntot = nrows_total >= 0 ? nrows_total : nrows;
        const double nnz = (double)spmv_nnz(rowptr, ntot) * (ntot > 0 ? (double)nrows / ntot : 0.0);
        ocmp_prof_bytes(cat, nnz * ((vals32 ? 4.0 : 8.0) + 4.0) + 20.0 * nrows);
    }
    ProfScope ps(cat, st);
    if (runs.runlen && runs.grouped > 0) {
        // node-grouped product: listed rows (or all rows) = nc * ngroups component rows followed by single rows
        const int ngroups = runs.grouped, nsingle = nrows - runs.nc * ngroups;
        const long long wantg = ((long long)(ngroups + nsingle) * 16 + threads - 1) / threads;
        const unsigned bg = (unsigned)(wantg < cap ? wantg : cap);
#define SPMV_VEC(VT, V, EP, NC) \
    k_spmv_vec<16, VT, EP, NC><<<bg, threads, 0, st>>>(ngroups, nsingle, rowptr, colidx, V, x, y, b, m, m2, rows, runs)
#define SPMV_VEC_EP(VT, V, NC)                                           \
    do {                                                                 \
        if (ep == EP_PLAIN) SPMV_VEC(VT, V, EP_PLAIN, NC);               \
        else if (ep == EP_RESID) SPMV_VEC(VT, V, EP_RESID, NC);          \
        else if (ep == EP_MASK) SPMV_VEC(VT, V, EP_MASK, NC);            \
        else SPMV_VEC(VT, V, EP_ADD, NC);                                \
    } while (0)
        if (vals32) { if (runs.nc == 3) SPMV_VEC_EP(float, vals32, 3); else SPMV_VEC_EP(float, vals32, 2); }
        else { if (runs.nc == 3) SPMV_VEC_EP(double, vals, 3); else SPMV_VEC_EP(double, vals, 2); }
#undef SPMV_VEC_EP
#undef SPMV_VEC
        return ocmp_check("ocmp_spmv (grouped)");
    }
#define SPMV_GO(VT, V, EP)                                                                                        \
    do {                                                                                                          \
        if (lanes == 16)                                                                                          \
            k_spmv<16, VT, EP><<<blocks, threads, 0, st>>>(nrows, rowptr, colidx, V, x, y, b, m, m2, rows, runs); \
        else if (lanes == 8)                                                                                      \
            k_spmv<8, VT, EP><<<blocks, threads, 0, st>>>(nrows, rowptr, colidx, V, x, y, b, m, m2, rows, runs);  \
        else                                                                                                      \
            k_spmv<4, VT, EP><<<blocks, threads, 0, st>>>(nrows, rowptr, colidx, V, x, y, b, m, m2, rows, runs);  \
    } while (0)
    if (vals32) {
        if (ep == EP_PLAIN) SPMV_GO(float, vals32, EP_PLAIN);
        else if (ep == EP_RESID) SPMV_GO(float, vals32, EP_RESID);
        else if (ep == EP_MASK) SPMV_GO(float, vals32, EP_MASK);
        else SPMV_GO(float, vals32, EP_ADD);
    } else {
        if (ep == EP_PLAIN) SPMV_GO(double, vals, EP_PLAIN);
        else if (ep == EP_RESID) SPMV_GO(double, vals, EP_RESID);
        else if (ep == EP_MASK) SPMV_GO(double, vals, EP_MASK);
        else SPMV_GO(double, vals, EP_ADD);
    }
#undef SPMV_GO
    return ocmp_check("ocmp_spmv");
}

extern "C" int ocmp_spmv(int nrows, const int* rowptr, const int* colidx, const double* vals, const double* x,
                         double* y, void* stream) {
    return spmv_ep(PROF_SPMV, EP_PLAIN, nrows, rowptr, colidx, vals, nullptr, x, y, nullptr, nullptr, nullptr,
                   (cudaStream_t)stream);
}
extern "C" int ocmp_spmv_runs(int nrows, const int* rowptr, const int* colidx, int shift, int nc, int* runlen,
                              int* grouped_bad_dev, void* stream) {
    if (nrows <= 0) return 0;
    if (nc < 2 || nc > 3 || shift <= 0 || (long long)nc * shift > nrows)
        return ocmp_fail(-1, "ocmp_spmv_runs: nc must be 2 or 3, 0 < nc * shift <= nrows");
    cudaStream_t st = (cudaStream_t)stream;
    k_spmv_runs<<<(nrows + 255) / 256, 256, 0, st>>>(nrows, rowptr, colidx, shift, nc, runlen);
    if (grouped_bad_dev) {
        cudaMemsetAsync(grouped_bad_dev, 0, sizeof(int), st);
        k_spmv_groups<<<(shift + 255) / 256, 256, 0, st>>>(shift, nc, rowptr, colidx, runlen, grouped_bad_dev);
    }
    return ocmp_check("ocmp_spmv_runs");
}
extern "C" int ocmp_spmv_compressed(int nrows, const int* rowptr, const int* colidx, const double* vals,
                                    const int* runlen, int shift, int nc, int grouped, const double* x, double* y,
                                    void* stream) {
    return spmv_ep(PROF_SPMV, EP_PLAIN, nrows, rowptr, colidx, vals, nullptr, x, y, nullptr, nullptr, nullptr,
                   (cudaStream_t)stream, nullptr, SpmvRuns{runlen, shift, nc, grouped ? shift : 0});
}

extern "C" int ocmp_to_f32(long long n, const double* src, float* dst, void* stream) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)ocmp_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PROF_SETUP, (cudaStream_t)stream);
    k_to_f32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, src, dst);
    return ocmp_check("ocmp_to_f32");
}

// ---- level-1 ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dot(long long n, const double* __restrict__ x, const double* __restrict__ y,
                                             double* __restrict__ out, const double* __restrict__ m = nullptr) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fma(m ? x[i] * m[i] : x[i], y[i], s);
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256) k_axpby(long long n, double a, const double* __restrict__ x, double b,
                                               double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}

__global__ void __launch_bounds__(256) k_masked_assign(long long n, double* __restrict__ dst,
                                                       const double* __restrict__ src, const double* __restrict__ inv,
                                                       const double* __restrict__ mask) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (mask[i] > 0.0) dst[i] = src[i] * inv[i];
}

// z (+)= scale * a .* m .* r   (a, m optional)
__global__ void __launch_bounds__(256) k_had(long long n, const double* __restrict__ a, const double* __restrict__ m,
                                             const double* __restrict__ r, double* __restrict__ z, double scale,
                                             int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double v = r[i] * scale;
        if (a) v *= a[i];
        if (m) v *= m[i];
        z[i] = accumulate ? z[i] + v : v;
    }
}

// r = m .* (b - r)
__global__ void __launch_bounds__(256) k_resid(long long n, const double* __restrict__ b, const double* __restrict__ m,
                                               double* __restrict__ r) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double v = b[i] - r[i];
        if (m) v *= m[i];
        r[i] = v;
    }
}

// out[j] += <V_j, w>, j < k <= K ; V_j = V + j*ld ; and, when `extra` is given, out[k] += <extra, w> (the squared
// norm of w rides along with the second Gram-Schmidt pass)
template <int K>
__global__ void __launch_bounds__(256) k_mdot(long long n, const double* __restrict__ V, long long ld, int k,
                                              const double* __restrict__ w, double* __restrict__ out,
                                              const double* __restrict__ m, const double* __restrict__ extra) {
    double s[K + 1];
#pragma unroll
    for (int j = 0; j <= K; ++j) s[j] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double wi = m ? w[i] * m[i] : w[i];
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j < k) s[j] = fma(V[j * ld + i], wi, s[j]);
        if (extra) s[K] = fma(extra[i], wi, s[K]);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            const double t = block_reduce_sum(s[j]);
            if (threadIdx.x == 0) atomicAdd(out + j, t);
        }
    }
    if (extra) {
        const double t = block_reduce_sum(s[K]);
        if (threadIdx.x == 0) atomicAdd(out + k, t);
    }
}

// w += sign * sum_j c[j] V_j  (c on device)
__global__ void __launch_bounds__(256) k_maxpy(long long n, const double* __restrict__ V, long long ld, int k,
                                               const double* __restrict__ c, double* __restrict__ w, double sign) {
    extern __shared__ double sc[];
    for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = sign * c[j];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double s = w[i];
        for (int j = 0; j < k; ++j) s = fma(sc[j], V[j * ld + i], s);
        w[i] = s;
    }
}

// Second Gram-Schmidt correction and normalisation in one pass: out = (w - sum_j c[j] V_j) / hn with
// hn^2 = c[k] - sum_j c[j]^2  (c[k] = <w, w> before the correction; V orthonormal). out = 0 when hn^2 <= 0.
__global__ void __launch_bounds__(256) k_gs_finish(long long n, const double* __restrict__ V, long long ld, int k,
                                                   const double* __restrict__ c, const double* __restrict__ w,
                                                   double* __restrict__ out) {
    extern __shared__ double sc[];
    __shared__ double s_inv;
    for (int j = threadIdx.x; j <= k; j += blockDim.x) sc[j] = c[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        double h2 = sc[k];
        for (int j = 0; j < k; ++j) h2 -= sc[j] * sc[j];
        s_inv = h2 > 0.0 ? 1.0 / sqrt(h2) : 0.0;
    }
    __syncthreads();
    const double inv = s_inv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double s = w[i];
        for (int j = 0; j < k; ++j) s = fma(-sc[j], V[j * ld + i], s);
        out[i] = s * inv;
    }
}

static inline unsigned grid_for(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)ocmp_sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" int ocmp_dot(long long n, const double* x, const double* y, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(double), st);
    ProfScope ps(PROF_VEC, st);
    if (n > 0) k_dot<<<grid_for(n), 256, 0, st>>>(n, x, y, out);
    return ocmp_check("ocmp_dot");
}
extern "C" int ocmp_axpby(long long n, double a, const double* x, double b, double* y, void* stream) {
    ProfScope ps(PROF_VEC, (cudaStream_t)stream);
    if (n > 0) k_axpby<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, a, x, b, y);
    return ocmp_check("ocmp_axpby");
}
// out[j] = <V_j, w>, j < k (V_j = V + j*ld), out on the device; w += sum_j coef[j] V_j (coef on the device).
// Building blocks of the Krylov drivers below, exported for the nonlinear mixers (opencmp_b200/mixing.py).
extern "C" int ocmp_mdot(long long n, const double* V, long long ld, int k, const double* w, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 0) return 0;
    cudaMemsetAsync(out, 0, sizeof(double) * k, st);
    ocmp_prof_begin(PROF_MDOT, st);
    for (int j0 = 0; j0 < k && n > 0; j0 += 8) {
        const int kk = (k - j0) < 8 ? (k - j0) : 8;
        k_mdot<8><<<grid_for(n), 256, 0, st>>>(n, V + (long long)j0 * ld, ld, kk, w, out + j0, nullptr, nullptr);
    }
    ocmp_prof_end(PROF_MDOT, st);
    return ocmp_check("ocmp_mdot");
}
extern "C" int ocmp_maxpy(long long n, const double* V, long long ld, int k, const double* coef, double* w,
                          void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 0 || n <= 0) return 0;
    ProfScope ps(PROF_MAXPY, st);
    k_maxpy<<<grid_for(n), 256, sizeof(double) * k, st>>>(n, V, ld, k, coef, w, 1.0);
    return ocmp_check("ocmp_maxpy");
}
extern "C" int ocmp_masked_assign(long long n, double* dst, const double* src, const double* inv, const double* mask,
                                  void* stream) {
    ProfScope ps(PROF_VEC, (cudaStream_t)stream);
    if (n > 0) k_masked_assign<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(n, dst, src, inv, mask);
    return ocmp_check("ocmp_masked_assign");
}

// ---- preconditioners -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jacobi_setup(int n, const int* __restrict__ diagpos,
                                                      const double* __restrict__ vals, const double* __restrict__ fm,
                                                      double* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = vals[diagpos[i]];
    const bool free_ = fm ? fm[i] > 0.0 : true;
    dinv[i] = (free_ && d != 0.0) ? 1.0 / d : 0.0;
}

extern "C" int ocmp_jacobi_setup(int nrows, const int* diagpos, const double* vals, const double* freemask,
                                 double* dinv, void* stream) {
    ProfScope ps(PROF_SETUP, (cudaStream_t)stream);
    if (nrows > 0)
        k_jacobi_setup<<<(nrows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(nrows, diagpos, vals, freemask, dinv);
    return ocmp_check("ocmp_jacobi_setup");
}

// One CTA per patch: gather the dense block straight from the CSR matrix (binary search of each column in its row),
// impose identity rows/cols on constrained or padded dofs, invert it IN PLACE by Gauss-Jordan with partial pivoting
// in shared memory and store the inverse transposed (inv[j*bs + i] = (A^-1)_{ij}) for coalesced application.
template <typename OutT>
__global__ void __launch_bounds__(256) k_asm_setup(int npatch, int bs, const int* __restrict__ pdofs,
                                                   const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                   const double* __restrict__ vals, const double* __restrict__ fm,
                                                   OutT* __restrict__ inv) {
    extern __shared__ double M[];          // bs x bs, then fcol[bs], then piv[bs] (int)
    double* fcol = M + bs * bs;
    int* piv = reinterpret_cast<int*>(fcol + bs);
    __shared__ int spiv;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < bs * bs; idx += blockDim.x) {
            const int i = idx / bs, j = idx % bs;
            const int di = pdofs[(long long)p * bs + i], dj = pdofs[(long long)p * bs + j];
            const bool fi = di >= 0 && (!fm || fm[di] > 0.0), fj = dj >= 0 && (!fm || fm[dj] > 0.0);
            double v = 0.0;
            if (!fi || !fj) v = (i == j) ? 1.0 : 0.0;
            else {
                int lo = __ldg(rowptr + di), hi = __ldg(rowptr + di + 1) - 1;
                while (lo <= hi) {
                    const int mid = lo + ((hi - lo) >> 1);     // lo + hi overflows int32 once nnz > 2^30
                    const int c = __ldg(colidx + mid);
                    if (c == dj) { v = vals[mid]; break; }
                    if (c < dj) lo = mid + 1; else hi = mid - 1;
                }
            }
            M[idx] = v;
        }
        __syncthreads();
        for (int k = 0; k < bs; ++k) {
            if (threadIdx.x < 32) {
                double best = -1.0;
                int bi = k;
                for (int i = k + threadIdx.x; i < bs; i += 32) {
                    const double a = fabs(M[i * bs + k]);
                    if (a > best) { best = a; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_down_sync(0xffffffffu, best, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (ob > best) { best = ob; bi = oi; }
                }
                if (threadIdx.x == 0) { spiv = bi; piv[k] = bi; }
            }
            __syncthreads();
            const int pr = spiv;
            if (pr != k)
                for (int j = threadIdx.x; j < bs; j += blockDim.x) {
                    const double t = M[k * bs + j];
                    M[k * bs + j] = M[pr * bs + j];
                    M[pr * bs + j] = t;
                }
            __syncthreads();
            // an exactly vanishing pivot even after the row interchange means the patch matrix is singular (an enclosed
            // pressure patch with all velocities constrained): that dof is dropped — zero row and column in the stored
            // inverse, i.e. no correction from this patch — instead of dividing by 0. No relative threshold: the
            // reference's 1e-10 regularisations make legitimate pivots tiny against the patch's largest entry.
            const double pv = M[k * bs + k];
            const double ip = fabs(pv) > 1e-300 ? 1.0 / pv : 0.0;
            for (int i = threadIdx.x; i < bs; i += blockDim.x) fcol[i] = (i == k) ? 0.0 : M[i * bs + k];
            __syncthreads();
            for (int j = threadIdx.x; j < bs; j += blockDim.x) M[k * bs + j] = (j == k) ? ip : M[k * bs + j] * ip;
            for (int i = threadIdx.x; i < bs; i += blockDim.x)
                if (i != k) M[i * bs + k] = 0.0;
            __syncthreads();
            for (int idx = threadIdx.x; idx < bs * bs; idx += blockDim.x) {
                const int i = idx / bs, j = idx % bs;
                if (i != k) M[idx] = fma(-fcol[i], M[k * bs + j], M[idx]);
            }
            __syncthreads();
        }
        for (int k = bs - 1; k >= 0; --k) {       // undo the row interchanges as column interchanges
            const int pr = piv[k];
            if (pr != k)
                for (int i = threadIdx.x; i < bs; i += blockDim.x) {
                    const double t = M[i * bs + k];
                    M[i * bs + k] = M[i * bs + pr];
                    M[i * bs + pr] = t;
                }
            __syncthreads();
        }
        for (int idx = threadIdx.x; idx < bs * bs; idx += blockDim.x) {
            const int j = idx / bs, i = idx % bs;
            inv[(long long)p * bs * bs + idx] = ocmp_store<OutT>(M[i * bs + j]);
        }
    }
}

static int invert_dispatch(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                           const double* fm, double* inv, int* flag, const int* pos, cudaStream_t st) {
    return ocmp_patch_invert_registers(npatch, bs, pd, rp, ci, vals, fm, inv, flag, pos, st);
}
static int invert_dispatch(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                           const double* fm, float* inv, int* flag, const int* pos, cudaStream_t st) {
    return ocmp_patch_invert_registers_f32(npatch, bs, pd, rp, ci, vals, fm, inv, flag, pos, st);
}

static int invert_dispatch(int npatch, int bs, const int* pd, const int* rp, const int* ci, const double* vals,
                           const double* fm, __nv_bfloat16* inv, int* flag, const int* pos, cudaStream_t st) {
    return ocmp_patch_invert_registers_bf16(npatch, bs, pd, rp, ci, vals, fm, inv, flag, pos, st);
}

template <typename OutT>
static int asm_setup(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                     const double* vals, const double* freemask, OutT* inv_blocks, const int* positions,
                     void* stream) {
    if (npatch <= 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)bs * bs + bs) + sizeof(int) * bs;
    if (smem > 220 * 1024) return ocmp_fail(-3, "patch too large for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_asm_setup<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int cap = ocmp_sm_count() * 4;
    ProfScope ps(PROF_SETUP, (cudaStream_t)stream);
    {   // fast path: register-tiled Gauss-Jordan without pivoting; falls back below if a pivot vanishes
        static int* flag_dev = nullptr;
        if (!flag_dev) cudaMalloc(&flag_dev, sizeof(int));
        cudaStream_t st = (cudaStream_t)stream;
        cudaMemsetAsync(flag_dev, 0, sizeof(int), st);
        if (positions && invert_dispatch(npatch, bs, patch_dofs, rowptr, colidx, vals, freemask, inv_blocks, flag_dev,
                                         positions, st)) {
            int flag = 0;
            cudaMemcpyAsync(&flag, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            if (!flag) return ocmp_check("ocmp_asm_setup");
        }
    }
    k_asm_setup<OutT><<<npatch < cap ? npatch : cap, 256, smem, (cudaStream_t)stream>>>(npatch, bs, patch_dofs, rowptr,
                                                                                        colidx, vals, freemask,
                                                                                        inv_blocks);
    return ocmp_check("ocmp_asm_setup");
}

extern "C" int ocmp_asm_setup(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                              const double* vals, const double* freemask, double* inv_blocks, const int* positions,
                              void* stream) {
    return asm_setup<double>(npatch, bs, patch_dofs, rowptr, colidx, vals, freemask, inv_blocks, positions, stream);
}

extern "C" int ocmp_asm_setup_bf16(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                                   const double* vals, const double* freemask, unsigned short* inv_blocks,
                                   const int* positions, void* stream) {
    return asm_setup<__nv_bfloat16>(npatch, bs, patch_dofs, rowptr, colidx, vals, freemask,
                                    reinterpret_cast<__nv_bfloat16*>(inv_blocks), positions, stream);
}

extern "C" int ocmp_asm_setup_f32(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                                  const double* vals, const double* freemask, float* inv_blocks, const int* positions,
                                  void* stream) {
    return asm_setup<float>(npatch, bs, patch_dofs, rowptr, colidx, vals, freemask, inv_blocks, positions, stream);
}


// z = sum over the patches of A_p^-1 r[dofs_p] scattered back: the streamed patch products into the patch-local
// buffer `ybuf` (npatch x bs doubles), then the per-dof gather through the incidence list (inc_ptr: n + 1, inc_idx:
// positions p * bs + i of every valid patch entry of the dof, ascending). Two launches, no atomics, deterministic.
static int patch_products(int npatch, int bs, const int* patch_dofs, const void* inv, int storage, const double* r,
                          double* ybuf, long long n, cudaStream_t st) {
    if (npatch <= 0) return 0;
    const double elem = storage == 0 ? 8.0 : storage == 1 ? 4.0 : 2.0;
    ocmp_prof_bytes(PROF_ASM_APPLY, elem * npatch * bs * bs + 16.0 * n);
    ProfScope ps(PROF_ASM_APPLY, st);
    int rc;
    if (storage == 2) rc = ocmp_patch_apply_y_bf16(npatch, bs, patch_dofs, (const __nv_bfloat16*)inv, r, ybuf, st);
    else if (storage == 1) rc = ocmp_patch_apply_y_f32(npatch, bs, patch_dofs, (const float*)inv, r, ybuf, st);
    else rc = ocmp_patch_apply_y(npatch, bs, patch_dofs, (const double*)inv, r, ybuf, st);
    return rc ? rc : ocmp_check("ocmp_asm_apply");
}
static int patch_gather(long long n, const int* inc_ptr, const int* inc_idx, const double* ybuf, const double* w,
                        const double* m, double scale, double* z, int accumulate, cudaStream_t st) {
    ProfScope ps(PROF_VEC, st);
    ocmp_patch_gather(n, inc_ptr, inc_idx, ybuf, w, m, scale, z, accumulate, st);
    return ocmp_check("ocmp_asm_apply (gather)");
}

static int asm_apply(int npatch, int bs, const int* patch_dofs, const void* inv, int storage, const int* inc_ptr,
                     const int* inc_idx, double* ybuf, const double* r, double* z, long long n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (npatch <= 0) { cudaMemsetAsync(z, 0, sizeof(double) * n, st); return 0; }
    if (int rc = patch_products(npatch, bs, patch_dofs, inv, storage, r, ybuf, n, st)) return rc;
    return patch_gather(n, inc_ptr, inc_idx, ybuf, nullptr, nullptr, 1.0, z, 0, st);
}

extern "C" int ocmp_asm_apply(int npatch, int bs, const int* patch_dofs, const double* inv_blocks, const int* inc_ptr,
                              const int* inc_idx, double* ybuf, const double* r, double* z, long long n,
                              void* stream) {
    return asm_apply(npatch, bs, patch_dofs, inv_blocks, 0, inc_ptr, inc_idx, ybuf, r, z, n, stream);
}
extern "C" int ocmp_asm_apply_f32(int npatch, int bs, const int* patch_dofs, const float* inv_blocks,
                                  const int* inc_ptr, const int* inc_idx, double* ybuf, const double* r, double* z,
                                  long long n, void* stream) {
    return asm_apply(npatch, bs, patch_dofs, inv_blocks, 1, inc_ptr, inc_idx, ybuf, r, z, n, stream);
}
extern "C" int ocmp_asm_apply_bf16(int npatch, int bs, const int* patch_dofs, const unsigned short* inv_blocks,
                                   const int* inc_ptr, const int* inc_idx, double* ybuf, const double* r, double* z,
                                   long long n, void* stream) {
    return asm_apply(npatch, bs, patch_dofs, inv_blocks, 2, inc_ptr, inc_idx, ybuf, r, z, n, stream);
}

// ---- Krylov drivers ---------------------------------------------------------------------------------------------
namespace {
// pinned host staging for the Gram-Schmidt coefficients (asynchronous device -> host copies)
double* pinned_scalars(size_t count) {
    static double* buf = nullptr;
    static size_t cap = 0;
    if (count > cap) {
        if (buf) cudaFreeHost(buf);
        cap = count < 1024 ? 1024 : 2 * count;
        if (cudaMallocHost(&buf, sizeof(double) * cap) != cudaSuccess) { buf = nullptr; cap = 0; }
    }
    return buf;
}

struct Ctx {
    const ocmp_system* s;
    cudaStream_t st;
    long long n;
    double* dscal;      // device scratch scalars
    double hscal[64];
    mutable int failed = 0;     // a matrix-free operator callback reported an error

    // the rows an operator application has to compute: a rank's own rows when the halo exchange supplies the others
    static const int* row_list(const ocmp_system* sy) { return (sy->halo_fwd && sy->spmv_rows) ? sy->spmv_rows : nullptr; }
    static int active_rows(const ocmp_system* sy) { return row_list(sy) ? sy->n_spmv_rows : sy->nrows; }
    // grouped product: all rows -> run_grouped nodes; with a row list -> the count of listed rows below run_shift
    static SpmvRuns runs_of(const ocmp_system* sy) {
        const int grouped = !sy->run_len ? 0 : row_list(sy) ? sy->n_spmv_groups : sy->run_grouped;
        return SpmvRuns{sy->run_len, sy->run_shift, sy->run_nc, grouped};
    }
    static void had(cudaStream_t st, long long n, const double* a, const double* m, const double* r, double* z,
                    double scale, int accumulate) {
        ProfScope ps(PROF_VEC, st);
        if (n > 0) k_had<<<grid_for(n), 256, 0, st>>>(n, a, m, r, z, scale, accumulate);
    }
    // y = m .* (A x) of the Krylov operator (mask optional); element-partitioned: followed by the ghost refresh, so
    // y is consistent like x
    void A(const double* x, double* y, bool masked) const {
        if (s->apply_fn) {      // matrix-free operator: the form's action (quadrature kernels), no CSR values read
            if (((ocmp_apply_fn)s->apply_fn)(s->apply_ctx, x, y, (void*)st) != 0 && !failed) failed = 1;
            if (masked && s->freemask) had(st, n, nullptr, s->freemask, y, y, 1.0, 0);
        } else {
            spmv_ep(PROF_SPMV, masked && s->freemask ? EP_MASK : EP_PLAIN, active_rows(s), s->rowptr, s->colidx,
                    s->vals, nullptr, x, y, nullptr, s->freemask, nullptr, st, row_list(s), runs_of(s), s->nrows);
        }
        if (s->halo_fwd) ocmp_halo_run(s->halo_fwd - 1, y, 0, st);
    }
    // r = m .* (b - A x) of the Krylov operator
    void residual(const double* b, const double* x, double* r) const {
        if (s->apply_fn) {
            A(x, r, false);
            ProfScope ps(PROF_VEC, st);
            k_resid<<<grid_for(n), 256, 0, st>>>(n, b, s->freemask, r);
        } else {
            spmv_ep(PROF_SPMV, EP_RESID, active_rows(s), s->rowptr, s->colidx, s->vals, nullptr, x, r, b, s->freemask,
                    nullptr, st, row_list(s), runs_of(s), s->nrows);
            if (s->halo_fwd) ocmp_halo_run(s->halo_fwd - 1, r, 0, st);
        }
    }
    // One application of the smoother / local preconditioner of system `sy`: z (+)= scale * S r (r masked, result
    // masked). `tmp` (n doubles) is only needed for accumulate on an element-partitioned patch smoother.
    static void smooth(const ocmp_system* sy, cudaStream_t st, const double* r, double* z, double scale,
                       int accumulate, double* tmp) {
        const long long n = sy->nrows;
        if (sy->pre_kind == 1) {
            had(st, n, sy->dinv, nullptr, r, z, scale, accumulate);
        } else if (sy->pre_kind == 2 || sy->pre_kind == 3) {
            patch_products(sy->npatch, sy->bs, sy->patch_dofs, sy->inv_blocks, sy->inv_storage, r, sy->patch_ybuf, n,
                           st);
            if (sy->halo_sum && accumulate) {
                // a rank applies the patches of the vertices it owns; DOFs shared with a neighbour get the sum
                patch_gather(n, sy->patch_inc_ptr, sy->patch_inc_idx, sy->patch_ybuf, sy->patch_weight, sy->freemask,
                             scale, tmp, 0, st);
                ocmp_halo_run(sy->halo_sum - 1, tmp, 1, st);
                ocmp_axpby(n, 1.0, tmp, 1.0, z, st);
            } else {
                patch_gather(n, sy->patch_inc_ptr, sy->patch_inc_idx, sy->patch_ybuf, sy->patch_weight, sy->freemask,
                             scale, z, accumulate, st);
                if (sy->halo_sum) ocmp_halo_run(sy->halo_sum - 1, z, 1, st);
            }
        } else if (sy->pre_kind == 4) {
            // explicit inverse of the coarsest level, stored as a (dense) CSR matrix
            spmv_ep(PROF_SPMV_MG, accumulate ? EP_ADD : EP_MASK, sy->nrows, sy->inv_rowptr, sy->inv_colidx,
                    sy->inv_vals, nullptr, r, z, nullptr, sy->freemask, nullptr, st);
            // replicated coarsest level: every rank solved with its own copy; the owners' values win
            if (sy->halo_fwd) ocmp_halo_run(sy->halo_fwd - 1, z, 0, st);
        } else if (sy->pre_kind == 5 && sy->direct) {
            // Preconditioner(a, 'direct'): exact solve with the factorised free-free block (zero on constrained dofs)
            const ocmp_band_lu* f = sy->direct;
            ocmp_band_gather(sy->nrows, f->perm, r, f->rhs, st);
            ocmp_band_solve(f->n, f->kl, f->ku, f->ubw, f->ab, f->ipiv, f->rhs, st);
            ocmp_band_scatter(sy->nrows, f->perm, f->rhs, z, accumulate, st);
        } else {
            had(st, n, nullptr, sy->freemask, r, z, scale, accumulate);
        }
    }
    // r = m .* m2 .* (b - A x) on a multigrid level: FP32-stored copy of the level matrix when the host provided one
    static void level_residual(const ocmp_system* sy, const double* b, const double* x, double* r, const double* m2,
                               bool refresh, cudaStream_t st) {
        spmv_ep(PROF_SPMV_MG, EP_RESID, active_rows(sy), sy->rowptr, sy->colidx, sy->vals, sy->vals32, x, r, b,
                sy->freemask, m2, st, row_list(sy), runs_of(sy), sy->nrows);
        if (refresh && sy->halo_fwd) ocmp_halo_run(sy->halo_fwd - 1, r, 0, st);
    }
    // V-cycle on level l: x = MG(b); b is masked on entry, x is masked on exit. Launches per level and cycle with
    // nu = 1: smoother 2 + residual 1 + restriction 1 + prolongation 1 + residual 1 + smoother 2.
    static void vcycle(const ocmp_mg_level* L, int l, cudaStream_t st, const double* b, double* x) {
        const ocmp_mg_level& lv = L[l];
        const ocmp_system* sy = &lv.sys;
        const long long n = sy->nrows;
        if (l == 0) { smooth(sy, st, b, x, 1.0, 0, nullptr); return; }
        double* r = lv.work + 2 * n;
        double* t = lv.work + 3 * n;
        smooth(sy, st, b, x, lv.omega, 0, t);                 // x = omega S b   (zero initial guess)
        for (int s = 1; s < lv.nu; ++s) {
            level_residual(sy, b, x, r, nullptr, true, st);
            smooth(sy, st, r, x, lv.omega, 1, t);
        }
        // residual to restrict: only the owned entries are used, so no ghost refresh after this product
        level_residual(sy, b, x, r, sy->owned, false, st);
        const ocmp_mg_level& lc = L[l - 1];
        const long long nc = lc.sys.nrows;
        double* xc = lc.work;
        double* bc = lc.work + nc;
        spmv_ep(PROF_SPMV_MG, EP_MASK, (int)nc, lv.r_rowptr, lv.r_colidx, lv.r_vals, nullptr, r, bc, nullptr,
                lc.sys.freemask, nullptr, st);
        if (lv.restrict_sum) ocmp_halo_run(lv.restrict_sum - 1, bc, 1, st);
        vcycle(L, l - 1, st, bc, xc);
        if (lv.handover) ocmp_halo_run(lv.handover - 1, xc, 0, st);
        spmv_ep(PROF_SPMV_MG, EP_ADD, sy->nrows, lv.p_rowptr, lv.p_colidx, lv.p_vals, nullptr, xc, x, nullptr,
                sy->freemask, nullptr, st);                  // x += m .* (P xc)
        for (int s = 0; s < lv.nu; ++s) {
            level_residual(sy, b, x, r, nullptr, true, st);
            smooth(sy, st, r, x, lv.omega, 1, t);
        }
    }
    // z = P (already masked r); result masked
    void P(const double* r, double* z) const {
        if (s->pre_kind == 3 && s->nlevels > 1) vcycle(s->levels, s->nlevels - 1, st, r, z);
        else smooth(s, st, r, z, 1.0, 0, nullptr);
    }
    // <x, y> over the owned entries, summed over the ranks (one FP64 all-reduce) when element-partitioned
    double dot(const double* x, const double* y) {
        cudaMemsetAsync(dscal, 0, sizeof(double), st);
        {
            ProfScope ps(PROF_VEC, st);
            if (n > 0) k_dot<<<grid_for(n), 256, 0, st>>>(n, x, y, dscal, s->owned);
        }
        if (s->owned) ocmp_allreduce_sum(dscal, 1, st);
        cudaMemcpyAsync(hscal, dscal, sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        return hscal[0];
    }
    void axpby(double a, const double* x, double b, double* y) const { ocmp_axpby(n, a, x, b, y, st); }
    // dout[j] = <V_j, w>, j < k, and dout[k] = <extra, w> when extra is given; all-reduced; stays on the device
    void mdot_dev(const double* V, int k, const double* w, const double* extra, double* dout) {
        const int tot = k + (extra ? 1 : 0);
        cudaMemsetAsync(dout, 0, sizeof(double) * tot, st);
        ocmp_prof_begin(PROF_MDOT, st);
        for (int j0 = 0; j0 < k || (j0 == 0 && extra); j0 += 8) {
            const int kk = (k - j0) < 8 ? (k - j0) : 8;
            const bool last = j0 + 8 >= k;
            k_mdot<8><<<grid_for(n), 256, 0, st>>>(n, V + (long long)j0 * n, n, kk, w, dout + j0, s->owned,
                                                   last ? extra : nullptr);
        }
        ocmp_prof_end(PROF_MDOT, st);
        if (s->owned) ocmp_allreduce_sum(dout, tot, st);
    }
    void maxpy_dev(const double* V, int k, const double* dc, double sign, double* w) {
        ProfScope ps(PROF_MAXPY, st);
        k_maxpy<<<grid_for(n), 256, sizeof(double) * k, st>>>(n, V, n, k, dc, w, sign);
    }
    void maxpy(const double* V, int k, const double* hc, double* w) {
        cudaMemcpyAsync(dscal, hc, sizeof(double) * k, cudaMemcpyHostToDevice, st);
        maxpy_dev(V, k, dscal, 1.0, w);
    }
};
}  // namespace

// relative (preconditioned) residual after every iteration of the last ocmp_krylov call
static std::vector<double> g_history;
extern "C" int ocmp_krylov_history(double* out, int cap) {
    const int n = (int)g_history.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = g_history[i];
    return n;
}

extern "C" long long ocmp_krylov_work_len(int nrows, int kind, int restart) {
    const long long n = nrows;
    if (kind == 0) return 4 * n + 64;
    if (kind == 1) return (long long)(restart + 2) * n + 2 * n + 2 * (restart + 2) + 128;
    if (kind == 3) return 9 * n + 64;
    return 2 * n + 64;
}

extern "C" int ocmp_krylov(const ocmp_system* sys, int kind, const double* b, double* x, double tol, int maxit,
                           int restart, double damp, double* work, long long work_len, int* iters, double* resid,
                           void* stream) {
    const long long n = sys->nrows;
    if (work_len < ocmp_krylov_work_len(sys->nrows, kind, restart)) return ocmp_fail(-4, "krylov work array too small");
    Ctx c;
    c.s = sys; c.st = (cudaStream_t)stream; c.n = n;
    int it = 0;
    double res = 0.0;
    g_history.clear();
    if (kind == 0) {                       // preconditioned CG on the free dofs
        double *r = work, *z = work + n, *p = work + 2 * n, *Ap = work + 3 * n;
        c.dscal = work + 4 * n;
        c.residual(b, x, r);
        c.P(r, z);
        c.axpby(1.0, z, 0.0, p);
        double rz = c.dot(r, z);
        const double err0 = sqrt(fabs(rz));
        res = err0;
        if (rz != 0.0) {
            for (it = 0; it < maxit;) {
                c.A(p, Ap, true);
                const double pAp = c.dot(p, Ap);
                const double alpha = rz / pAp;
                c.axpby(alpha, p, 1.0, x);
                c.axpby(-alpha, Ap, 1.0, r);
                c.P(r, z);
                const double rzn = c.dot(r, z);
                c.axpby(1.0, z, rzn / rz, p);
                ++it;
                res = sqrt(fabs(rzn));
                g_history.push_back(res / err0);
                rz = rzn;
                if (res < tol * err0 || rzn == 0.0) break;
            }
        }
    } else if (kind == 1) {                // left-preconditioned restarted GMRES, CGS2 orthogonalisation
        // Per iteration: operator + preconditioner, two batched Gram-Schmidt passes whose coefficients stay on the
        // device (the corrections read them there), the squared norm of the new vector riding along with the second
        // pass (||w - V h2||^2 = ||w||^2 - ||h2||^2), correction + normalisation fused — ONE host synchronisation and,
        // element-partitioned, TWO all-reduces per iteration. The Hessenberg / Givens recurrences run on the host
        // from the coefficients copied back asynchronously into pinned memory.
        const int m = restart;
        double* V = work;
        double* w = work + (long long)(m + 1) * n;
        double* t = w + n;
        c.dscal = t + n;                                      // [0, 64): dot(); then two coefficient arrays of m + 2
        double* d1 = c.dscal + 64;
        double* d2 = d1 + (m + 2);
        double* pin = pinned_scalars(2 * (size_t)(m + 2));
        if (!pin) return ocmp_fail(-6, "ocmp_krylov: cannot allocate pinned host memory");
        double* h1 = pin;
        double* h2 = pin + (m + 2);
        std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), yv(m);
        double beta0 = -1.0;
        bool done = false;
        while (!done && it < maxit) {
            c.residual(b, x, t);
            c.P(t, V);
            double beta = sqrt(c.dot(V, V));
            if (beta0 < 0.0) beta0 = beta;
            res = beta;
            if (beta == 0.0 || beta < tol * beta0) break;
            c.axpby(0.0, V, 1.0 / beta, V);      // V0 *= 1/beta
            std::fill(g.begin(), g.end(), 0.0);
            g[0] = beta;
            int k = 0;
            for (; k < m && it < maxit; ++k) {
                double* vn = V + (long long)(k + 1) * n;
                c.A(V + (long long)k * n, t, true);
                c.P(t, w);
                c.mdot_dev(V, k + 1, w, nullptr, d1);
                cudaMemcpyAsync(h1, d1, sizeof(double) * (k + 1), cudaMemcpyDeviceToHost, c.st);
                c.maxpy_dev(V, k + 1, d1, -1.0, w);
                c.mdot_dev(V, k + 1, w, w, d2);              // second pass; d2[k + 1] = <w, w>
                cudaMemcpyAsync(h2, d2, sizeof(double) * (k + 2), cudaMemcpyDeviceToHost, c.st);
                {
                    ProfScope ps(PROF_MAXPY, c.st);
                    k_gs_finish<<<grid_for(n), 256, sizeof(double) * (k + 2), c.st>>>(n, V, n, k + 1, d2, w, vn);
                }
                cudaStreamSynchronize(c.st);
                double ww = h2[k + 1], hh = ww;
                for (int j = 0; j <= k; ++j) hh -= h2[j] * h2[j];
                double hn = hh > 0.0 ? sqrt(hh) : 0.0;
                if (hn > 0.0 && hh < 1e-4 * ww) {
                    // the second pass removed a large part of w (cancellation in ww - |h2|^2): measure the stored
                    // vector instead and renormalise it
                    const double nv = sqrt(c.dot(vn, vn));
                    if (nv > 0.0) { c.axpby(0.0, vn, 1.0 / nv, vn); hn *= nv; } else hn = 0.0;
                }
                for (int j = 0; j <= k; ++j) H[(size_t)j * m + k] = h1[j] + h2[j];
                H[(size_t)(k + 1) * m + k] = hn;
                for (int i = 0; i < k; ++i) {
                    const double a = H[(size_t)i * m + k], bb = H[(size_t)(i + 1) * m + k];
                    H[(size_t)i * m + k] = cs[i] * a + sn[i] * bb;
                    H[(size_t)(i + 1) * m + k] = -sn[i] * a + cs[i] * bb;
                }
                const double a = H[(size_t)k * m + k], bb = H[(size_t)(k + 1) * m + k];
                const double den = hypot(a, bb);
                cs[k] = den == 0.0 ? 1.0 : a / den;
                sn[k] = den == 0.0 ? 0.0 : bb / den;
                H[(size_t)k * m + k] = cs[k] * a + sn[k] * bb;
                H[(size_t)(k + 1) * m + k] = 0.0;
                g[k + 1] = -sn[k] * g[k];
                g[k] = cs[k] * g[k];
                ++it;
                res = fabs(g[k + 1]);
                g_history.push_back(res / beta0);
                if (res < tol * beta0 || hn == 0.0) { done = true; ++k; break; }
            }
            for (int i = k - 1; i >= 0; --i) {
                double sacc = g[i];
                for (int j = i + 1; j < k; ++j) sacc -= H[(size_t)i * m + j] * yv[j];
                yv[i] = sacc / H[(size_t)i * m + i];
            }
            c.maxpy(V, k, yv.data(), x);
        }
    } else if (kind == 2) {                // damped preconditioned Richardson
        double *r = work, *z = work + n;
        c.dscal = work + 2 * n;
        double r0 = -1.0;
        for (it = 0; it < maxit; ++it) {
            c.residual(b, x, r);
            res = sqrt(c.dot(r, r));
            if (r0 < 0.0) r0 = res;
            if (res < tol * r0 || res == 0.0) break;
            c.P(r, z);
            c.axpby(damp, z, 1.0, x);
        }
    } else if (kind == 3) {                // preconditioned MINRES (Paige-Saunders; symmetric A, SPD preconditioner)
        // the Lanczos process in the P-inner product with the two-term Givens update of the solution; stops on the
        // recurrence value of the P-norm of the residual, relative to the initial one (what NGSolve's MinRes monitors)
        double* v0 = work;          double* v1 = work + n;     double* v2 = work + 2 * n;
        double* z1 = work + 3 * n;  double* z2 = work + 4 * n;
        double* w0 = work + 5 * n;  double* w1 = work + 6 * n; double* w2 = work + 7 * n;
        double* Az = work + 8 * n;
        c.dscal = work + 9 * n;
        cudaMemsetAsync(v0, 0, sizeof(double) * n, c.st);
        cudaMemsetAsync(w0, 0, sizeof(double) * n, c.st);
        cudaMemsetAsync(w1, 0, sizeof(double) * n, c.st);
        c.residual(b, x, v1);
        c.P(v1, z1);
        double gamma0 = 1.0, gamma1 = sqrt(fabs(c.dot(z1, v1)));
        double eta = gamma1, s0 = 0.0, s1 = 0.0, c0 = 1.0, c1 = 1.0;
        const double err0 = gamma1;
        res = err0;
        if (gamma1 != 0.0) {
            for (it = 0; it < maxit;) {
                c.axpby(0.0, z1, 1.0 / gamma1, z1);                 // z_j normalised
                c.A(z1, Az, true);
                const double delta = c.dot(Az, z1);
                // v_{j+1} = A z_j - (delta / gamma_j) v_j - (gamma_j / gamma_{j-1}) v_{j-1}
                c.axpby(1.0, Az, 0.0, v2);
                c.axpby(-delta / gamma1, v1, 1.0, v2);
                c.axpby(-gamma1 / gamma0, v0, 1.0, v2);
                c.P(v2, z2);
                const double gamma2 = sqrt(fabs(c.dot(z2, v2)));
                const double a0 = c1 * delta - c0 * s1 * gamma1;
                const double a1 = sqrt(a0 * a0 + gamma2 * gamma2);
                const double a2 = s1 * delta + c0 * c1 * gamma1;
                const double a3 = s0 * gamma1;
                c0 = c1; s0 = s1;
                c1 = a1 == 0.0 ? 1.0 : a0 / a1;
                s1 = a1 == 0.0 ? 0.0 : gamma2 / a1;
                // w_{j+1} = (z_j - a3 w_{j-1} - a2 w_j) / a1
                c.axpby(1.0, z1, 0.0, w2);
                c.axpby(-a3, w0, 1.0, w2);
                c.axpby(-a2, w1, 1.0, w2);
                if (a1 != 0.0) c.axpby(0.0, w2, 1.0 / a1, w2);
                c.axpby(c1 * eta, w2, 1.0, x);
                eta = -s1 * eta;
                ++it;
                res = fabs(eta);
                g_history.push_back(res / err0);
                if (res < tol * err0 || gamma2 == 0.0) break;
                double* tv = v0; v0 = v1; v1 = v2; v2 = tv;
                double* tz = z1; z1 = z2; z2 = tz;
                double* tw = w0; w0 = w1; w1 = w2; w2 = tw;
                gamma0 = gamma1; gamma1 = gamma2;
            }
        }
    } else return ocmp_fail(-1, "unknown krylov kind");
    cudaStreamSynchronize(c.st);
    if (iters) *iters = it;
    if (resid) *resid = res;
    if (c.failed) return ocmp_fail(-5, "matrix-free operator callback failed inside ocmp_krylov");
    return ocmp_check("ocmp_krylov");
}
