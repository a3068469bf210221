// Assembly kernels of the opencmp_b200 backend (sm_100a).
//
//   k_coef      evaluates the coefficient bytecode of one integral at every (item, quadrature point):
//               geometry (x, n, h, measure), field values from DOF vectors, then the register machine -> D slots.
//   k_contract  forms the local matrices  A = sum_q B^T (D B)  with physical basis tables, Z = D B rows and the
//               accumulators in shared memory, and scatter-adds them through the element -> nnz map (no colouring).
//   k_lin       the same for linear forms (local vectors, scatter through the cell dof lists).
//
// Stand in for NGSolve's SymbolicBilinearFormIntegrator / SymbolicLinearFormIntegrator element loops that
// `a.Assemble()` / `L.Assemble()` run (reference opencmp/solvers/base_solver.py:368-377).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "../../include/opencmp_b200.h"
#include "ocmp_common.cuh"

enum { OP_CONST = 0, OP_PARAM, OP_COORD, OP_NORMAL, OP_MESHSIZE, OP_FIELD, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG,
       OP_ABS, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_EXP, OP_LOG, OP_POW, OP_IFPOS, OP_MIN, OP_MAX, OP_TANH, OP_ERF,
       OP_FLOOR, OP_CEIL, OP_ROUND, OP_TRUNC, OP_SGN, OP_ATAN, OP_OUT, OP_MOV, OP_MEASURE };

#define MAX_FSLOTS 40
#define MAX_ROWS 12

template <int DIM> struct GeoT { static constexpr int GS = DIM + 2 * DIM * DIM + 1; };

// physical rows of one block from its reference-row sums R (scalar: value + gradient; hdiv: Piola value + gradient)
template <int DIM>
__device__ __forceinline__ void phys_rows(int kind, const double* __restrict__ R, const double* __restrict__ J,
                                          const double* __restrict__ Jinv, double det, double* __restrict__ PH) {
    if (kind == 0) {
        PH[0] = R[0];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; ++b) s += Jinv[b * DIM + a] * R[1 + b];
            PH[1 + a] = s;
        }
    } else {
        const double idet = 1.0 / det;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; ++b) s += J[c * DIM + b] * R[b];
            PH[c] = s * idet;
        }
        // grad_{c,a} = sum_{b,m} J[c][b] R[DIM + b*DIM + m] Jinv[m][a] / det
        double T[DIM * DIM];
#pragma unroll
        for (int b = 0; b < DIM; ++b)
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < DIM; ++m) s += R[DIM + b * DIM + m] * Jinv[m * DIM + a];
                T[b * DIM + a] = s;
            }
#pragma unroll
        for (int c = 0; c < DIM; ++c)
#pragma unroll
            for (int a = 0; a < DIM; ++a) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < DIM; ++b) s += J[c * DIM + b] * T[b * DIM + a];
                PH[DIM + c * DIM + a] = s * idet;
            }
    }
}

template <int DIM>
__device__ __forceinline__ double facet_measure(const double* __restrict__ J, const double* __restrict__ tref) {
    double T[(DIM - 1) * DIM];
#pragma unroll
    for (int k = 0; k < DIM - 1; ++k)
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) s += J[i * DIM + a] * tref[k * DIM + a];
            T[k * DIM + i] = s;
        }
    if (DIM == 2) return sqrt(T[0] * T[0] + T[1] * T[1]);
    const double cx = T[1] * T[DIM + 2] - T[2] * T[DIM + 1];
    const double cy = T[2] * T[DIM + 0] - T[0] * T[DIM + 2];
    const double cz = T[0] * T[DIM + 1] - T[1] * T[DIM + 0];
    return sqrt(cx * cx + cy * cy + cz * cz);
}

// ---------------------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_coef(const __grid_constant__ ocmp_coef_plan P, int item0, int nitems,
                                              double* __restrict__ dbuf) {
    constexpr int GS = GeoT<DIM>::GS;
    const long long total = (long long)nitems * P.nq;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int it = (int)(gid / P.nq), q = (int)(gid % P.nq);
    const int id = P.items ? __ldg(P.items + item0 + it) : item0 + it;
    int cell[2], lf[2];
    if (P.kind == 0) {
        cell[0] = id; lf[0] = 0; cell[1] = id; lf[1] = 0;
    } else {
        cell[0] = __ldg(P.facet_cells + 2 * id); lf[0] = __ldg(P.facet_local + 2 * id);
        cell[1] = __ldg(P.facet_cells + 2 * id + 1); lf[1] = __ldg(P.facet_local + 2 * id + 1);
        if (cell[1] < 0) { cell[1] = cell[0]; lf[1] = lf[0]; }
    }
    const double* g0 = P.geo + (long long)cell[0] * GS;
    const double* J0 = g0 + DIM;
    const double* Ji0 = g0 + DIM + DIM * DIM;
    const double det0 = g0[GS - 1];
    const double* xi = P.qpts + ((P.kind == 0 ? 0 : lf[0] * P.nq) + q) * DIM;
    double X[3] = {0.0, 0.0, 0.0}, N[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        double s = g0[i];
#pragma unroll
        for (int a = 0; a < DIM; ++a) s += J0[i * DIM + a] * xi[a];
        X[i] = s;
    }
    double measure, h;
    if (P.kind == 0) {
        measure = fabs(det0);
        h = (DIM == 2) ? sqrt(measure) : cbrt(measure);
    } else {
        const double* fr = P.fref + lf[0] * DIM * DIM;
        measure = facet_measure<DIM>(J0, fr + DIM);
        double nn = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DIM; ++a) s += Ji0[a * DIM + i] * fr[a];
            N[i] = s; nn += s * s;
        }
        nn = 1.0 / sqrt(nn);
#pragma unroll
        for (int i = 0; i < DIM; ++i) N[i] *= nn;
        h = fabs(det0) / measure;
    }
    // ---- field values -------------------------------------------------------------------------------------
    double fval[MAX_FSLOTS];
    for (int g = 0; g < P.nfgroups; ++g) {
        const int* fg = P.fgroup + 8 * g;
        const int vec = fg[0], darr = fg[1], side = fg[2], kind = fg[3], nloc = fg[4], nrows = fg[5], toff = fg[6],
                  doff = fg[7];
        const int dstride = P.fgroup2[2 * g], dadd = P.fgroup2[2 * g + 1];
        const int c = cell[side];
        const double* tab = P.ftab + toff + (long long)((P.kind == 0 ? 0 : lf[side] * P.nq) + q) * nrows * nloc;
        const int* dofs = P.fdof[darr] + (long long)c * dstride + doff;
        const double* v = P.fvec[vec];
        double R[MAX_ROWS];
#pragma unroll
        for (int r = 0; r < MAX_ROWS; ++r) R[r] = 0.0;
        for (int i = 0; i < nloc; ++i) {
            const double ci = __ldg(v + (__ldg(dofs + i) + dadd));
#pragma unroll
            for (int r = 0; r < MAX_ROWS; ++r)
                if (r < nrows) R[r] += ci * __ldg(tab + r * nloc + i);
        }
        const double* gs = P.geo + (long long)c * GS;
        double PH[MAX_ROWS];
        phys_rows<DIM>(kind, R, gs + DIM, gs + DIM + DIM * DIM, gs[GS - 1], PH);
        for (int s = 0; s < P.nfslots; ++s)
            if (P.fslot[2 * s] == g) {
                const int row = P.fslot[2 * s + 1];
                double val = 0.0;
#pragma unroll
                for (int r = 0; r < MAX_ROWS; ++r)
                    if (r == row) val = PH[r];
                fval[s] = val;
            }
    }
    // ---- register machine ---------------------------------------------------------------------------------
    double r[OCMP_MAX_REGS];
    const long long ostride = total;
    const double wq = P.qw[q] * measure;
    for (int pc = 0; pc < P.ninstr; ++pc) {
        const int4 ins = __ldg(reinterpret_cast<const int4*>(P.code) + pc);
        const int op = ins.x & 0xff, d = ins.x >> 8;
        double v;
        switch (op) {
            case OP_CONST: v = P.consts[ins.y]; break;
            case OP_PARAM: v = P.params[ins.y]; break;
            case OP_COORD: v = X[ins.y]; break;
            case OP_NORMAL: v = N[ins.y]; break;
            case OP_MESHSIZE: v = h; break;
            case OP_MEASURE: v = measure; break;
            case OP_FIELD: v = fval[ins.y]; break;
            case OP_ADD: v = r[ins.y] + r[ins.z]; break;
            case OP_SUB: v = r[ins.y] - r[ins.z]; break;
            case OP_MUL: v = r[ins.y] * r[ins.z]; break;
            case OP_DIV: v = r[ins.y] / r[ins.z]; break;
            case OP_NEG: v = -r[ins.y]; break;
            case OP_ABS: v = fabs(r[ins.y]); break;
            case OP_SQRT: v = sqrt(r[ins.y]); break;
            case OP_SIN: v = sin(r[ins.y]); break;
            case OP_COS: v = cos(r[ins.y]); break;
            case OP_TAN: v = tan(r[ins.y]); break;
            case OP_EXP: v = exp(r[ins.y]); break;
            case OP_LOG: v = log(r[ins.y]); break;
            case OP_POW: v = pow(r[ins.y], r[ins.z]); break;
            case OP_IFPOS: v = r[ins.y] > 0.0 ? r[ins.z] : r[ins.w]; break;
            case OP_MIN: v = fmin(r[ins.y], r[ins.z]); break;
            case OP_MAX: v = fmax(r[ins.y], r[ins.z]); break;
            case OP_TANH: v = tanh(r[ins.y]); break;
            case OP_ERF: v = erf(r[ins.y]); break;
            case OP_FLOOR: v = floor(r[ins.y]); break;
            case OP_CEIL: v = ceil(r[ins.y]); break;
            case OP_ROUND: v = rint(r[ins.y]); break;
            case OP_TRUNC: v = trunc(r[ins.y]); break;
            case OP_SGN: v = (r[ins.y] > 0.0) - (r[ins.y] < 0.0); break;
            case OP_ATAN: v = atan(r[ins.y]); break;
            case OP_MOV: v = r[ins.y]; break;
            case OP_OUT: dbuf[(long long)ins.y * ostride + gid] = r[ins.z] * wq; continue;
            default: v = 0.0; break;
        }
        r[d] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Shared helper: physical table of local dof `il` of block (kind,nloc,nrows) at quadrature point q into out[r*nloc]
template <int DIM>
__device__ __forceinline__ void dof_phys_rows(int kind, int nloc, int nrows, const double* __restrict__ tabq, int il,
                                              const double* __restrict__ g, double* __restrict__ PH) {
    constexpr int GS = GeoT<DIM>::GS;
    double R[MAX_ROWS];
#pragma unroll
    for (int r = 0; r < MAX_ROWS; ++r) R[r] = (r < nrows) ? __ldg(tabq + r * nloc + il) : 0.0;
    phys_rows<DIM>(kind, R, g + DIM, g + DIM + DIM * DIM, g[GS - 1], PH);
}

// Contraction kernel: local matrices A = sum_q B^T (D B) in REGISTERS. One 256-thread CTA handles EB items, 256 / EB
// threads per item; every thread owns up to MAXT 4 x 4 tiles of the item's active block pairs (16 FP64 accumulators
// each) for the whole quadrature loop. Per quadrature point the physical basis tables (sB), the D values (sD) and the
// rows Z_g = sum_k D_k B[trial row_k] (sZ) are staged in shared memory; a tile update reads, per segment g of its
// pair, 4 B values and 4 Z values as four 16-byte loads (rows padded to multiples of 4; neighbouring lanes share
// tile rows / columns, so the loads are mostly broadcasts) for 16 FMAs. Structurally empty block pairs are neither
// computed nor scattered. The finished tiles are scatter-added through the element -> nnz map (no colouring).
template <int DIM, int MAXT>
__global__ void __launch_bounds__(256, (MAXT == 1 ? 2 : 1)) k_contract(const __grid_constant__ ocmp_contract_plan P, int item0, int nitems,
                                                  const double* __restrict__ dbuf, double* __restrict__ values) {
    constexpr int GS = GeoT<DIM>::GS;
    extern __shared__ double smem[];
    const int EB = P.eb, TPE = 256 / EB;
    const int nside = P.nside;
    const int QB = P.qb;                                 // quadrature points staged per barrier round
    double* sZ = smem;                                   // [EB][QB][zsz]         (zsz % 4 == 0: 32-byte aligned rows)
    double* sB = sZ + EB * QB * P.zsz;                   // [EB][QB][nside][sbsz] (sbsz % 4 == 0)
    double* sD = sB + EB * QB * nside * P.sbsz;          // [EB][QB][nslots]
    double* sG = sD + EB * QB * P.nslots;                // [EB][nside][GS]
    // [EB][4]: cell0, cell1, lf0, lf1 — rounded up to 16 bytes (the descriptor tables behind it are read with int4
    // loads; contract_smem() reserves the slack)
    int* sI = reinterpret_cast<int*>((reinterpret_cast<size_t>(sG + EB * nside * GS) + 15) & ~(size_t)15);
    int* tDof = sI + 4 * EB;                             // [nside*nloc][4]: kind | nr << 8 | padded nl << 16, nl, tab offset, sB offset(+il)
    int* tZ = tDof + 4 * nside * P.nloc;                 // [nzd][4]: entry k0, k1, sB base (incl. j), z index
    int* tEnt = tZ + 4 * P.nzd;                          // [nent][2]: slot, row offset (row * padded stride)
    int* tSeg = tEnt + 2 * P.nent;                       // [nseg][2]: test row offset (row * padded ni), z offset
    const int tid = threadIdx.x;
    const int e_own = tid / TPE, lane = tid % TPE;
    const long long dstride = (long long)nitems * P.nq;
    const int ngroups = (nitems + EB - 1) / EB;
    const int n2 = P.nloc * P.nloc;

    for (int i = tid; i < 4 * nside * P.nloc; i += 256) tDof[i] = __ldg(P.dofdesc + i);
    for (int i = tid; i < 4 * P.nzd; i += 256) tZ[i] = __ldg(P.zdesc + i);
    for (int i = tid; i < 2 * P.nent; i += 256) tEnt[i] = __ldg(P.ent + i);
    for (int i = tid; i < 2 * P.nseg; i += 256) tSeg[i] = __ldg(P.seg + i);
    for (int i = tid; i < EB * QB * P.zsz; i += 256) sZ[i] = 0.0;
    for (int i = tid; i < EB * QB * nside * P.sbsz; i += 256) sB[i] = 0.0;  // the row padding stays zero

    // my tiles: (sB offset of the tile's test rows, j0, first segment, segments, sides, row / column of the local
    // matrix, valid rows | valid columns << 8)
    int tB[MAXT], tJ[MAXT], tS0[MAXT], tNS[MAXT], tSide[MAXT], tRow[MAXT], tCol[MAXT], tRem[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int tile = lane + t * TPE;
        tNS[t] = 0; tB[t] = 0; tJ[t] = 0; tS0[t] = 0; tSide[t] = 0; tRow[t] = 0; tCol[t] = 0; tRem[t] = 0;
        if (tile < P.ntiles) {
            const int4 a = __ldg(reinterpret_cast<const int4*>(P.tiles) + 2 * tile);
            const int4 b = __ldg(reinterpret_cast<const int4*>(P.tiles) + 2 * tile + 1);
            tB[t] = a.x; tJ[t] = a.y; tS0[t] = a.z; tNS[t] = a.w;
            tSide[t] = b.x; tRow[t] = b.y; tCol[t] = b.z; tRem[t] = b.w;
        }
    }

    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        __syncthreads();
        if (tid < EB) {
            const int it = grp * EB + tid;
            int c0 = -1, c1 = -1, l0 = 0, l1 = 0;
            if (it < nitems) {
                const int id = P.items ? __ldg(P.items + item0 + it) : item0 + it;
                if (P.kind == 0) { c0 = id; }
                else {
                    c0 = __ldg(P.facet_cells + 2 * id); l0 = __ldg(P.facet_local + 2 * id);
                    c1 = __ldg(P.facet_cells + 2 * id + 1); l1 = __ldg(P.facet_local + 2 * id + 1);
                }
            }
            sI[4 * tid] = c0; sI[4 * tid + 1] = c1; sI[4 * tid + 2] = l0; sI[4 * tid + 3] = l1;
        }
        __syncthreads();
        for (int idx = tid; idx < EB * nside * GS; idx += 256) {
            const int e = idx / (nside * GS), rem = idx % (nside * GS), s = rem / GS, k = rem % GS;
            const int c = sI[4 * e + s];
            sG[idx] = (c >= 0) ? __ldg(P.geo + (long long)c * GS + k) : (k == GS - 1 ? 1.0 : 0.0);
        }
        const int it_own = grp * EB + e_own;
        double acc[MAXT][16];
#pragma unroll
        for (int t = 0; t < MAXT; ++t)
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[t][k] = 0.0;

        // Quadrature loop in rounds of QB points: the staging phases (tables, D, Z rows) have far fewer independent
        // work items per point than the CTA has threads, so QB points are staged together between two barriers.
        double* bE = sB + e_own * QB * nside * P.sbsz;
        double* dE = sD + e_own * QB * P.nslots;
        double* zE = sZ + e_own * QB * P.zsz;
        const int ncol = nside * P.nloc;
        for (int q0 = 0; q0 < P.nq; q0 += QB) {
            const int nqq = min(QB, P.nq - q0);
            __syncthreads();
            // D values of these quadrature points (nqq consecutive doubles per slot)
            for (int sl = lane; sl < P.nslots; sl += TPE) {
                const double* src = dbuf + (long long)sl * dstride + (long long)it_own * P.nq + q0;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq)
                    if (qq < nqq) dE[qq * P.nslots + sl] = (it_own < nitems) ? __ldg(src + qq) : 0.0;
            }
            // physical basis tables of my item: one (side, dof) column per thread, all points of the round
            for (int sd = lane; sd < ncol; sd += TPE) {
                const int4 dd = *reinterpret_cast<const int4*>(tDof + 4 * sd);
                const int s = sd >= P.nloc ? 1 : 0;
                const int kind = dd.x & 0xff, nr = (dd.x >> 8) & 0xff, nlp = dd.x >> 16, nl = dd.y;
                const int c = sI[4 * e_own + s];
                double* out0 = bE + s * P.sbsz + dd.w;
                if (c < 0) {
                    for (int qq = 0; qq < nqq; ++qq)
                        for (int r = 0; r < nr; ++r) out0[qq * nside * P.sbsz + r * nlp] = 0.0;
                    continue;
                }
                const int lfq0 = (P.kind == 0 ? 0 : sI[4 * e_own + 2 + s] * P.nq) + q0;
                const double* tq0 = P.tab + dd.z + (long long)lfq0 * nr * nl;        // dd.z includes the local dof
                const double* g = sG + (e_own * nside + s) * GS;
                for (int qq = 0; qq < nqq; ++qq) {
                    const double* tq = tq0 + (long long)qq * nr * nl;
                    double* out = out0 + qq * nside * P.sbsz;
                    if (kind == 0) {
                        double R[1 + DIM], PH[MAX_ROWS];
#pragma unroll
                        for (int r = 0; r < 1 + DIM; ++r) R[r] = __ldg(tq + r * nl);
                        phys_rows<DIM>(0, R, g + DIM, g + DIM + DIM * DIM, g[GS - 1], PH);
#pragma unroll
                        for (int r = 0; r < 1 + DIM; ++r) out[r * nlp] = PH[r];
                    } else {
                        double R[DIM + DIM * DIM], PH[MAX_ROWS];
#pragma unroll
                        for (int r = 0; r < DIM + DIM * DIM; ++r) R[r] = __ldg(tq + r * nl);
                        phys_rows<DIM>(1, R, g + DIM, g + DIM + DIM * DIM, g[GS - 1], PH);
#pragma unroll
                        for (int r = 0; r < DIM + DIM * DIM; ++r) out[r * nlp] = PH[r];
                    }
                }
            }
            __syncthreads();
            // Z rows of these points: the descriptors of a row are read once for all points of the round
            {
                const int bstr = nside * P.sbsz;
                for (int z = lane; z < P.nzd; z += TPE) {
                    const int4 zd = *reinterpret_cast<const int4*>(tZ + 4 * z);
                    const double* bq = bE + zd.z;
                    double sacc[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int k = zd.x; k < zd.y; ++k) {
                        const int2 en = *reinterpret_cast<const int2*>(tEnt + 2 * k);
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq)
                            if (qq < nqq) sacc[qq] = fma(dE[qq * P.nslots + en.x], bq[qq * bstr + en.y], sacc[qq]);
                    }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
                        if (qq < nqq) zE[qq * P.zsz + zd.w] = sacc[qq];
                }
            }
            __syncthreads();
            // tile updates: A[i0 + a][j0 + b] += B[row_g][i0 + a] * Z_g[j0 + b] over the segments g of the tile's pair
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                const int s1 = tS0[t] + tNS[t];
                for (int qq = 0; qq < nqq; ++qq) {
                    const double* bp = bE + qq * nside * P.sbsz + tB[t];
                    const double* zp = zE + qq * P.zsz + tJ[t];
                    for (int sg = tS0[t]; sg < s1; ++sg) {
                        const int2 so = *reinterpret_cast<const int2*>(tSeg + 2 * sg);
                        const double2 b01 = *reinterpret_cast<const double2*>(bp + so.x);
                        const double2 b23 = *reinterpret_cast<const double2*>(bp + so.x + 2);
                        const double2 z01 = *reinterpret_cast<const double2*>(zp + so.y);
                        const double2 z23 = *reinterpret_cast<const double2*>(zp + so.y + 2);
                        const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
                        const double zv[4] = {z01.x, z01.y, z23.x, z23.y};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[t][4 * a + b] = fma(bv[a], zv[b], acc[t][4 * a + b]);
                    }
                }
            }
        }
        // ---- deterministic path: the finished tiles go to the item-local buffer (16 doubles per tile, coalesced);
        //      ocmp_gather_add then sums, for every non-zero, its contributions in a fixed order
        if (P.abuf) {
            if (it_own < nitems) {
#pragma unroll
                for (int t = 0; t < MAXT; ++t) {
                    const int tile = lane + t * TPE;
                    if (tile >= P.ntiles) continue;
                    double2* dst = reinterpret_cast<double2*>(P.abuf + ((long long)it_own * P.ntiles + tile) * 16);
#pragma unroll
                    for (int k = 0; k < 8; ++k) dst[k] = make_double2(acc[t][2 * k], acc[t][2 * k + 1]);
                }
            }
            continue;
        }
        // ---- scatter-add through the element -> nnz map ------------------------------------------------------
        if (it_own < nitems) {
            const int c0 = sI[4 * e_own], c1 = sI[4 * e_own + 1];
#pragma unroll
            for (int t = 0; t < MAXT; ++t) {
                if (tNS[t] == 0) continue;
                const int st = tSide[t] & 1, su = (tSide[t] >> 1) & 1;
                const int ni = tRem[t] & 0xff, nj = tRem[t] >> 8;
                const int* map = (st == su) ? P.cell2nnz + (long long)(st ? c1 : c0) * n2
                                            : P.facet2nnz + ((long long)(item0 + it_own) * 2 + st) * n2;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (a < ni && b < nj)
                            atomicAdd(values + __ldg(map + (tRow[t] + a) * P.nloc + tCol[t] + b), acc[t][4 * a + b]);
            }
        }
    }
}

// linear forms: one thread per (item, side, local dof)
template <int DIM>
__global__ void __launch_bounds__(128) k_lin(const __grid_constant__ ocmp_contract_plan P, int item0, int nitems,
                                             const double* __restrict__ dbuf, double* __restrict__ vec) {
    constexpr int GS = GeoT<DIM>::GS;
    const long long total = (long long)nitems * P.nside * P.nloc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int it = (int)(gid / (P.nside * P.nloc)), rem = (int)(gid % (P.nside * P.nloc));
    const int s = rem / P.nloc, i = rem % P.nloc;
    const int id = P.items ? __ldg(P.items + item0 + it) : item0 + it;
    int c, lf = 0;
    if (P.kind == 0) c = id;
    else { c = __ldg(P.facet_cells + 2 * id + s); lf = __ldg(P.facet_local + 2 * id + s); }
    if (c < 0) return;
    int b = 0;
    while (b + 1 < P.nblk && i >= P.blk[6 * (b + 1) + 4]) ++b;
    const int* bd = P.blk + 6 * b;
    const int kind = bd[0], nl = bd[1], nr = bd[2], toff = bd[3], loff = bd[4];
    const int il = i - loff, sbk = s * P.nblk + b;
    // does any entry test against this side-block?
    const int nent = P.nent;
    bool any = false;
    for (int k = 0; k < nent; ++k) any |= ((P.ent[2 * k + 1] >> 8) == sbk);
    if (!any) return;
    const double* g = P.geo + (long long)c * GS;
    const long long dstride = (long long)nitems * P.nq;
    double acc = 0.0;
    for (int q = 0; q < P.nq; ++q) {
        const int lfq = (P.kind == 0 ? 0 : lf * P.nq) + q;
        double PH[MAX_ROWS];
        dof_phys_rows<DIM>(kind, nl, nr, P.tab + toff + (long long)lfq * nr * nl, il, g, PH);
        for (int k = 0; k < nent; ++k) {
            const int code = P.ent[2 * k + 1];
            if ((code >> 8) != sbk) continue;
            const int row = code & 0xff;
            double val = 0.0;
#pragma unroll
            for (int r = 0; r < MAX_ROWS; ++r)
                if (r == row) val = PH[r];
            acc = fma(__ldg(dbuf + (long long)P.ent[2 * k] * dstride + (long long)it * P.nq + q), val, acc);
        }
    }
    if (P.abuf) P.abuf[gid] = acc;             // deterministic path: local vectors, summed by ocmp_gather_add
    else atomicAdd(vec + __ldg(P.cell_dofs + (long long)c * P.nloc + i), acc);
}

// dst[seg_tgt[s]] += sum of src[order[k]], k in [seg_ptr[s], seg_ptr[s + 1]): every target has exactly one thread and
// a fixed summation order, so the result does not depend on the scheduling (unlike atomicAdd scatter)
__global__ void __launch_bounds__(256) k_gather_add(int nseg, const int* __restrict__ seg_ptr,
                                                    const int* __restrict__ seg_tgt, const int* __restrict__ order,
                                                    const double* __restrict__ src, double* __restrict__ dst) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int a = __ldg(seg_ptr + s), e = __ldg(seg_ptr + s + 1);
    double v = 0.0;
    for (int k = a; k < e; ++k) v += __ldg(src + __ldg(order + k));
    dst[__ldg(seg_tgt + s)] += v;
}

__global__ void __launch_bounds__(256) k_sum(const double* __restrict__ x, long long n, double* __restrict__ out) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += x[i];
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// ---- C ABI ------------------------------------------------------------------------------------------------------
extern "C" int ocmp_eval_coefficients(const ocmp_coef_plan* plan, int item0, int nitems, double* dbuf, void* stream) {
    if (nitems <= 0) return 0;
    if (plan->nfslots > MAX_FSLOTS) return ocmp_fail(-2, "too many field slots in one integral");
    if (plan->nreg > OCMP_MAX_REGS) return ocmp_fail(-2, "coefficient program needs too many registers");
    const long long total = (long long)nitems * plan->nq;
    const int threads = 128;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_COEF, st);
    if (plan->dim == 2) k_coef<2><<<blocks, threads, 0, st>>>(*plan, item0, nitems, dbuf);
    else if (plan->dim == 3) k_coef<3><<<blocks, threads, 0, st>>>(*plan, item0, nitems, dbuf);
    else return ocmp_fail(-1, "dim must be 2 or 3");
    return ocmp_check("ocmp_eval_coefficients");
}

static size_t contract_smem(const ocmp_contract_plan* p) {
    const int gs = p->dim + 2 * p->dim * p->dim + 1;
    size_t dbl = (size_t)p->eb * p->qb * ((size_t)p->nside * p->sbsz + p->zsz + p->nslots) +
                 (size_t)p->eb * p->nside * gs;
    dbl += dbl & 1;
    const size_t ints = 4 * (size_t)p->eb + 4 * (size_t)p->nside * p->nloc + 4 * (size_t)p->nzd + 2 * (size_t)p->nent +
                        2 * (size_t)p->nseg;
    return sizeof(double) * dbl + sizeof(int) * ints + 16;
}

template <int DIM, int MAXT>
static int launch_contract(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf, double* values,
                           cudaStream_t st) {
    const size_t smem = contract_smem(plan);
    if (smem > 220 * 1024) return ocmp_fail(-3, "contraction plan needs more than 220 KB of shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_contract<DIM, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int ngroups = (nitems + plan->eb - 1) / plan->eb;
    int sms = ocmp_sm_count();
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_contract<DIM, MAXT>, 256, smem);
    if (per_sm < 1) per_sm = 1;
    const int grid = ngroups < sms * per_sm ? ngroups : sms * per_sm;
    ProfScope ps(PROF_CONTRACT, st);
    k_contract<DIM, MAXT><<<grid, 256, smem, st>>>(*plan, item0, nitems, dbuf, values);
    return ocmp_check("ocmp_contract_matrix");
}

template <int DIM>
static int launch_contract_t(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf,
                             double* values, cudaStream_t st) {
    if (plan->maxt <= 1) return launch_contract<DIM, 1>(plan, item0, nitems, dbuf, values, st);
    if (plan->maxt == 2) return launch_contract<DIM, 2>(plan, item0, nitems, dbuf, values, st);
    return launch_contract<DIM, 4>(plan, item0, nitems, dbuf, values, st);
}

extern "C" int ocmp_contract_matrix(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf,
                                    double* values, void* stream) {
    if (nitems <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int eb = plan->eb;
    if (eb < 1 || eb > 16 || (eb & (eb - 1))) return ocmp_fail(-1, "eb must be a power of two <= 16");
    if (plan->qb < 1 || plan->qb > 4) return ocmp_fail(-1, "contraction plan: qb must be 1 .. 4");
    if (plan->ntiles > plan->maxt * (256 / eb) || plan->maxt > 4)
        return ocmp_fail(-1, "contraction plan: more tiles than the threads of an item can hold");
    if (plan->dim == 2) return launch_contract_t<2>(plan, item0, nitems, dbuf, values, st);
    if (plan->dim == 3) return launch_contract_t<3>(plan, item0, nitems, dbuf, values, st);
    return ocmp_fail(-1, "dim must be 2 or 3");
}

extern "C" int ocmp_contract_vector(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf,
                                    double* vec, void* stream) {
    if (nitems <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)nitems * plan->nside * plan->nloc;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    ProfScope ps(PROF_LIN, st);
    if (plan->dim == 2) k_lin<2><<<blocks, 128, 0, st>>>(*plan, item0, nitems, dbuf, vec);
    else if (plan->dim == 3) k_lin<3><<<blocks, 128, 0, st>>>(*plan, item0, nitems, dbuf, vec);
    else return ocmp_fail(-1, "dim must be 2 or 3");
    return ocmp_check("ocmp_contract_vector");
}

extern "C" int ocmp_gather_add(int nseg, const int* seg_ptr, const int* seg_tgt, const int* order, const double* src,
                               double* dst, void* stream) {
    if (nseg <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_CONTRACT, st);
    k_gather_add<<<(nseg + 255) / 256, 256, 0, st>>>(nseg, seg_ptr, seg_tgt, order, src, dst);
    return ocmp_check("ocmp_gather_add");
}

extern "C" int ocmp_sum(const double* x, long long n, double* out, void* stream) {
    if (n <= 0) return 0;
    long long blocks = (n + 255) / 256;
    const int cap = ocmp_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PROF_VEC, (cudaStream_t)stream);
    k_sum<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out);
    return ocmp_check("ocmp_sum");
}
