/* opencmp_b200 — C ABI of the B200 backend for OpenCMP's assembly + linear-solve hot path.
 *
 * The reference has no FFI of its own: its hot path calls NGSolve's Python API. Each entry point
 * below names the NGSolve call (and the OpenCMP call site, path:line under the reference tree) it stands in for.
 * All pointers are raw device pointers unless marked "host"; sizes are element counts; every function returns
 * 0 on success and a negative code otherwise (ocmp_last_error() gives the text). Nothing throws across the ABI.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef OPENCMP_B200_H
#define OPENCMP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OCMP_MAX_FVEC 8      /* distinct DOF vectors a single integral may read (wind, u^n, phi, masks ...) */
#define OCMP_MAX_REGS 48     /* registers of the coefficient bytecode machine (ir.py MAX_REGS) */

/* ---- plans: flat descriptors of one lowered integral (built by opencmp_b200/backend.py from ir.Integral) ------ */

/* Coefficient program of an integral: evaluated at every (item, quadrature point). */
typedef struct {
    int dim;                 /* 2 | 3 */
    int kind;                /* 0 cell, 1 interior facet, 2 boundary facet */
    int nq;                  /* quadrature points per item */
    int nout;                /* output slots D_k */
    int ninstr, nreg;
    int nfgroups, nfslots;
    const int* items;        /* item -> cell / facet id, NULL = identity */
    const double* geo;       /* per cell: x0[dim], J[dim*dim] (J[i*dim+a] = dx_i/dxi_a), Jinv[dim*dim], det */
    const int* facet_cells;  /* (nfacets, 2) */
    const int* facet_local;  /* (nfacets, 2) */
    const double* qpts;      /* cell: (nq, dim); facet: (nfc, nq, dim) reference coordinates */
    const double* qw;        /* (nq) */
    const double* fref;      /* facets: per local facet outward reference normal (dim) then tangents ((dim-1)*dim) */
    const int* code;         /* (ninstr, 4): op | dst << 8, a, b, c — opcode table in ir.py */
    const double* consts;
    const double* params;    /* run-time scalars (t, dt, ...), refreshed per call */
    const int* fgroup;       /* (nfgroups, 8): vec, dofarr, side, kind, nloc, nrows, tab_off, dof_off ; see below */
    const int* fgroup2;      /* (nfgroups, 2): dof_stride, dof_add */
    const int* fslot;        /* (nfslots, 2): group, physical row */
    const double* ftab;      /* reference tables of the field blocks */
    const double* fvec[OCMP_MAX_FVEC];
    const int* fdof[OCMP_MAX_FVEC];
} ocmp_coef_plan;

/* Contraction plan: B^T D B per item, scatter through the element -> nnz map. */
typedef struct {
    int dim, kind, nq, nside;
    int nblk;                /* blocks per side */
    int nloc;                /* local dofs per side */
    int sbsz;                /* doubles of one side's physical table at one quadrature point (rows padded to x4) */
    int zsz;                 /* doubles of the Z rows of one item (every segment padded to a multiple of 4) */
    int nact;                /* active (structurally non-zero) local entries per item */
    int nslots;              /* D slots (== coef plan nout) */
    int eb;                  /* items per 256-thread CTA */
    int ntiles;              /* 4 x 4 accumulator tiles covering the active block pairs of one item */
    int maxt;                /* tiles per thread: 1 | 2 | 4, ntiles <= maxt * 256 / eb */
    int nzd, nent, nseg;
    int qb;                  /* quadrature points staged in shared memory per barrier round */
    const int* items;
    const double* geo;
    const int* facet_cells;
    const int* facet_local;
    const int* blk;          /* (nblk, 6): kind, nloc, nrows, tab_off, loc_off, sb_off — vector assembly */
    const double* tab;       /* reference tables of the form's blocks at this rule */
    const int* zdesc;        /* (nzd, 4): entry k0, k1, sB base (side*sbsz + sb_off + j), z index */
    const int* ent;          /* matrices (nent, 2): slot, trial row * padded row stride; vectors: slot, side-block << 8 | row */
    const int* dofdesc;      /* (nside*nloc, 4): kind | nrows << 8 | padded nloc << 16, nloc of the block, tab_off + il,
                                sb_off + il */
    const int* tiles;        /* (ntiles, 8): sB offset of the test rows (side*sbsz + sb_off + i0), j0, first segment,
                                segments, test side | trial side << 1, row i0 and column j0 of the local matrix (within
                                one side), valid rows | valid columns << 8 */
    const int* seg;          /* (nseg, 2): test row * padded ni, z offset of the segment */
    const int* cell2nnz;     /* (ncells, nloc*nloc) */
    const int* facet2nnz;    /* (n interior facets, 2, nloc*nloc) indexed by item */
    const int* cell_dofs;    /* (ncells, nloc) — vector assembly */
    double* abuf;            /* deterministic assembly (NULL = atomicAdd scatter): matrices — nitems x ntiles x 16
                                doubles, the finished 4 x 4 tiles of every item; vectors — nitems x nside x nloc local
                                vector entries. ocmp_gather_add then adds them into the CSR values / the vector */
} ocmp_contract_plan;

/* ---- assembly: stands in for BilinearForm.Assemble / LinearForm.Assemble (reference
 *      opencmp/solvers/base_solver.py:368-377) and ngs.Integrate (opencmp/helpers/error.py:66-77) ------------- */
int ocmp_eval_coefficients(const ocmp_coef_plan* plan, int item0, int nitems, double* dbuf, void* stream);
int ocmp_contract_matrix(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf,
                         double* values, void* stream);
int ocmp_contract_vector(const ocmp_contract_plan* plan, int item0, int nitems, const double* dbuf,
                         double* vec, void* stream);
int ocmp_sum(const double* x, long long n, double* out_accumulate, void* stream);
/* Second phase of the deterministic assembly: dst[seg_tgt[s]] += sum_k src[order[k]], k in [seg_ptr[s], seg_ptr[s+1]).
 * The lists are built once per plan and chunk by the host (stable sort of the element -> nnz map): every target has
 * one writer and a fixed summation order, so two assemblies of the same form are bit-identical — on one GPU and
 * between the ranks of an element-partitioned run. */
int ocmp_gather_add(int nseg, const int* seg_ptr, const int* seg_tgt, const int* order, const double* src,
                    double* dst, void* stream);

/* ---- sparse / dense vector kernels: stand in for `a.mat * x`, BaseVector arithmetic and InnerProduct
 *      (reference opencmp/models/base_model.py:918-922) ------------------------------------------------------- */
int ocmp_spmv(int nrows, const int* rowptr, const int* colidx, const double* vals, const double* x, double* y,
              void* stream);
/* Column-index compression for vector-valued spaces whose components are numbered one after the other, `shift` apart
 * (VectorH1: u_x, u_y[, u_z]): ocmp_spmv_runs marks every row whose columns start with nc runs of equal length L
 * holding the same scalar columns shifted by 0, shift, (2 shift) — runlen[row] = L, else 0. Products over such rows
 * read only the first run's indices (4 index bytes per nc values). ocmp_spmv_compressed = ocmp_spmv with the marks. */
int ocmp_spmv_runs(int nrows, const int* rowptr, const int* colidx, int shift, int nc, int* runlen,
                   int* grouped_bad_dev, void* stream);
/* grouped_bad_dev (optional, one int): left 0 when, in addition, the nc component rows of every node (row, row +
 * shift, ...; row < shift) have identical column lists. Then `grouped` products let one lane group compute the nc
 * rows of a node together: x values and indices are fetched once per nc x nc matrix values, which divides the L2
 * gather traffic of a scattered CG numbering by nc. */
int ocmp_spmv_compressed(int nrows, const int* rowptr, const int* colidx, const double* vals, const int* runlen,
                         int shift, int nc, int grouped, const double* x, double* y, void* stream);
int ocmp_dot(long long n, const double* x, const double* y, double* out, void* stream);
int ocmp_axpby(long long n, double a, const double* x, double b, double* y, void* stream);   /* y = a x + b y */
int ocmp_masked_assign(long long n, double* dst, const double* src, const double* inv, const double* mask,
                       void* stream);
int ocmp_to_f32(long long n, const double* src, float* dst, void* stream);   /* dst[i] = (float) src[i] */
/* Batched level-1 building blocks of the Krylov drivers, also used by the stationary nonlinear mixers that stand in for
 * reference opencmp/solvers/nonlinear_mixing.py:19-141 (Anderson / DiagBroyden / linear mixing on the DOF vector):
 * out_dev[j] = <V_j, w>, j < k, V_j = V + j*ld;   w += sum_j coef_dev[j] V_j. */
int ocmp_mdot(long long n, const double* V, long long ld, int k, const double* w, double* out_dev, void* stream);
int ocmp_maxpy(long long n, const double* V, long long ld, int k, const double* coef_dev, double* w, void* stream);

/* ---- preconditioners: stand in for ngs.Preconditioner(a, type) + .Update() (reference
 *      opencmp/models/base_model.py:365-383, opencmp/solvers/base_solver.py:711-719) -------------------------- */
/* point Jacobi on the free dofs: dinv[i] = free[i] && diag != 0 ? 1/diag : 0 */
int ocmp_jacobi_setup(int nrows, const int* diagpos, const double* vals, const double* freemask, double* dinv,
                      void* stream);
/* additive Schwarz over dof patches (patch_dofs: npatch x bs, padded with -1): gather the dense blocks from the CSR
 * matrix and invert them in shared memory (batched Gauss-Jordan with partial pivoting) */
int ocmp_asm_setup(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                   const double* vals, const double* freemask, double* inv_blocks, const int* positions,
                   void* stream);
/* one-off: CSR positions of every (padded to a multiple of 16) patch entry, npatch x NP x NP int32, -1 = not in the
 * pattern; NULL positions in ocmp_asm_setup selects the slower pivoted shared-memory kernel */
int ocmp_patch_positions(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                         int* positions, void* stream);
/* Application z = sum_p scatter(A_p^-1 r[dofs_p]): the stored inverses are streamed once (bulk copies into a shared-
 * memory ring), the patch products land in the patch-local scratch `ybuf` (npatch x bs doubles) and a per-dof gather
 * through the incidence list sums them in a fixed order — no atomics, bit-reproducible. inc_ptr (n + 1) / inc_idx:
 * for every dof the positions p * bs + i of its valid patch entries, ascending (built once per patch table by the
 * host). The patch stride bs must keep the stored columns 16-byte aligned (even for FP64, % 4 for FP32, % 8 for
 * bfloat16), bs <= 256. */
int ocmp_asm_apply(int npatch, int bs, const int* patch_dofs, const double* inv_blocks, const int* inc_ptr,
                   const int* inc_idx, double* ybuf, const double* r, double* z, long long n, void* stream);
/* The same smoother with the patch inverses STORED in FP32 (inverted and applied in FP64 arithmetic): its application
 * is bound by streaming the inverses from HBM, so this halves the bytes of the dominant kernel; as a preconditioner
 * inside FP64 GMRES it leaves iteration counts and solutions unchanged. bs must be a multiple of 4. */
int ocmp_asm_setup_f32(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                       const double* vals, const double* freemask, float* inv_blocks, const int* positions,
                       void* stream);
int ocmp_asm_apply_f32(int npatch, int bs, const int* patch_dofs, const float* inv_blocks, const int* inc_ptr,
                       const int* inc_idx, double* ybuf, const double* r, double* z, long long n, void* stream);
/* ... and in bfloat16 (a quarter of the FP64 bytes; bs a multiple of 8). Iteration counts on the CPU restatement:
 * unchanged in 2-D, +-2 in 3-D (profiles/r1_solver_convergence.md). */
int ocmp_asm_setup_bf16(int npatch, int bs, const int* patch_dofs, const int* rowptr, const int* colidx,
                        const double* vals, const double* freemask, unsigned short* inv_blocks, const int* positions,
                        void* stream);
int ocmp_asm_apply_bf16(int npatch, int bs, const int* patch_dofs, const unsigned short* inv_blocks,
                        const int* inc_ptr, const int* inc_idx, double* ybuf, const double* r, double* z, long long n,
                        void* stream);

/* ---- Krylov: stand in for ngs.solvers.CG / GMRes / PreconditionedRichardson and for mat.Inverse applied to a
 *      residual (reference opencmp/models/base_model.py:886-947) ---------------------------------------------- */
struct ocmp_mg_level;
struct ocmp_band_lu;
typedef struct ocmp_system {
    int nrows;
    const int* rowptr;
    const int* colidx;
    const double* vals;
    const double* freemask;  /* 1.0 free / 0.0 constrained, NULL = all free */
    int pre_kind;            /* 0 none, 1 Jacobi (dinv), 2 additive Schwarz patches, 3 geometric multigrid V-cycle,
                                4 explicit inverse stored as CSR (coarsest multigrid level), 5 band LU (direct) */
    const double* dinv;
    int npatch, bs;
    const int* patch_dofs;
    const void* inv_blocks;  /* npatch x bs x bs, transposed per patch; double / float / bfloat16, see inv_storage */
    const double* patch_weight; /* optional per-dof weight applied after the additive patch solves (restricted / averaged AS) */
    int nlevels;             /* pre_kind 3: levels[0] coarsest ... levels[nlevels-1] = this system */
    const struct ocmp_mg_level* levels;
    const int* inv_rowptr;   /* pre_kind 4 */
    const int* inv_colidx;
    const double* inv_vals;
    /* element-partitioned runs; all zero / NULL on a single GPU. Vectors are local (owned + ghost)
     * and kept consistent: ghost entries equal the owner's value. Plan fields hold ocmp_halo_plan handle + 1. */
    const double* owned;     /* 1.0 on the entries this rank owns: dot products (followed by one ncclAllReduce) and
                                restriction only see those; NULL = not distributed */
    int halo_fwd;            /* refresh ghost entries after an SpMV / after the coarsest-level solve (0: none) */
    int halo_sum;            /* sum the neighbours' partial patch corrections after the smoother (0: none) */
    int inv_storage;         /* type of inv_blocks: 0 double, 1 float (ocmp_asm_setup_f32), 2 bfloat16 (.._bf16) */
    const float* vals32;     /* optional FP32 copy of vals (ocmp_to_f32): used for the operator applications inside
                                the multigrid cycle only — the Krylov method around it always applies `vals` */
    const void* apply_fn;    /* optional matrix-free operator (an ocmp_apply_fn, see below): when set, the Krylov method
                                applies it instead of the CSR product — NGSolve's BilinearForm(nonassemble=True).mat
                                handed to solvers.CG / GMRes; rowptr / colidx / vals may then be NULL unless the
                                preconditioner reads them. NULL = stored CSR operator */
    void* apply_ctx;         /* first argument of apply_fn */
    const struct ocmp_band_lu* direct; /* pre_kind 5: factorised free-free block (ocmp_band_factor) applied as the
                                preconditioner — ngs.Preconditioner(a, 'direct') */
    const int* patch_inc_ptr;   /* pre_kind 2 / 3: incidence list of the patch table (see ocmp_asm_apply) */
    const int* patch_inc_idx;
    double* patch_ybuf;         /* npatch x bs doubles of scratch for the patch products */
    const int* spmv_rows;       /* element-partitioned: the rows this rank owns (ascending). Operator applications and
                                   residuals only compute those — the ghost rows of a local matrix are incomplete and
                                   their entries are overwritten by the halo exchange that follows. NULL = all rows */
    int n_spmv_rows;
    const int* run_len;         /* optional marks of ocmp_spmv_runs for the CSR pattern of this system (NULL = none) */
    int run_shift, run_nc;
    int run_grouped;            /* 0, or run_shift when the grouped product is valid for the full row range */
    int n_spmv_groups;          /* with spmv_rows: 0, or the number of listed rows below run_shift when the list holds,
                                   in this order, those rows, their nc - 1 shifted copies, then the other rows */
} ocmp_system;
/* A band LU as ocmp_band_fill / ocmp_band_factor leave it, plus the permutation and a work vector of n doubles. */
typedef struct ocmp_band_lu {
    int n, kl, ku, ubw;
    const double* ab;
    const int* ipiv;
    const int* perm;         /* nrows of the system: position in band order, -1 = constrained */
    double* rhs;             /* n doubles of scratch */
} ocmp_band_lu;
/* y = A x on device vectors of length nrows, enqueued on `stream`; returns 0 on success */
typedef int (*ocmp_apply_fn)(void* ctx, const double* x, double* y, void* stream);

/* One multigrid level: operator + smoother (sys), transfer from the next coarser level, work space.
 * Stands in for ngs.Preconditioner(a, 'multigrid') (reference opencmp/models/base_model.py:365-383). */
typedef struct ocmp_mg_level {
    ocmp_system sys;
    int ncoarse;
    const int* p_rowptr; const int* p_colidx; const double* p_vals;   /* prolongation (nrows x ncoarse), CSR */
    const int* r_rowptr; const int* r_colidx; const double* r_vals;   /* restriction = its transpose, CSR */
    double* work;            /* 4 * nrows doubles: x, b, r, t */
    int nu;                  /* pre- and post-smoothing sweeps */
    double omega;            /* smoother damping */
    int restrict_sum;        /* plan + 1 on the COARSE level: sum the partial restricted residuals (0: none) */
    int handover;            /* plan + 1 on the COARSE level: a replicated coarse level hands the owners' correction up
                                to a distributed one (redundant solves differ by round-off) (0: none) */
} ocmp_mg_level;

/* kind: 0 CG, 1 GMRES(restart), 2 Richardson, 3 MINRES (symmetric operator, SPD preconditioner). x holds the initial guess (and Dirichlet values) on entry.
 * Stops when the preconditioned residual norm drops below tol * initial. iters / resid are host outputs. */
int ocmp_krylov(const ocmp_system* sys, int kind, const double* b, double* x, double tol, int maxit, int restart,
                double damp, double* work, long long work_len, int* iters, double* resid, void* stream);
long long ocmp_krylov_work_len(int nrows, int kind, int restart);
/* Relative preconditioned residual after every iteration of the last ocmp_krylov call (CG, GMRES); returns the number
 * of iterations recorded and copies at most `cap` of them. NGSolve's solvers print these with printrates=True
 * (reference base_model.py:925-943 passes printrates=self.verbose). */
int ocmp_krylov_history(double* out_host, int cap);

/* ---- sparse direct solve: stands in for `a.mat.Inverse(freedofs=..., inverse="umfpack"|"pardiso")` applied to a
 *      residual, the reference's default linear_solver = direct (reference opencmp/models/base_model.py:908-922).
 * The free-free block is permuted (perm[i] = position of dof i in a bandwidth-reducing order, -1 = constrained; built
 * once per pattern by the host) and held as a LAPACK-style band array `ab` of ocmp_band_len(n, kl, ku) doubles
 * (column j: rows j-kl-ku .. j+kl, leading dimension 2 kl + ku + 1). ocmp_band_factor = dgbtrf (partial pivoting,
 * one persistent cooperative kernel), info_dev[0] = 1-based first zero pivot or 0, info_dev[1] = super-diagonals of U;
 * ocmp_band_solve = dgbtrs on one right-hand side in place. gather / scatter move between dof order and band order. */
long long ocmp_band_len(int n, int kl, int ku);
int ocmp_band_fill(int nrows, const int* rowptr, const int* colidx, const double* vals, const int* perm, int n,
                   int kl, int ku, double* ab, void* stream);
int ocmp_band_factor(int n, int kl, int ku, double* ab, int* ipiv, int* info_dev, void* stream);
int ocmp_band_solve(int n, int kl, int ku, int ubw, const double* ab, const int* ipiv, double* b, void* stream);
int ocmp_band_gather(int nrows, const int* perm, const double* r, double* b, void* stream);
int ocmp_band_scatter(int nrows, const int* perm, const double* x, double* out, int accumulate, void* stream);

/* ---- diffuse-interface phase-field generation: the voxel pipeline of the reference's pre-processing,
 *      reference opencmp/diffuse_interface/interface.py:31-57 (get_binary_2d; ray tracing mesh_helpers.py:268-302) and
 *      :137-180 (get_phi: erosion -> border -> exact Euclidean distance transform -> erf profile). Arrays are C-ordered
 *      (n0, n1[, n2]) like the reference's NumPy arrays; n_k = N_k + 1 grid nodes. */
int ocmp_dim_raytrace_2d(int n0, int n1, double scale0, double scale1, double offset0, double offset1, int N0, int N1,
                         int npoly, const double* poly_xy, double* binary, void* stream);
/* fg[v] = 0 on the border voxels of the shape (binary - 3^d erosion), 1 elsewhere: the argument of edt.edt() */
int ocmp_dim_border(int dim, int n0, int n1, int n2, const double* binary, unsigned char* fg, void* stream);
/* exact Euclidean distance (in voxels, FP32 like the edt package) of every voxel with fg != 0 to the nearest voxel with
 * fg == 0; work_a / work_b: n0*n1*n2 int64 each */
int ocmp_dim_edt(int dim, int n0, int n1, int n2, const unsigned char* fg, long long* work_a, long long* work_b,
                 float* dist, void* stream);
/* Rigid-body motion of a node field on the structured grid (reference opencmp/helpers/ngsolve_.py:212-296, evaluated
 * every time step for moving diffuse interfaces): out[node] = orig(inv_rotation * x_node), multilinear interpolation,
 * 1 where the pre-image leaves the box. scale / offset / inv_rotation (dim x dim, row-major) are host arrays. */
int ocmp_dim_rigid_motion(int dim, int n0, int n1, int n2, const double* scale_host, const double* offset_host,
                          const double* inv_rotation_host, const double* orig, double* out, void* stream);
/* phi = (erf(dist * h / lmbda) * (2 binary - 1) + 1) / 2 */
int ocmp_dim_phi(long long n, const float* dist, const double* binary, double h, double lmbda, double* phi,
                 void* stream);

/* ---- instrumentation: per-category device time from CUDA events recorded around every launch on its own stream.
 * categories: 0 spmv, 1 asm_apply, 2 coefficient eval, 3 matrix contraction, 4 vector contraction, 5 multi-dot,
 * 6 multi-axpy, 7 other vector kernels, 8 preconditioner setup, 9 SpMVs inside the multigrid cycle, 10 halo exchange */
void ocmp_profile_enable(int on);
void ocmp_profile_reset(void);
int ocmp_profile_read(int category, long long* count, double* ms);
double ocmp_profile_bytes(int category);   /* algorithmic bytes of the category's launches (asm_apply only) */
long long ocmp_launch_count(void);

/* ---- element-partitioned runs (no reference counterpart): NCCL halo exchange in one call -----------
 * ocmp_comm_unique_id: rank 0 creates the 128-byte NCCL id (broadcast it with any transport), ocmp_comm_init: all ranks.
 * ocmp_halo_plan: per neighbour `peers[i]` the counts of values sent / received and the concatenated device index lists
 * into the local vector; returns a plan handle (>= 0). ocmp_halo_run: pack, grouped ncclSend/ncclRecv on `stream`,
 * unpack — add = 0 overwrites (ghost refresh), add = 1 accumulates (partial contributions). */
int ocmp_comm_unique_id(void* out128_host);
int ocmp_comm_init(const void* id128_host, int nranks, int rank);
int ocmp_halo_plan(int nnbr, const int* peers_host, const int* send_cnt_host, const int* recv_cnt_host,
                   const int* send_idx_dev, const int* recv_idx_dev, double* sendbuf_dev, double* recvbuf_dev);
int ocmp_halo_run(int plan, double* x, int add, void* stream);
int ocmp_allreduce_sum(double* buf_dev, int n, void* stream);

const char* ocmp_last_error(void);
int ocmp_version(void);

#ifdef __cplusplus
}
#endif
#endif
