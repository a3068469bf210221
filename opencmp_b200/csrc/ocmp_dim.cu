// Diffuse-interface phase-field generation on the device: the voxel pipeline of the reference's
// pre-processing — reference opencmp/diffuse_interface/interface.py:31-57 (get_binary_2d: ray tracing of every grid
// node against the boundary polygon, mesh_helpers.py:268-302), :159-180 (get_phi: erosion -> border -> exact Euclidean
// distance transform (the third-party `edt` package) -> erf profile). One-off per run, except with rigid-body motion
// where it is repeated every time step.
//
//   k_raytrace_2d   one thread per grid node, the reference's crossing test edge by edge (including its carried
//                   intersection abscissa on horizontal edges), same coordinate arithmetic -> identical masks
//   k_border        3^d erosion (outside the array counts as background) and border = 1 - (binary - erosion)
//   k_edt_pass      exact squared distance transform, one separable pass per axis: out[v] = min_k in[line(v), k] +
//                   (pos - k)^2 in 64-bit integers (exact for any array that fits the device); lines are short
//                   (<= a few hundred nodes), so the O(n) scan per voxel is cheaper than building lower envelopes
//   k_phi           dt = sqrt(d2) in FP32 like `edt`, phi = (erf(dt h / lambda) (2 binary - 1) + 1) / 2
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/opencmp_b200.h"
#include "ocmp_common.cuh"

namespace {
constexpr long long EDT_INF = 1LL << 60;

__global__ void __launch_bounds__(256) k_raytrace_2d(int n0, int n1, double s0, double s1, double o0, double o1,
                                                     int N0, int N1, int npoly, const double* __restrict__ poly,
                                                     double* __restrict__ binary) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n0 * n1) return;
    const int i = (int)(gid / n1), j = (int)(gid % n1);
    const double x = i * s0 / N0 - o0;              // interface.py:52-53, same operation order
    const double y = j * s1 / N1 - o1;
    bool inside = false;
    double xints = 0.0;
    double p1x = poly[0], p1y = poly[1];
    for (int k = 0; k <= npoly; ++k) {
        const int kk = k % npoly;
        const double p2x = poly[2 * kk], p2y = poly[2 * kk + 1];
        if (y > fmin(p1y, p2y) && y <= fmax(p1y, p2y) && x <= fmax(p1x, p2x)) {
            if (p1y != p2y) xints = (y - p1y) * (p2x - p1x) / (p2y - p1y) + p1x;
            if (p1x == p2x || x <= xints) inside = !inside;
        }
        p1x = p2x; p1y = p2y;
    }
    binary[gid] = inside ? 1.0 : 0.0;
}

// fg = 0 on the border voxels (inside the shape, but with a 3^d neighbour outside it or outside the array), else 1
__global__ void __launch_bounds__(256) k_border(int dim, int n0, int n1, int n2, const double* __restrict__ binary,
                                                unsigned char* __restrict__ fg) {
    const long long total = (long long)n0 * n1 * n2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int k = (int)(gid % n2), j = (int)((gid / n2) % n1), i = (int)(gid / ((long long)n1 * n2));
    const bool in = binary[gid] != 0.0;
    bool eroded = in;
    if (in) {
        const int dk = dim == 3 ? 1 : 0;
        for (int a = -1; a <= 1 && eroded; ++a)
            for (int b = -1; b <= 1 && eroded; ++b)
                for (int c = -dk; c <= dk && eroded; ++c) {
                    const int ii = i + a, jj = j + b, kk = k + c;
                    if (ii < 0 || ii >= n0 || jj < 0 || jj >= n1 || kk < 0 || kk >= n2) { eroded = false; break; }
                    if (binary[((long long)ii * n1 + jj) * n2 + kk] == 0.0) eroded = false;
                }
    }
    fg[gid] = (in && !eroded) ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_edt_init(long long total, const unsigned char* __restrict__ fg,
                                                  long long* __restrict__ d2) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) d2[gid] = fg[gid] ? EDT_INF : 0;
}

// one pass along the axis with `len` entries `stride` apart: out[v] = min_k in[base + k * stride] + (pos - k)^2
__global__ void __launch_bounds__(256) k_edt_pass(long long total, int len, long long stride,
                                                  const long long* __restrict__ in, long long* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int pos = (int)((gid / stride) % len);
    const long long base = gid - (long long)pos * stride;
    long long best = EDT_INF;
    for (int k = 0; k < len; ++k) {
        const long long v = __ldg(in + base + (long long)k * stride);
        if (v < EDT_INF) {
            const long long d = (long long)(pos - k) * (pos - k) + v;
            best = d < best ? d : best;
        }
    }
    out[gid] = best;
}

__global__ void __launch_bounds__(256) k_edt_sqrt(long long total, const long long* __restrict__ d2,
                                                  float* __restrict__ dist) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) dist[gid] = d2[gid] >= EDT_INF ? INFINITY : sqrtf((float)d2[gid]);
}

__global__ void __launch_bounds__(256) k_phi(long long total, const float* __restrict__ dist,
                                             const double* __restrict__ binary, float h, float lmbda,
                                             double* __restrict__ phi) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const float dt = dist[gid] * h;                  // interface.py:168: FP32 array scaled in place
    const double e = (double)erff(dt / lmbda);       // :173-174: FP32 erf, promoted by the FP64 mask it multiplies
    const double b = binary[gid];
    phi[gid] = (e * b + e * (b - 1.0) + 1.0) / 2.0;  // :173-178
}

// Rigid-body motion of a node field (reference helpers/ngsolve_.py:212-296, the per-time-step triple Python loop of
// moving diffuse interfaces): out[node] = orig(R^-1 x_node) by multilinear interpolation on the structured grid when the
// pre-image lies inside the grid's box, 1 otherwise. Node (i, j[, k]) sits at x_a = -offset_a + scale_a * idx_a / N_a.
struct RigidArgs {
    double scale[3], offset[3], invR[9];
    int n[3];
};
__global__ void __launch_bounds__(256) k_rigid_motion(int dim, RigidArgs A, const double* __restrict__ orig,
                                                      double* __restrict__ out) {
    const long long total = (long long)A.n[0] * A.n[1] * A.n[2];
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    int idx[3];
    idx[2] = (int)(gid % A.n[2]);
    idx[1] = (int)((gid / A.n[2]) % A.n[1]);
    idx[0] = (int)(gid / ((long long)A.n[1] * A.n[2]));
    double xn[3] = {0.0, 0.0, 0.0}, xo[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < dim; ++a) xn[a] = -A.offset[a] + A.scale[a] * idx[a] / (A.n[a] - 1);
    for (int a = 0; a < dim; ++a) {
        double sacc = 0.0;
        for (int b = 0; b < dim; ++b) sacc += A.invR[a * dim + b] * xn[b];
        xo[a] = sacc;
    }
    bool inside = true;
    for (int a = 0; a < dim; ++a) inside = inside && xo[a] >= -A.offset[a] && xo[a] <= A.scale[a] - A.offset[a];
    double val = 1.0;
    if (inside) {
        int c[3] = {0, 0, 0};
        double f[3] = {0.0, 0.0, 0.0};
        for (int a = 0; a < dim; ++a) {
            const int N = A.n[a] - 1;
            const double u = (xo[a] + A.offset[a]) / A.scale[a] * N;
            int i0 = (int)floor(u);
            i0 = i0 < 0 ? 0 : (i0 > N - 1 ? N - 1 : i0);
            c[a] = i0;
            f[a] = u - i0;
        }
        val = 0.0;
        const int nc = 1 << dim;
        for (int corner = 0; corner < nc; ++corner) {
            double w = 1.0;
            long long off = 0;
            for (int a = 0; a < 3; ++a) {
                const int bit = a < dim ? (corner >> a) & 1 : 0;
                if (a < dim) w *= bit ? f[a] : 1.0 - f[a];
                off = off * A.n[a] + (c[a] + bit);
            }
            val += w * orig[off];
        }
    }
    out[gid] = val;
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }
}  // namespace

extern "C" {

int ocmp_dim_raytrace_2d(int n0, int n1, double scale0, double scale1, double offset0, double offset1, int N0, int N1,
                         int npoly, const double* poly_xy, double* binary, void* stream) {
    if (n0 <= 0 || n1 <= 0) return 0;
    if (npoly < 1) return ocmp_fail(-1, "ocmp_dim_raytrace_2d: empty polygon");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    k_raytrace_2d<<<blocks_for((long long)n0 * n1), 256, 0, st>>>(n0, n1, scale0, scale1, offset0, offset1, N0, N1,
                                                                  npoly, poly_xy, binary);
    return ocmp_check("ocmp_dim_raytrace_2d");
}

int ocmp_dim_border(int dim, int n0, int n1, int n2, const double* binary, unsigned char* fg, void* stream) {
    if (dim != 2 && dim != 3) return ocmp_fail(-1, "ocmp_dim_border: dim must be 2 or 3");
    if (dim == 2) n2 = 1;
    const long long total = (long long)n0 * n1 * n2;
    if (total <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    k_border<<<blocks_for(total), 256, 0, st>>>(dim, n0, n1, n2, binary, fg);
    return ocmp_check("ocmp_dim_border");
}

int ocmp_dim_edt(int dim, int n0, int n1, int n2, const unsigned char* fg, long long* work_a, long long* work_b,
                 float* dist, void* stream) {
    if (dim != 2 && dim != 3) return ocmp_fail(-1, "ocmp_dim_edt: dim must be 2 or 3");
    if (dim == 2) n2 = 1;
    const long long total = (long long)n0 * n1 * n2;
    if (total <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    const unsigned nb = blocks_for(total);
    k_edt_init<<<nb, 256, 0, st>>>(total, fg, work_a);
    long long* cur = work_a;
    long long* nxt = work_b;
    const int lens[3] = {n0, n1, n2};
    const long long strides[3] = {(long long)n1 * n2, (long long)n2, 1};
    for (int ax = 2; ax >= 0; --ax) {
        if (lens[ax] == 1) continue;
        k_edt_pass<<<nb, 256, 0, st>>>(total, lens[ax], strides[ax], cur, nxt);
        long long* t = cur; cur = nxt; nxt = t;
    }
    k_edt_sqrt<<<nb, 256, 0, st>>>(total, cur, dist);
    return ocmp_check("ocmp_dim_edt");
}

int ocmp_dim_rigid_motion(int dim, int n0, int n1, int n2, const double* scale_host, const double* offset_host,
                          const double* inv_rotation_host, const double* orig, double* out, void* stream) {
    if (dim != 2 && dim != 3) return ocmp_fail(-1, "ocmp_dim_rigid_motion: dim must be 2 or 3");
    RigidArgs A;
    A.n[0] = n0; A.n[1] = n1; A.n[2] = dim == 3 ? n2 : 1;
    for (int a = 0; a < 3; ++a) { A.scale[a] = a < dim ? scale_host[a] : 1.0; A.offset[a] = a < dim ? offset_host[a] : 0.0; }
    for (int k = 0; k < 9; ++k) A.invR[k] = k < dim * dim ? inv_rotation_host[k] : 0.0;
    const long long total = (long long)A.n[0] * A.n[1] * A.n[2];
    if (total <= 0) return 0;
    for (int a = 0; a < dim; ++a)
        if (A.n[a] < 2) return ocmp_fail(-1, "ocmp_dim_rigid_motion: at least two nodes per direction");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    k_rigid_motion<<<blocks_for(total), 256, 0, st>>>(dim, A, orig, out);
    return ocmp_check("ocmp_dim_rigid_motion");
}

int ocmp_dim_phi(long long n, const float* dist, const double* binary, double h, double lmbda, double* phi,
                 void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_SETUP, st);
    k_phi<<<blocks_for(n), 256, 0, st>>>(n, dist, binary, (float)h, (float)lmbda, phi);
    return ocmp_check("ocmp_dim_phi");
}

}  // extern "C"
