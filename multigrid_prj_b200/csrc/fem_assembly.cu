// fem_assembly.cu -- operator construction on the device: P1 finite-element assembly of
// the Poisson problem on a triangle mesh, and the synthetic unstructured triangulation of BASELINE config 5.
//
// Replaces (reference, relative to AMG/): the assembly loop of src/main.cpp:34-117 with LinearFE::set_dofs
// (include/FEM.hpp:174-258) and the problem functions of src/Utilities.cpp:3-28.  The element formulas are restated
// in the reference's evaluation order with unfused operations, every contribution is expanded in the order the
// reference adds it (element by element; per element i, j, q), and a STABLE sort groups them -- so with
// `exact_order` the assembled matrix is bit-identical to the reference's (tests/test_fem_assembly_gpu.py, mesh1.msh);
// the right-hand side agrees to the last bits of sin / cos / sqrt (device libm vs glibc).
// Unknowns are the interior nodes in order of appearance (src/FEM.cpp:291-303).
#include "dev_util.cuh"

#include <vector>

struct mgb_system {
    int device = 0;
    int n = 0, nnz = 0;
    int *ptr = nullptr, *col = nullptr;
    double *val = nullptr, *rhs = nullptr;
    cudaStream_t st = nullptr;
};

namespace {

using mgb::dev::DBuf;

// src/Utilities.cpp:3-28
__device__ __forceinline__ double boundary_function(double x, double y) { return sin(5 * sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)))); }
__device__ __forceinline__ double forcing_term(double x, double y)
{
    const double r = sqrt(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
    return __dmul_rn(-5., __dsub_rn(__ddiv_rn(cos(5 * r), r), __dmul_rn(5., sin(5 * r))));
}

struct Element {
    double area;                 // element_area of FEM.hpp:178-186 (twice the geometric area)
    double gx[3], gy[3];         // gradients of the three basis functions
    double cx[3], cy[3], c0[3];  // coefficients of the basis functions as the reference stores them
};

// LinearFE::set_dofs (FEM.hpp:174-258), same operations in the same order, never contracted
__device__ __forceinline__ void element_setup(const double (&x)[3], const double (&y)[3], Element &e)
{
    e.area = fabs(__dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(x[1], y[2]), __dmul_rn(x[2], y[1])),
                                      __dsub_rn(__dmul_rn(y[0], x[2]), __dmul_rn(x[0], y[2]))),
                            __dsub_rn(__dmul_rn(x[0], y[1]), __dmul_rn(y[0], x[1]))));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        const double vj0 = __dsub_rn(x[j], x[i]), vj1 = __dsub_rn(y[j], y[i]), vj2 = -1.0;
        const double vk0 = __dsub_rn(x[k], x[i]), vk1 = __dsub_rn(y[k], y[i]), vk2 = -1.0;
        const double n0 = __dsub_rn(__dmul_rn(vj1, vk2), __dmul_rn(vk1, vj2));
        const double n1 = __dsub_rn(__dmul_rn(vj2, vk0), __dmul_rn(vk2, vj0));
        const double n2 = __dsub_rn(__dmul_rn(vj0, vk1), __dmul_rn(vk0, vj1));
        e.gx[i] = __ddiv_rn(-n0, n2);
        e.gy[i] = __ddiv_rn(-n1, n2);
        e.cx[i] = __ddiv_rn(n0, n2);
        e.cy[i] = __ddiv_rn(n1, n2);
        e.c0[i] = __dsub_rn(__dsub_rn(1., __dmul_rn(e.cx[i], x[i])), __dmul_rn(e.cy[i], y[i]));
    }
}
__device__ __forceinline__ double basis(const Element &e, int i, double px, double py)
{
    return __dsub_rn(2., __dadd_rn(__dadd_rn(e.c0[i], __dmul_rn(e.cx[i], px)), __dmul_rn(e.cy[i], py)));
}
// alpha * (grad_i . grad_j) * w   (alpha == 1, src/Utilities.cpp:25-28)
__device__ __forceinline__ double stiffness_term(const Element &e, int i, int j, double w)
{
    return __dmul_rn(__dmul_rn(1.0, __dadd_rn(__dmul_rn(e.gx[i], e.gx[j]), __dmul_rn(e.gy[i], e.gy[j]))), w);
}

// Contributions of one element.  COUNT: number of matrix / rhs contributions; otherwise they are written at mo / ro.
// q_terms = 3: one contribution per quadrature point, as the reference adds them (exact_order); 1: the three equal
// terms are added up first.
template <bool COUNT>
__device__ __forceinline__ void element_contributions(const double *__restrict__ X, const double *__restrict__ Y,
                                                      const unsigned char *__restrict__ bnd, const int *__restrict__ dof,
                                                      const int (&v)[3], int q_terms, int &nm, int &nr, size_t mo, size_t ro,
                                                      uint64_t *mkeys, double *mvals, uint64_t *rkeys, double *rvals)
{
    bool b[3] = {bnd[v[0]] != 0, bnd[v[1]] != 0, bnd[v[2]] != 0};
    const int n_int = (!b[0]) + (!b[1]) + (!b[2]), n_bnd = 3 - n_int;
    if (COUNT) {
        nm = n_int * n_int * q_terms;
        nr = n_int * (q_terms + n_bnd * q_terms);
        return;
    }
    if (n_int == 0) return;
    double x[3] = {X[v[0]], X[v[1]], X[v[2]]}, y[3] = {Y[v[0]], Y[v[1]], Y[v[2]]};
    Element e;
    element_setup(x, y, e);
    const double w = __ddiv_rn(e.area, 3.0);
    for (int i = 0; i < 3; ++i) {
        if (b[i]) continue;
        const int di = dof[v[i]];
        for (int j = 0; j < 3; ++j) {
            if (b[j]) continue;
            const double t = stiffness_term(e, i, j, w);
            const uint64_t key = mgb::dev::esc_key(di, dof[v[j]]);
            if (q_terms == 3) { for (int q = 0; q < 3; ++q) { mkeys[mo] = key; mvals[mo] = t; ++mo; } }
            else { mkeys[mo] = key; mvals[mo] = __dadd_rn(__dadd_rn(t, t), t); ++mo; }
        }
        const double f = forcing_term(x[i], y[i]);
        const uint64_t rkey = mgb::dev::esc_key(di, 0);
        if (q_terms == 3) {
            for (int q = 0; q < 3; ++q) { rkeys[ro] = rkey; rvals[ro] = __dmul_rn(__dmul_rn(f, basis(e, i, x[q], y[q])), w); ++ro; }
        } else {
            double s = 0.;
            for (int q = 0; q < 3; ++q) s = __dadd_rn(s, __dmul_rn(__dmul_rn(f, basis(e, i, x[q], y[q])), w));
            rkeys[ro] = rkey; rvals[ro] = s; ++ro;
        }
    }
    if (n_bnd == 0) return;
    // Dirichlet lifting F - B g (src/main.cpp:88-113), after all the load terms of the element
    // NB the reference interleaves nothing here: within one row the order is load terms, then lifting terms by j, q
    for (int i = 0; i < 3; ++i) {
        if (b[i]) continue;
        const uint64_t rkey = mgb::dev::esc_key(dof[v[i]], 0);
        for (int j = 0; j < 3; ++j) {
            if (!b[j]) continue;
            const double g = boundary_function(x[j], y[j]);
            const double t = -__dmul_rn(__dmul_rn(__dmul_rn(g, 1.0), __dadd_rn(__dmul_rn(e.gx[i], e.gx[j]), __dmul_rn(e.gy[i], e.gy[j]))), w);
            if (q_terms == 3) { for (int q = 0; q < 3; ++q) { rkeys[ro] = rkey; rvals[ro] = t; ++ro; } }
            else { rkeys[ro] = rkey; rvals[ro] = __dadd_rn(__dadd_rn(t, t), t); ++ro; }
        }
    }
}

__global__ void __launch_bounds__(256)
k_fem_count(const double *X, const double *Y, const unsigned char *bnd, const int *dof, const int *tri, int n_tri, int q_terms,
            int *__restrict__ cm, int *__restrict__ cr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tri) return;
    int nm = 0, nr = 0;
    if (t < n_tri) {
        const int v[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
        element_contributions<true>(X, Y, bnd, dof, v, q_terms, nm, nr, 0, 0, nullptr, nullptr, nullptr, nullptr);
    }
    cm[t] = nm; cr[t] = nr;
}

__global__ void __launch_bounds__(128)
k_fem_expand(const double *X, const double *Y, const unsigned char *bnd, const int *dof, const int *tri, int n_tri, int q_terms,
             const int *__restrict__ om, const int *__restrict__ orr, uint64_t *mkeys, double *mvals, uint64_t *rkeys, double *rvals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const int v[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
    int nm, nr;
    element_contributions<false>(X, Y, bnd, dof, v, q_terms, nm, nr, (size_t)om[t], (size_t)orr[t], mkeys, mvals, rkeys, rvals);
}

__global__ void __launch_bounds__(256)
k_interior_flag(const unsigned char *__restrict__ bnd, int n, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flag[i] = (i < n && !bnd[i]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_scatter_rhs(const int *__restrict__ ptr, const double *__restrict__ val, int n, double *__restrict__ rhs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rhs[i] = ptr[i + 1] > ptr[i] ? val[ptr[i]] : 0.0;
}

// ---- synthetic unstructured triangulation (BASELINE config 5) ---------------------------------
// side x side lattice on [0,2]^2, interior nodes jittered by <= 0.2 h, every cell split by one of its two diagonals;
// jitter and diagonal come from a counter-based hash of (seed, index), so host and device generate the same mesh.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double unit01(uint64_t seed, uint64_t stream, uint64_t idx)
{
    return (double)(mix64(mix64(seed + stream) ^ idx) >> 11) * (1.0 / 9007199254740992.0);       // 53 bits -> [0, 1)
}
// one rounding per operation on the host AND on the device (nvcc would contract a * b + c into an FMA)
__host__ __device__ __forceinline__ double mul_r(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__host__ __device__ __forceinline__ double add_r(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// coordinates of lattice node p = i * side + j: (j h, i h), interior nodes moved by at most 0.2 h in each direction
__host__ __device__ __forceinline__ void synth_node(uint64_t side, uint64_t seed, uint64_t p, double &x, double &y, bool &interior)
{
    const uint64_t i = p / side, j = p % side;
    const double h = 2.0 / (double)(side - 1);
    interior = i > 0 && i < side - 1 && j > 0 && j < side - 1;
    x = mul_r((double)j, h); y = mul_r((double)i, h);
    if (interior) {
        x = add_r(x, mul_r(add_r(mul_r(unit01(seed, 1, p), 0.4), -0.2), h));
        y = add_r(y, mul_r(add_r(mul_r(unit01(seed, 2, p), 0.4), -0.2), h));
    }
}

__global__ void __launch_bounds__(256)
k_synth_nodes(int side, uint64_t seed, double *__restrict__ X, double *__restrict__ Y, unsigned char *__restrict__ bnd)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)side * side) return;
    double x, y;
    bool interior;
    synth_node((uint64_t)side, seed, (uint64_t)p, x, y, interior);
    X[p] = x; Y[p] = y; bnd[p] = interior ? 0 : 1;
}

__global__ void __launch_bounds__(256)
k_synth_triangles(int side, uint64_t seed, int *__restrict__ tri)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_cells = (size_t)(side - 1) * (side - 1);
    if (c >= n_cells) return;
    const int i = (int)(c / (side - 1)), j = (int)(c % (side - 1));
    const int a00 = i * side + j, a01 = a00 + 1, a10 = a00 + side, a11 = a10 + 1;
    const bool flip = (mix64(mix64(seed + 3) ^ c) >> 63) != 0;
    int *t = tri + 6 * c;           // vertices ascending, as the reference's reader orders them (FEM.cpp:162-170)
    if (flip) { t[0] = a00; t[1] = a01; t[2] = a10; t[3] = a01; t[4] = a10; t[5] = a11; }
    else { t[0] = a00; t[1] = a01; t[2] = a11; t[3] = a00; t[4] = a10; t[5] = a11; }
}

// nodes / triangles / flags on the device -> system
int assemble_device(int n_nodes, const double *X, const double *Y, const unsigned char *bnd, int n_tri, const int *tri,
                    int exact_order, mgb_system *S)
{
    cudaStream_t st = S->st;
    const int q_terms = exact_order ? 3 : 1;
    DBuf<int> flag, dof;
    DCK(flag.alloc((size_t)n_nodes + 1)); DCK(dof.alloc((size_t)n_nodes + 2));
    k_interior_flag<<<(n_nodes + 1 + 255) / 256, 256, 0, st>>>(bnd, n_nodes, flag.p);
    DCK(cudaGetLastError());
    if (int rc = mgb::dev::exclusive_scan(flag.p, dof.p, (size_t)n_nodes, st)) return rc;
    int n = 0;
    if (int rc = mgb::dev::read_int(dof.p + n_nodes, &n, st)) return rc;
    flag.free();
    DBuf<int> cm, cr, om, orr;
    DCK(cm.alloc((size_t)n_tri + 1)); DCK(cr.alloc((size_t)n_tri + 1)); DCK(om.alloc((size_t)n_tri + 2)); DCK(orr.alloc((size_t)n_tri + 2));
    k_fem_count<<<(n_tri + 1 + 255) / 256, 256, 0, st>>>(X, Y, bnd, dof.p, tri, n_tri, q_terms, cm.p, cr.p);
    DCK(cudaGetLastError());
    size_t tm = 0, tr = 0;
    if (int rc = mgb::dev::sum_int64(cm.p, n_tri, &tm, st)) return rc;
    if (int rc = mgb::dev::sum_int64(cr.p, n_tri, &tr, st)) return rc;
    if (tm >= ((size_t)1 << 31) || tr >= ((size_t)1 << 31)) return mgb_set_error(MGB_ERR_ARG, "mesh too large for one assembly pass");
    if (int rc = mgb::dev::exclusive_scan(cm.p, om.p, (size_t)n_tri, st)) return rc;
    if (int rc = mgb::dev::exclusive_scan(cr.p, orr.p, (size_t)n_tri, st)) return rc;
    cm.free(); cr.free();
    DBuf<uint64_t> mkeys, rkeys;
    DBuf<double> mvals, rvals;
    DCK(mkeys.alloc(tm)); DCK(mvals.alloc(tm)); DCK(rkeys.alloc(tr)); DCK(rvals.alloc(tr));
    if (n_tri) k_fem_expand<<<(n_tri + 127) / 128, 128, 0, st>>>(X, Y, bnd, dof.p, tri, n_tri, q_terms, om.p, orr.p, mkeys.p, mvals.p, rkeys.p, rvals.p);
    DCK(cudaGetLastError());
    om.free(); orr.free(); dof.free();
    S->n = n;
    if (int rc = mgb::dev::esc_to_csr(mkeys.p, mvals.p, tm, n, n, &S->ptr, &S->col, &S->val, &S->nnz, st)) return rc;
    mkeys.free(); mvals.free();
    int *rp = nullptr, *rc_ = nullptr, rn = 0;
    double *rv = nullptr;
    int rc = mgb::dev::esc_to_csr(rkeys.p, rvals.p, tr, n, 1, &rp, &rc_, &rv, &rn, st);
    if (!rc) {
        cudaError_t e = cudaMalloc(&S->rhs, sizeof(double) * (size_t)std::max(n, 1));
        if (e != cudaSuccess) rc = mgb_set_error(MGB_ERR_CUDA, cudaGetErrorString(e));
        else if (n) k_scatter_rhs<<<(n + 255) / 256, 256, 0, st>>>(rp, rv, n, S->rhs);
        cudaStreamSynchronize(st);
    }
    cudaFree(rp); cudaFree(rc_); cudaFree(rv);
    return rc;
}

int new_system(int device, mgb_system **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return mgb_set_error(MGB_ERR_CUDA, "no CUDA device: libmgb200 has no CPU fallback");
    }
    DCK(cudaSetDevice(device));
    mgb_system *S = new mgb_system();
    S->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete S; return mgb_set_error(MGB_ERR_CUDA, cudaGetErrorString(e)); }
    *out = S;
    return MGB_OK;
}

}  // namespace

extern "C" {

void mgb_system_destroy(mgb_system_t S)
{
    if (!S) return;
    cudaSetDevice(S->device);
    cudaFree(S->ptr); cudaFree(S->col); cudaFree(S->val); cudaFree(S->rhs);
    if (S->st) cudaStreamDestroy(S->st);
    delete S;
}

int mgb_fem_assemble_p1(size_t n_nodes, const double *x, const double *y, const unsigned char *on_boundary, size_t n_tri,
                        const int64_t *tri, int exact_order, int device, mgb_system_t *out)
{
    if (!x || !y || !on_boundary || !tri || !out) return mgb_set_error(MGB_ERR_ARG, "null argument");
    if (n_nodes == 0 || n_nodes >= ((size_t)1 << 31) || n_tri >= ((size_t)1 << 29)) return mgb_set_error(MGB_ERR_ARG, "mesh too large for int32 indices");
    *out = nullptr;
    mgb_system *S = nullptr;
    if (int rc = new_system(device, &S)) return rc;
    struct Guard { mgb_system *s; ~Guard() { if (s) mgb_system_destroy(s); } } guard{S};
    std::vector<int> t32(3 * n_tri);
    for (size_t k = 0; k < 3 * n_tri; ++k) {
        if (tri[k] < 0 || tri[k] >= (int64_t)n_nodes) return mgb_set_error(MGB_ERR_ARG, "triangle vertex out of range");
        t32[k] = (int)tri[k];
    }
    DBuf<double> X, Y;
    DBuf<unsigned char> B;
    DBuf<int> T;
    DCK(X.alloc(n_nodes)); DCK(Y.alloc(n_nodes)); DCK(B.alloc(n_nodes)); DCK(T.alloc(3 * n_tri));
    DCK(cudaMemcpy(X.p, x, sizeof(double) * n_nodes, cudaMemcpyHostToDevice));
    DCK(cudaMemcpy(Y.p, y, sizeof(double) * n_nodes, cudaMemcpyHostToDevice));
    DCK(cudaMemcpy(B.p, on_boundary, n_nodes, cudaMemcpyHostToDevice));
    if (n_tri) DCK(cudaMemcpy(T.p, t32.data(), sizeof(int) * 3 * n_tri, cudaMemcpyHostToDevice));
    if (int rc = assemble_device((int)n_nodes, X.p, Y.p, B.p, (int)n_tri, T.p, exact_order, S)) return rc;
    guard.s = nullptr;
    *out = S;
    return MGB_OK;
}

int mgb_fem_synthetic(size_t side, uint64_t seed, int device, mgb_system_t *out)
{
    if (!out || side < 3 || side > 30000) return mgb_set_error(MGB_ERR_ARG, "3 <= side <= 30000");
    *out = nullptr;
    mgb_system *S = nullptr;
    if (int rc = new_system(device, &S)) return rc;
    struct Guard { mgb_system *s; ~Guard() { if (s) mgb_system_destroy(s); } } guard{S};
    const size_t n_nodes = side * side, n_cells = (side - 1) * (side - 1);
    DBuf<double> X, Y;
    DBuf<unsigned char> B;
    DBuf<int> T;
    DCK(X.alloc(n_nodes)); DCK(Y.alloc(n_nodes)); DCK(B.alloc(n_nodes)); DCK(T.alloc(6 * n_cells));
    k_synth_nodes<<<(unsigned)((n_nodes + 255) / 256), 256, 0, S->st>>>((int)side, seed, X.p, Y.p, B.p);
    k_synth_triangles<<<(unsigned)((n_cells + 255) / 256), 256, 0, S->st>>>((int)side, seed, T.p);
    DCK(cudaGetLastError());
    if (int rc = assemble_device((int)n_nodes, X.p, Y.p, B.p, (int)(2 * n_cells), T.p, 0, S)) return rc;
    guard.s = nullptr;
    *out = S;
    return MGB_OK;
}

// the mesh mgb_fem_synthetic assembles, for callers that want to check it on the host (x, y: side*side; tri: 6*(side-1)^2)
int mgb_fem_synthetic_mesh(size_t side, uint64_t seed, double *x, double *y, unsigned char *on_boundary, int64_t *tri)
{
    if (side < 3 || !x || !y || !on_boundary || !tri) return mgb_set_error(MGB_ERR_ARG, "bad argument");
    for (size_t p = 0; p < side * side; ++p) {
        bool interior;
        synth_node((uint64_t)side, seed, (uint64_t)p, x[p], y[p], interior);
        on_boundary[p] = interior ? 0 : 1;
    }
    for (size_t c = 0; c < (side - 1) * (side - 1); ++c) {
        const size_t i = c / (side - 1), j = c % (side - 1);
        const int64_t a00 = (int64_t)(i * side + j), a01 = a00 + 1, a10 = a00 + (int64_t)side, a11 = a10 + 1;
        const bool flip = (mix64(mix64(seed + 3) ^ c) >> 63) != 0;
        int64_t *t = tri + 6 * c;
        if (flip) { t[0] = a00; t[1] = a01; t[2] = a10; t[3] = a01; t[4] = a10; t[5] = a11; }
        else { t[0] = a00; t[1] = a01; t[2] = a11; t[3] = a00; t[4] = a10; t[5] = a11; }
    }
    return MGB_OK;
}

int mgb_system_info(mgb_system_t S, size_t *n, size_t *nnz)
{
    if (!S) return mgb_set_error(MGB_ERR_ARG, "null system");
    if (n) *n = (size_t)S->n;
    if (nnz) *nnz = (size_t)S->nnz;
    return MGB_OK;
}

int mgb_system_get(mgb_system_t S, int64_t *ptr, int64_t *col, double *val, double *rhs)
{
    if (!S) return mgb_set_error(MGB_ERR_ARG, "null system");
    DCK(cudaSetDevice(S->device));
    if (ptr) {
        std::vector<int> p((size_t)S->n + 1);
        DCK(cudaMemcpy(p.data(), S->ptr, sizeof(int) * p.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < p.size(); ++i) ptr[i] = p[i];
    }
    if (col && S->nnz) {
        std::vector<int> c((size_t)S->nnz);
        DCK(cudaMemcpy(c.data(), S->col, sizeof(int) * c.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < c.size(); ++i) col[i] = c[i];
    }
    if (val && S->nnz) DCK(cudaMemcpy(val, S->val, sizeof(double) * (size_t)S->nnz, cudaMemcpyDeviceToHost));
    if (rhs && S->n) DCK(cudaMemcpy(rhs, S->rhs, sizeof(double) * (size_t)S->n, cudaMemcpyDeviceToHost));
    return MGB_OK;
}

// used by amg_solver.cu (mgb_amg_create_from_system): raw device views
int mgb_system_device_view(mgb_system_t S, int *device, int *n, int *nnz, const int **ptr, const int **col, const double **val, const double **rhs)
{
    if (!S) return mgb_set_error(MGB_ERR_ARG, "null system");
    *device = S->device; *n = S->n; *nnz = S->nnz; *ptr = S->ptr; *col = S->col; *val = S->val; *rhs = S->rhs;
    return MGB_OK;
}

}  // extern "C"
