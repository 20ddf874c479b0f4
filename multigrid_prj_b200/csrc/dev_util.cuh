// dev_util.cuh -- building blocks of the device-side SETUP code (AMG hierarchy, FEM assembly): scans, sorts and the
// expand-sort-compress construction of a CSR matrix from a list of (row, col, value) contributions.
//
// The solve phase (smoothers, residuals, transfers) is hand-written and lives in amg_kernels.cuh / gmg_kernels.cuh.
// The setup runs once per hierarchy; its scans / radix sorts / selections come from CUB (shipped with the CUDA
// toolkit), everything problem-specific (strength, splitting, interpolation, expansion of the triple product, the
// compression of duplicates in a fixed order) is written here.
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/mgb200.h"

int mgb_set_error(int code, const std::string &msg);   // gmg_solver.cu

#define DCK(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return mgb_set_error(MGB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace mgb {
namespace dev {

// MGB_TRACE_SETUP=1: wall-clock of the setup phases on stderr (each mark synchronises the device)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char *who;
    explicit Trace(const char *w) : on(std::getenv("MGB_TRACE_SETUP") != nullptr), t0(std::chrono::steady_clock::now()), who(w) {}
    void mark(const char *what, long long a = -1, long long b = -1)
    {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mgb setup] %-14s %-28s %9.3f ms", who, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        if (a >= 0) std::fprintf(stderr, "  %lld", a);
        if (b >= 0) std::fprintf(stderr, "  %lld", b);
        std::fprintf(stderr, "\n");
        t0 = std::chrono::steady_clock::now();
    }
};

// device array owned by a scope (setup temporaries)
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count)
    {
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        return cudaMalloc(&p, sizeof(T) * (count ? count : 1));
    }
    void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
    T *release() { T *q = p; p = nullptr; n = 0; return q; }
};

// out[i] = sum of in[0..i), i = 0..n  (out has n + 1 entries; in and out may not alias)
inline int exclusive_scan(const int *in, int *out, size_t n, cudaStream_t st)
{
    size_t bytes = 0;
    DCK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)(n + 1), st));
    DBuf<char> tmp;
    DCK(tmp.alloc(bytes));
    // the scan reads in[n] as well: callers allocate n + 1 input entries with in[n] = anything (its value only lands beyond out[n])
    DCK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)(n + 1), st));
    DCK(cudaStreamSynchronize(st));
    return MGB_OK;
}

inline int read_int(const int *d, int *h, cudaStream_t st)
{
    DCK(cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, st));
    DCK(cudaStreamSynchronize(st));
    return MGB_OK;
}

inline int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

// stable radix sort of (key, value) pairs on the low `key_bits` bits; results land in keys_out / vals_out
template <class K, class V>
inline int sort_pairs(const K *keys_in, K *keys_out, const V *vals_in, V *vals_out, size_t n, int key_bits, cudaStream_t st)
{
    if (n >= ((size_t)1 << 31)) return mgb_set_error(MGB_ERR_ARG, "setup: more than 2^31 items in one sort");
    size_t bytes = 0;
    DCK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, key_bits, st));
    DBuf<char> tmp;
    DCK(tmp.alloc(bytes));
    DCK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, key_bits, st));
    DCK(cudaStreamSynchronize(st));
    return MGB_OK;
}

// 64-bit sum of an int array (counts whose total may pass 2^31)
static __global__ void __launch_bounds__(256)
k_sum_int64(const int *__restrict__ v, int n, unsigned long long *out)
{
    unsigned long long acc = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += (unsigned long long)v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
inline int sum_int64(const int *v, int n, size_t *total, cudaStream_t st)
{
    DBuf<unsigned long long> d;
    DCK(d.alloc(1));
    DCK(cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), st));
    if (n > 0) k_sum_int64<<<std::min((n + 255) / 256, 1184), 256, 0, st>>>(v, n, d.p);
    DCK(cudaGetLastError());
    unsigned long long h = 0;
    DCK(cudaMemcpyAsync(&h, d.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    DCK(cudaStreamSynchronize(st));
    *total = (size_t)h;
    return MGB_OK;
}

// ---- expand-sort-compress -----------------------------------------------------------------------------------------
// key = row << 32 | col.  After a STABLE sort the contributions to one entry are adjacent and keep the order in which
// they were expanded, so adding them front to back gives a result that does not depend on the launch geometry.
__device__ __forceinline__ uint64_t esc_key(int row, int col) { return ((uint64_t)(uint32_t)row << 32) | (uint32_t)col; }

struct EscHead {
    const uint64_t *k;
    __device__ __forceinline__ bool operator()(const unsigned &p) const { return p == 0 || k[p] != k[p - 1]; }
};

static __global__ void __launch_bounds__(256)
k_esc_compress(const uint64_t *__restrict__ keys, const double *__restrict__ vals, unsigned m, const unsigned *__restrict__ start,
               int nnz, int *__restrict__ row_of, int *__restrict__ col, double *__restrict__ val)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    const unsigned p0 = start[t], p1 = (t + 1 < nnz) ? start[t + 1] : m;
    double s = 0.;
    for (unsigned p = p0; p < p1; ++p) s += vals[p];
    const uint64_t k = keys[p0];
    row_of[t] = (int)(k >> 32);
    col[t] = (int)(k & 0xffffffffu);
    val[t] = s;
}

struct EscNonZero {
    const double *v;
    __device__ __forceinline__ bool operator()(const unsigned &t) const { return v[t] != 0.0; }
};
static __global__ void __launch_bounds__(256)
k_esc_gather(const unsigned *__restrict__ keep, int n, const int *__restrict__ row_in, const int *__restrict__ col_in,
             const double *__restrict__ val_in, int *__restrict__ row_out, int *__restrict__ col_out, double *__restrict__ val_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned s = keep[t];
    row_out[t] = row_in[s]; col_out[t] = col_in[s]; val_out[t] = val_in[s];
}

// ptr[r] = first entry whose row is >= r (entries sorted by row), ptr[n_rows] = nnz
static __global__ void __launch_bounds__(256)
k_rows_to_ptr(const int *__restrict__ row_of, int nnz, int n_rows, int *__restrict__ ptr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nnz) return;
    const int prev = t == 0 ? -1 : row_of[t - 1];
    const int cur = t == nnz ? n_rows : row_of[t];
    for (int r = prev + 1; r <= cur; ++r) ptr[r] = t;
}

// (keys, vals)[0..m) -> CSR with sorted, duplicate-free rows.  keys/vals are consumed (used as sort buffers).
// ptr/col/val are cudaMalloc'ed here; *nnz_out entries.
inline int esc_to_csr(uint64_t *keys, double *vals, size_t m, int n_rows, int n_cols, int **ptr_out, int **col_out,
                      double **val_out, int *nnz_out, cudaStream_t st)
{
    *ptr_out = nullptr; *col_out = nullptr; *val_out = nullptr; *nnz_out = 0;
    DBuf<int> ptr;
    DCK(ptr.alloc((size_t)n_rows + 1));
    if (m == 0) {
        DCK(cudaMemsetAsync(ptr.p, 0, sizeof(int) * ((size_t)n_rows + 1), st));
        DCK(cudaStreamSynchronize(st));
        *ptr_out = ptr.release();
        DCK(cudaMalloc(col_out, sizeof(int)));
        DCK(cudaMalloc(val_out, sizeof(double)));
        return MGB_OK;
    }
    if (m >= ((size_t)1 << 31)) return mgb_set_error(MGB_ERR_ARG, "setup: more than 2^31 contributions in one product");
    DBuf<uint64_t> keys2;
    DBuf<double> vals2;
    DCK(keys2.alloc(m));
    DCK(vals2.alloc(m));
    const int key_bits = 32 + bits_for((uint64_t)(n_rows > 0 ? n_rows - 1 : 0));
    (void)n_cols;
    if (int rc = sort_pairs(keys, keys2.p, vals, vals2.p, m, key_bits, st)) return rc;
    // positions of the first contribution of every distinct key
    DBuf<unsigned> start;
    DBuf<int> d_count;
    DCK(start.alloc(m));
    DCK(d_count.alloc(1));
    {
        thrust::counting_iterator<unsigned> it(0u);
        EscHead pred{keys2.p};
        size_t bytes = 0;
        DCK(cub::DeviceSelect::If(nullptr, bytes, it, start.p, d_count.p, (int)m, pred, st));
        DBuf<char> tmp;
        DCK(tmp.alloc(bytes));
        DCK(cub::DeviceSelect::If(tmp.p, bytes, it, start.p, d_count.p, (int)m, pred, st));
        DCK(cudaStreamSynchronize(st));
    }
    int nnz = 0;
    if (int rc = read_int(d_count.p, &nnz, st)) return rc;
    DBuf<int> row_of, col;
    DBuf<double> val;
    DCK(row_of.alloc((size_t)nnz));
    DCK(col.alloc((size_t)nnz));
    DCK(val.alloc((size_t)nnz));
    k_esc_compress<<<(nnz + 255) / 256, 256, 0, st>>>(keys2.p, vals2.p, (unsigned)m, start.p, nnz, row_of.p, col.p, val.p);
    DCK(cudaGetLastError());
    DCK(cudaStreamSynchronize(st));
    keys2.free(); vals2.free();
    // entries that cancelled to exactly 0 are dropped, as CSRMatrix::copy_from does (AMG/src/CSRMatrix.cpp:13-14)
    {
        thrust::counting_iterator<unsigned> it(0u);
        EscNonZero pred{val.p};
        size_t bytes = 0;
        DCK(cub::DeviceSelect::If(nullptr, bytes, it, start.p, d_count.p, nnz, pred, st));
        DBuf<char> tmp;
        DCK(tmp.alloc(bytes));
        DCK(cub::DeviceSelect::If(tmp.p, bytes, it, start.p, d_count.p, nnz, pred, st));
        int kept = 0;
        if (int rc = read_int(d_count.p, &kept, st)) return rc;
        if (kept != nnz) {
            DBuf<int> row2, col2;
            DBuf<double> val2;
            DCK(row2.alloc((size_t)kept)); DCK(col2.alloc((size_t)kept)); DCK(val2.alloc((size_t)kept));
            if (kept) k_esc_gather<<<(kept + 255) / 256, 256, 0, st>>>(start.p, kept, row_of.p, col.p, val.p, row2.p, col2.p, val2.p);
            DCK(cudaGetLastError());
            DCK(cudaStreamSynchronize(st));
            std::swap(row_of.p, row2.p); std::swap(col.p, col2.p); std::swap(val.p, val2.p);
            nnz = kept;
        }
    }
    k_rows_to_ptr<<<(nnz + 1 + 255) / 256, 256, 0, st>>>(row_of.p, nnz, n_rows, ptr.p);
    DCK(cudaGetLastError());
    DCK(cudaStreamSynchronize(st));
    *ptr_out = ptr.release(); *col_out = col.release(); *val_out = val.release(); *nnz_out = nnz;
    return MGB_OK;
}

}  // namespace dev
}  // namespace mgb
