// amg_dev.h -- device-resident containers of the AMG solve phase shared by amg_solver.cu (host driver + C ABI) and
// amg_setup.cu (device-side setup: strength, C/F splitting, interpolation, Galerkin product, SELL copies).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "amg_types.h"

int mgb_set_error(int code, const std::string &msg);   // gmg_solver.cu

namespace mgb {
namespace amg {

// rank r owns the entries [n r / R, n (r+1) / R) of a length-n index space (rows of an operator, entries of a vector)
struct Block { int r0 = 0, r1 = 0; int size() const { return r1 - r0; } };
inline Block block_of(int n, int n_ranks, int rank)
{
    Block b;
    b.r0 = (int)((long long)n * rank / n_ranks);
    b.r1 = (int)((long long)n * (rank + 1) / n_ranks);
    return b;
}
struct DevCsr {
    int n_rows = 0, n_cols = 0, nnz = 0;
    int *ptr = nullptr, *col = nullptr;
    double *val = nullptr;
    mgb::CsrDev view() const { return mgb::CsrDev{n_rows, n_cols, nnz, ptr, col, val}; }
    void release() { cudaFree(ptr); cudaFree(col); cudaFree(val); ptr = col = nullptr; val = nullptr; }
};

struct Schedule {          // rows grouped into independent sets (wavefronts or colours)
    int n_groups = 0;
    std::vector<int> h_ptr, h_group;
    int *d_ptr = nullptr, *d_rows = nullptr;
    void release() { cudaFree(d_ptr); cudaFree(d_rows); d_ptr = d_rows = nullptr; }
};

struct SellCopy {              // colour-sorted SELL-32 copy of A for the fast kernels (amg_kernels.cuh)
    int n_slots = 0;
    std::vector<int> colour_slot_ptr;      // first slot of each colour (+ end)
    int *d_colour_slot_ptr = nullptr;      // the same on the device (cooperative whole-sweep kernel)
    int coop_blocks = 0;                   // grid of the cooperative kernel: co-resident, at most the largest colour
    int *slice_ptr = nullptr, *col = nullptr, *row_of_slot = nullptr;
    double *val = nullptr, *diag_s = nullptr, *b_s = nullptr;
    size_t stored = 0;                     // entries incl. padding
    mgb::SellDev view() const { return mgb::SellDev{n_slots, slice_ptr, col, val, row_of_slot, diag_s, b_s}; }
    void release() { cudaFree(slice_ptr); cudaFree(col); cudaFree(row_of_slot); cudaFree(val); cudaFree(diag_s); cudaFree(b_s); cudaFree(d_colour_slot_ptr); }
};


// ---- device-side setup (amg_setup.cu); every function returns MGB_OK or records the error text ---------------------------
int dev_diagonals(const DevCsr &A, double *diag, double *dl1, cudaStream_t st);
int dev_coarsen(const DevCsr &A, double eps, unsigned seed, DevCsr &P, DevCsr &R, DevCsr &Ac, cudaStream_t st, int *rounds_out,
                int **cf_out /* may be null: the C/F state (1 coarse, 0 fine), cudaMalloc'ed, owned by the caller */);
int dev_build_sell(const DevCsr &M, const int *list, int n_slots, bool skip_diag, const double *diag, const double *rhs,
                   bool slot_vectors, SellCopy &S, cudaStream_t st);
int dev_build_sell_range(const DevCsr &M, Block rows, bool skip_diag, SellCopy &S, cudaStream_t st);
int dev_group_schedule(const DevCsr &A, const int *group, int n_groups, Block own, Schedule &Sch, SellCopy *sell,
                       const double *diag, const double *rhs, cudaStream_t st);

}  // namespace amg
}  // namespace mgb
