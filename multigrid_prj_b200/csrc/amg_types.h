// amg_types.h -- plain views of the device-resident AMG operators (shared by the kernels and the setup code)
#pragma once

namespace mgb {

struct CsrDev {
    int n_rows, n_cols, nnz;
    const int *ptr, *col;
    const double *val;
};

// sliced-ELLPACK (SELL-32) copy of an operator: see amg_kernels.cuh ("fast path")
struct SellDev {
    int n_slots;                 // rows including padding (multiple of 32)
    const int *slice_ptr;        // [n_slots/32 + 1] offsets into col/val
    const int *col;              // column of each stored entry (padding: 0)
    const double *val;           // value (padding: 0.0)
    const int *row_of_slot;      // original row of a slot, -1 for padding slots
    const double *diag_s, *b_s;  // diagonal and right-hand side in slot order
};

}  // namespace mgb
