// gmg_stream2.h -- host-side interface of the second-generation streaming red-black kernel (gmg_stream2.cuh).
// The kernels are instantiated in gmg_stream2.cu; gmg_solver.cu picks them through these three functions.
#pragma once
#include <cuda_runtime.h>
#include "gmg_common.cuh"

namespace mgb {

struct Stream2Args {
    LevelGeom g, gc;
    const double *in, *rhs;
    double *out, *ucorr, *aux;
    int rows_per_chunk, restr;
    double rscale;
};

// true when k_rb_stream2<S, EXACT, MODE, PIN> is compiled into the library
bool stream2_has(int S, bool exact, int mode, bool pin);
// raises the dynamic shared-memory limit (once) and returns the resident CTAs per SM in *occ
cudaError_t stream2_occupancy(int S, bool exact, int mode, bool pin, int *occ);
cudaError_t stream2_launch(int S, bool exact, int mode, bool pin, dim3 grid, cudaStream_t st, const Stream2Args &a);
// row steps of one unrolled period of the steady loop (rows_per_chunk is chosen against it)
int stream2_period(int S);

}  // namespace mgb
