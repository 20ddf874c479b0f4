// gmg_stream2.cu -- instantiations and dispatch of k_rb_stream2 (see gmg_stream2.cuh)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "gmg_stream2.h"
#include "gmg_stream2.cuh"

namespace mgb {
namespace {

template <int S, bool EXACT, int MODE, bool PIN>
struct Inst {
    static cudaError_t occupancy(int *occ)
    {
        static int cached = 0;
        if (!cached) {
            constexpr int smem = S2Layout<S, MODE, PIN>::bytes;
            cudaError_t e = cudaFuncSetAttribute(k_rb_stream2<S, EXACT, MODE, PIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for a kernel that allocates tensor memory, although the
            // SM co-schedules CTAs as long as their TMEM columns, registers and shared memory fit: count them here
            cudaFuncAttributes fa;
            e = cudaFuncGetAttributes(&fa, k_rb_stream2<S, EXACT, MODE, PIN>);
            if (e != cudaSuccess) return e;
            int dev = 0, regs_sm = 65536, smem_sm = 233472;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
            cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
            const int regs_cta = ((fa.numRegs * 32 + 255) / 256) * 256 * (kS2NT / 32);
            const int smem_cta = smem + (int)fa.sharedSizeBytes + 1024;
            int o = std::min(regs_sm / regs_cta, smem_sm / smem_cta);
            if (kS2TmemB) o = std::min(o, 512 / (int)kS2TmemCols);
            o = std::min(o, s2_min_ctas<S>());           // __launch_bounds__ of the kernel
            cached = o > 0 ? o : 1;
            if (getenv("MGB_DEBUG")) fprintf(stderr, "k_rb_stream2<%d,%d,%d,%d>: smem %d B, occupancy %d\n", S, (int)EXACT, MODE, (int)PIN, smem, o);
        }
        *occ = cached;
        return cudaSuccess;
    }
    static cudaError_t launch(dim3 grid, cudaStream_t st, const Stream2Args &a)
    {
        constexpr int smem = S2Layout<S, MODE, PIN>::bytes;
        k_rb_stream2<S, EXACT, MODE, PIN><<<grid, kS2NT, smem, st>>>(a.g, a.in, a.rhs, a.out, a.rows_per_chunk, a.ucorr, a.aux, a.gc,
                                                                      a.restr, a.rscale);
        return cudaGetLastError();
    }
};

// the combinations the fused fast path launches (everything else stays on the first-generation kernel)
#define MGB_S2_LIST(F) \
    F(10, false, 1, true) F(10, false, 0, true) F(4, false, 3, false) F(4, false, 2, false)

}  // namespace

bool stream2_has(int S, bool exact, int mode, bool pin)
{
#define F(s, e, m, p) if (S == s && exact == e && mode == m && pin == p) return true;
    MGB_S2_LIST(F)
#undef F
    return false;
}

cudaError_t stream2_occupancy(int S, bool exact, int mode, bool pin, int *occ)
{
#define F(s, e, m, p) if (S == s && exact == e && mode == m && pin == p) return Inst<s, e, m, p>::occupancy(occ);
    MGB_S2_LIST(F)
#undef F
    return cudaErrorInvalidValue;
}

cudaError_t stream2_launch(int S, bool exact, int mode, bool pin, dim3 grid, cudaStream_t st, const Stream2Args &a)
{
#define F(s, e, m, p) if (S == s && exact == e && mode == m && pin == p) return Inst<s, e, m, p>::launch(grid, st, a);
    MGB_S2_LIST(F)
#undef F
    return cudaErrorInvalidValue;
}

int stream2_period(int S) { return (2 * S + 2 <= 12) ? 12 : 24; }

}  // namespace mgb
