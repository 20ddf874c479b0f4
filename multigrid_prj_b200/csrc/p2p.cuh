// p2p.cuh -- direct peer-memory exchange between the GPUs of one box (one process per GPU).
//
// NCCL's grouped ncclSend/ncclRecv costs 60-100 us per group on 8 B200 (profiles/r02_slab_trace_8gpu_nccl.log), which
// is a third of a slab iteration.  The hot exchanges therefore go through NVLink peer memory directly:
//   * every rank allocates ONE pool (all level arrays + a small header) and exports it with cudaIpcGetMemHandle; the
//     handles travel once by ncclAllGather and every rank maps the pools of the others (cudaIpcOpenMemHandle);
//   * an exchange is ONE kernel that copies a list of segments from local memory into the peers' pools with 16-byte
//     stores over NVLink; the last CTA to finish bumps this rank's sequence number of the channel and publishes it in
//     the header of every peer it signals (st.release.sys after __threadfence_system);
//   * the consumer side is a one-warp kernel that spins (ld.acquire.sys) until the flags of the expected sources reach
//     its own sequence number of the channel.  Both counters live in device memory, so a CUDA graph that contains the
//     pair can be replayed any number of times.
// The wait kernel gives up after ~20 s (globaltimer) and raises header.error instead of hanging the GPU.
// Write-after-read safety is the caller's business: the slab schedule alternates two exchange points between the same
// peers, so a buffer is overwritten only after the peer has signalled the NEXT exchange (see gmg_solver.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "nccl_dyn.h"

namespace mgb {

constexpr int kP2PChannels = 8;
constexpr int kP2PMaxRanks = 16;
constexpr int kP2PMaxSeg = 28;
constexpr int kP2PChunk = 2048;          // doubles per work item (16 KB)
constexpr size_t kP2PHeaderBytes = 16384;

struct P2PHeader {                        // at offset 0 of every rank's pool
    unsigned long long flags[kP2PMaxRanks][kP2PChannels];   // flags[src][channel]: written by rank src
    unsigned long long push_seq[kP2PChannels];
    unsigned long long wait_seq[kP2PChannels];
    unsigned int done[kP2PChannels];
    unsigned int error;
    unsigned long long magic;
    // neighbour-only protocol (AMG ghost exchanges): messages sent to / consumed from every other rank so far; the flag a
    // rank publishes in flags[src][0] of a peer is its pair_push count for that peer
    unsigned long long pair_push[kP2PMaxRanks];
    unsigned long long pair_wait[kP2PMaxRanks];
};
static_assert(sizeof(P2PHeader) <= 4096, "header layout");

struct P2PSeg { const double *src; double *dst; unsigned long long n; };      // n doubles

struct P2PPush {
    int nseg;
    int first_item[kP2PMaxSeg + 1];       // work items (chunks) of segment s are [first_item[s], first_item[s+1])
    P2PSeg seg[kP2PMaxSeg];
    int n_sig;
    unsigned long long *sig[kP2PMaxRanks];    // flags[my rank][channel] in the header of every peer to signal
    unsigned long long *seq;                  // this rank's push_seq[channel]
    unsigned int *done;                       // this rank's done[channel]
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

static __global__ void __launch_bounds__(256)
k_p2p_push(const __grid_constant__ P2PPush a)
{
    const int total = a.first_item[a.nseg];
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int s = 0;
        while (item >= a.first_item[s + 1]) ++s;
        const P2PSeg sg = a.seg[s];
        const unsigned long long o = (unsigned long long)(item - a.first_item[s]) * kP2PChunk;
        const unsigned long long n = min((unsigned long long)kP2PChunk, sg.n - o);
        const double *src = sg.src + o;
        double *dst = sg.dst + o;
        if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            const unsigned long long n2 = n >> 1;
            for (unsigned long long i = threadIdx.x; i < n2; i += blockDim.x)
                reinterpret_cast<double2 *>(dst)[i] = reinterpret_cast<const double2 *>(src)[i];
            if ((n & 1) && threadIdx.x == 0) dst[n - 1] = src[n - 1];
        } else
            for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(a.done, 1u);
        if (prev == gridDim.x - 1) {                    // the last CTA: every CTA's stores are fenced before its atomicAdd
            __threadfence_system();
            *a.done = 0u;
            const unsigned long long s = *a.seq + 1ull;
            *a.seq = s;
            for (int i = 0; i < a.n_sig; ++i) st_release_sys(a.sig[i], s);
        }
    }
}

// waits until flags[src][channel] >= (++wait_seq[channel]) for every src in `mask`
static __global__ void k_p2p_wait(P2PHeader *h, int channel, unsigned int mask)
{
    const int src = threadIdx.x;
    const unsigned long long expect = h->wait_seq[channel] + 1ull;
    __syncwarp();
    if (src < kP2PMaxRanks && ((mask >> src) & 1u)) {
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned int spins = 0;
        while (ld_acquire_sys(&h->flags[src][channel]) < expect) {
            if ((++spins & 1023u) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) { atomicExch(&h->error, 1u + (unsigned)src); break; }
            }
        }
    }
    __syncwarp();
    if (threadIdx.x == 0) h->wait_seq[channel] = expect;
}

// ---- host side -----------------------------------------------------------------------------------------------------------
struct P2PComm {
    bool on = false;
    int rank = 0, n = 1;
    char *local = nullptr;
    size_t bytes = 0;
    std::vector<char *> peer;                 // peer[r]: base of rank r's pool in this process (peer[rank] = local)
    std::string why;                          // why it is off

    P2PHeader *hdr(int r) const { return reinterpret_cast<P2PHeader *>(peer[r]); }
    template <class T> T *at(int r, size_t byte_off) const { return reinterpret_cast<T *>(peer[r] + byte_off); }

    // collective over the communicator.  Leaves on = false (with `why`) when the pools cannot be mapped; never fails hard.
    void init(char *pool, size_t pool_bytes, int rank_, int n_, NcclComm comm, cudaStream_t st)
    {
        rank = rank_; n = n_; local = pool; bytes = pool_bytes;
        peer.assign(n, nullptr);
        peer[rank] = pool;
        on = false;
        if (n < 2) { why = "single rank"; return; }
        auto &N = nccl();
        int ok = (n <= kP2PMaxRanks) ? 1 : 0;
        if (!ok) why = "more ranks than one box holds";
        cudaIpcMemHandle_t mine{};
        if (ok && cudaIpcGetMemHandle(&mine, pool) != cudaSuccess) { cudaGetLastError(); ok = 0; why = "cudaIpcGetMemHandle failed"; }
        // header: zero flags and counters, magic = f(rank)
        P2PHeader h0{};
        h0.magic = 0xC0FFEE0000ull + (unsigned long long)rank;
        cudaMemcpyAsync(pool, &h0, sizeof(h0), cudaMemcpyHostToDevice, st);
        // handles of all ranks
        unsigned char *d_buf = nullptr;
        const size_t hb = sizeof(cudaIpcMemHandle_t);
        if (cudaMalloc(&d_buf, hb * (size_t)(n + 1)) != cudaSuccess) { cudaGetLastError(); why = "cudaMalloc"; return; }
        cudaMemcpyAsync(d_buf + hb * n, &mine, hb, cudaMemcpyHostToDevice, st);
        std::vector<cudaIpcMemHandle_t> all(n);
        bool coll_ok = N.AllGather(d_buf + hb * n, d_buf, hb, kNcclUint8, comm, st) == kNcclSuccess;
        cudaMemcpyAsync(all.data(), d_buf, hb * n, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        if (!coll_ok) { ok = 0; why = "ncclAllGather failed"; }
        for (int r = 0; ok && r < n; ++r) {
            if (r == rank) continue;
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError(); ok = 0; why = "cudaIpcOpenMemHandle failed"; break;
            }
            peer[r] = static_cast<char *>(p);
        }
        // every rank has written its header (stream order + the all-gather above was after the memcpy): check the magic of
        // every mapped pool -- a handle that maps an enclosing block instead of the pool would show here
        for (int r = 0; ok && r < n; ++r) {
            if (r == rank) continue;
            unsigned long long m = 0;
            if (cudaMemcpy(&m, &hdr(r)->magic, sizeof(m), cudaMemcpyDeviceToHost) != cudaSuccess || m != 0xC0FFEE0000ull + (unsigned long long)r) {
                cudaGetLastError(); ok = 0; why = "peer pool does not show the expected header";
            }
        }
        // all ranks must agree
        int *d_ok = reinterpret_cast<int *>(d_buf);
        int h_ok = ok;
        cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, st);
        // min over ranks via sum of (1 - ok)
        double *d_bad = reinterpret_cast<double *>(d_buf + 64);
        double bad = ok ? 0. : 1.;
        cudaMemcpyAsync(d_bad, &bad, sizeof(double), cudaMemcpyHostToDevice, st);
        N.AllReduce(d_bad, d_bad, 1, kNcclFloat64, kNcclSum, comm, st);
        cudaMemcpyAsync(&bad, d_bad, sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(d_buf);
        if (bad != 0.) { if (why.empty()) why = "another rank could not map the pools"; close_peers(); return; }
        on = true;
    }

    void close_peers()
    {
        for (int r = 0; r < (int)peer.size(); ++r)
            if (r != rank && peer[r]) { cudaIpcCloseMemHandle(peer[r]); peer[r] = nullptr; }
        on = false;
    }
};

// builder of one push launch
struct P2PPushBuilder {
    P2PPush a{};
    int items = 0;
    bool overflow = false;
    void seg(const double *src, double *dst, size_t n)
    {
        if (n == 0) return;
        if (a.nseg >= kP2PMaxSeg) { overflow = true; return; }
        a.seg[a.nseg] = P2PSeg{src, dst, (unsigned long long)n};
        a.first_item[a.nseg] = items;
        items += (int)((n + kP2PChunk - 1) / kP2PChunk);
        a.nseg++;
        a.first_item[a.nseg] = items;
    }
    void signal(unsigned long long *flag)
    {
        if (a.n_sig < kP2PMaxRanks) a.sig[a.n_sig++] = flag; else overflow = true;
    }
};

}  // namespace mgb
