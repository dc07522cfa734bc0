// Gradient all-reduce over NVLink peer memory ("per-Gaussian gradient allreduce over NVLink"), as ONE kernel
// of our own instead of two or three NCCL collectives: the view-sharded training step ends with every rank holding its
// partial gradients in a buffer that all ranks can address (a symmetric allocation, one per rank, mapped into every
// process: torch.distributed._symmetric_memory on the host side, plain device pointers here). One launch per rank does
//   (1) a cross-rank barrier through flag words in peer memory ("my gradients are complete"),
//   (2) a two-shot all-reduce IN PLACE: rank r owns slice r of every span — it loads slice r from all N buffers over
//       NVLink (16-B loads), adds them in rank order (so every rank ends up with bit-identical sums, and the result does
//       not depend on timing) and stores the sum into slice r of all N buffers (16-B stores). Slice r of any buffer is
//       read and written by rank r only, so the in-place update has no cross-rank hazard,
//   (3) a second barrier ("my slice is written everywhere"), after which the kernel — and with it the stream — proceeds.
// 12.6 MB of PBR-stage gradients (materials + light textures) at 8 GPUs: each rank moves 1.4 MB in and 1.4 MB out per
// peer; the cost is the two barrier latencies plus ~10 us of transfers, against ~60-150 us for the NCCL sequence it
// replaces (launch + protocol latency of each collective). A peer that has not arrived after PR_TIMEOUT_CYCLES (about a
// minute of GPU clock: stragglers — a rank writing a checkpoint, evaluating, loading data — are waited for, like
// NCCL does) is a fatal error: the error word is set and the kernel traps, so the next CUDA call of this process
// fails instead of the ranks training on with unreduced, diverging gradients.
#include <cstdlib>
#include "common.cuh"

namespace gigs {

constexpr int PR_MAX_WORLD = 16;
constexpr int PR_MAX_SPANS = 16;
constexpr int PR_THREADS = 512;

struct PeerArgs {
    int world, rank, n_spans;
    unsigned int epoch;                       // call counter, starts at 1
    float* buf[PR_MAX_WORLD];                 // the symmetric gradient buffer of every rank (buf[rank] is local)
    float* mc;                                // multicast (NVLS) mapping of the same buffer, or NULL
    unsigned int* flags[PR_MAX_WORLD];        // every rank's flag block: [0..W) ready, [W..2W) done, [2W] local CTA counter, [2W+1] error
    unsigned long long lo[PR_MAX_SPANS], hi[PR_MAX_SPANS];   // float offsets
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long PR_TIMEOUT_CYCLES = 120000000000ll;   // ~60 s at 1.9 GHz
// wait until *p >= want (flags only grow); false on timeout
__device__ __forceinline__ bool spin_ge(const unsigned int* p, unsigned int want)
{
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - want) < 0) {
        if (clock64() - t0 > PR_TIMEOUT_CYCLES) return false;
        __nanosleep(64);
    }
    return true;
}

// NVLS: one load that the NVSwitch answers with the SUM of the addressed 16 bytes over every rank's copy of the buffer,
// and one store that the switch writes into every rank's copy (PTX multimem.*, sm_90+). Per rank and slice the links
// carry the slice once out and once in instead of (N-1) times each way.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc_addr)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc_addr, const float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(PR_THREADS) peer_allreduce_kernel(const __grid_constant__ PeerArgs A)
{
    const int W = A.world, r = A.rank;
    unsigned int* my = A.flags[r];
    const unsigned int ready = A.epoch, done = A.epoch;
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    // (1) everybody's gradients are complete: CTA 0 tells the peers, every CTA waits for all of them
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(A.flags[threadIdx.x] + r, ready);
    }
    if (threadIdx.x < W && !spin_ge(my + threadIdx.x, ready)) ok = 0;
    __syncthreads();
    if (!ok) {
        // a peer never arrived: there is nothing to reduce and nothing sane to continue with. Record which call it
        // was (host-readable: PeerBuffer.error_epoch) and abort the kernel; the context's next call returns the error.
        if (threadIdx.x == 0) {
            my[2 * W + 1] = A.epoch;
            __threadfence_system();
        }
        __syncthreads();
        __trap();
    }
    // (2) reduce my slice of every span and write it to everybody
    for (int s = 0; s < A.n_spans; ++s) {
        const unsigned long long lo = A.lo[s], n = A.hi[s] - lo;
        const bool vec = ((lo | n) & 3ull) == 0;
        if (vec) {
            const unsigned long long n4 = n >> 2;
            const unsigned long long per = (n4 + W - 1) / W;
            const unsigned long long b = (unsigned long long)r * per, e = (b + per < n4) ? b + per : n4;
            if (A.mc != nullptr) {
                // the switch adds (order fixed by the fabric, the same for every element of a run) and broadcasts
                // UM load-reduces in flight per thread: one at a time left a thread waiting a switch round trip per 16 bytes
                constexpr int UM = 4;
                const unsigned long long mstride = (unsigned long long)gridDim.x * PR_THREADS;
                for (unsigned long long i0 = b + (unsigned long long)blockIdx.x * PR_THREADS + threadIdx.x; i0 < e;
                     i0 += UM * mstride) {
                    float4 acc[UM];
#pragma unroll
                    for (int u = 0; u < UM; ++u)
                        if (i0 + u * mstride < e) acc[u] = multimem_ld_reduce_f4(A.mc + lo + 4 * (i0 + u * mstride));
#pragma unroll
                    for (int u = 0; u < UM; ++u)
                        if (i0 + u * mstride < e) multimem_st_f4(A.mc + lo + 4 * (i0 + u * mstride), acc[u]);
                }
                continue;
            }
            // U elements per thread and trip: U * W independent 16-byte peer loads in flight (at W = 2 one element per
            // trip leaves the links waiting on latency)
            constexpr int U = 4;
            const unsigned long long stride = (unsigned long long)gridDim.x * PR_THREADS;
            for (unsigned long long i0 = b + (unsigned long long)blockIdx.x * PR_THREADS + threadIdx.x; i0 < e;
                 i0 += U * stride) {
                float4 acc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = 0; p < W; ++p) {
                    const float4* src = reinterpret_cast<const float4*>(A.buf[p] + lo);
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        v[u] = (i0 + u * stride < e) ? src[i0 + u * stride] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < U; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
                }
                for (int p = 0; p < W; ++p) {
                    float4* dst = reinterpret_cast<float4*>(A.buf[p] + lo);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (i0 + u * stride < e) dst[i0 + u * stride] = acc[u];
                }
            }
        } else {
            const unsigned long long per = (n + W - 1) / W;
            const unsigned long long b = (unsigned long long)r * per, e = (b + per < n) ? b + per : n;
            for (unsigned long long i = b + (unsigned long long)blockIdx.x * PR_THREADS + threadIdx.x; i < e;
                 i += (unsigned long long)gridDim.x * PR_THREADS) {
                float acc = 0.f;
                for (int p = 0; p < W; ++p) acc += A.buf[p][lo + i];
                for (int p = 0; p < W; ++p) A.buf[p][lo + i] = acc;
            }
        }
    }
    // (3) my slice is written everywhere: the last CTA of this rank tells the peers and waits for theirs
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) {
        // calls on one flag block are stream-ordered (never concurrent), so the last CTA can hand the counter back at 0:
        // the grid size may differ from call to call
        const unsigned int c = atomicAdd(my + 2 * W, 1u) + 1u;
        last = (c == gridDim.x);
        if (last) my[2 * W] = 0u;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(A.flags[threadIdx.x] + W + r, done);
        if (!spin_ge(my + W + threadIdx.x, done)) {
            my[2 * W + 1] = A.epoch;
            __threadfence_system();
            __trap();
        }
    }
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_peer_allreduce(int32_t world, int32_t rank, const uint64_t* peer_bufs, const uint64_t* peer_flags,
                        uint64_t multicast_buf, uint32_t epoch, int32_t n_spans, const uint64_t* span_begin,
                        const uint64_t* span_end, int32_t n_ctas, void* stream)
{
    if (world < 1 || world > PR_MAX_WORLD || rank < 0 || rank >= world || n_spans < 0 || n_spans > PR_MAX_SPANS ||
        !peer_bufs || !peer_flags || epoch == 0 || (n_spans && (!span_begin || !span_end))) {
        set_error("gigs_peer_allreduce: bad arguments (world 1..%d, spans 0..%d, epoch >= 1)", PR_MAX_WORLD, PR_MAX_SPANS);
        return -1;
    }
    PeerArgs A;
    A.world = world; A.rank = rank; A.n_spans = n_spans; A.epoch = epoch;
    A.mc = (float*)multicast_buf;
    for (int p = 0; p < world; ++p) {
        if (!peer_bufs[p] || !peer_flags[p]) { set_error("gigs_peer_allreduce: peer %d has a NULL pointer", p); return -1; }
        A.buf[p] = (float*)peer_bufs[p];
        A.flags[p] = (unsigned int*)peer_flags[p];
    }
    for (int s = 0; s < n_spans; ++s) {
        if (span_end[s] < span_begin[s]) { set_error("gigs_peer_allreduce: span %d is reversed", s); return -1; }
        A.lo[s] = span_begin[s]; A.hi[s] = span_end[s];
    }
    if (n_ctas <= 0) {
        // scale with the bytes: ~128 KB of span per CTA, between 8 (latency-bound small exchanges: fewer CTAs to
        // gather at the barriers) and 128 (an 80 MB first-stage buffer needs the loads of many SMs in flight)
        uint64_t floats = 0;
        for (int s = 0; s < n_spans; ++s) floats += span_end[s] - span_begin[s];
        static const uint64_t kb_per_cta = getenv("GIGS_PEER_KB_PER_CTA") ? (uint64_t)atoi(getenv("GIGS_PEER_KB_PER_CTA")) : 128;
        const uint64_t want = (floats * 4 / (uint64_t)world) / ((kb_per_cta ? kb_per_cta : 128) * 1024) + 1;
        n_ctas = (int)(want < 8 ? 8 : (want > 128 ? 128 : want));
    }
    if (n_ctas > 148) n_ctas = 148;     // every CTA of every call must be counted exactly once by the counter protocol
    ProfScope ps(ST_PEER_ALLREDUCE, (cudaStream_t)stream);
    peer_allreduce_kernel<<<n_ctas, PR_THREADS, 0, (cudaStream_t)stream>>>(A);
    GIGS_LAUNCH_CHECK("peer_allreduce_kernel");
    return 0;
}

}  // extern "C"
