// Onesweep LSD radix sort of (u64 key, u32 value) pairs — replaces cub::DeviceRadixSort::SortPairs at
// reference cuda_rasterizer/rasterizer_impl.cu:612-617. Stable, so the result is bit-identical to any
// other stable sort on the same key bits [0, end_bit).
//
// Structure (Adinets & Merrill, "Onesweep"): one histogram kernel computes the digit histograms of ALL
// passes in a single read of the keys; each pass is then ONE kernel that ranks a 4096-key tile in
// shared memory, resolves its global digit offsets with a decoupled look-back over the preceding
// tiles (chained scan, tiles ordered by an atomic ticket so look-back never waits on a tile that has
// not started), and scatters key+value together. HBM traffic per pass = read 12 B + write 12 B per pair.
#include <cstdlib>
#include "common.cuh"

namespace gigs {

constexpr int RS_HTHREADS = 256;  // histogram kernel
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u, FLAG_INC = 2u;

__global__ void __launch_bounds__(RS_HTHREADS)
rs_histogram_kernel(const uint64_t* __restrict__ keys, const uint32_t n, const int passes, const int end_bit,
                    uint32_t* __restrict__ hist /*[passes][256]*/)
{
    __shared__ uint32_t s_hist[RS_MAX_PASSES][RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_HTHREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * RS_HTHREADS;
    for (uint32_t i = blockIdx.x * RS_HTHREADS + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
        for (int p = 0; p < passes; ++p) {
            const int shift = p * 8;
            const int bits = min(8, end_bit - shift);
            const uint32_t d = (uint32_t)(k >> shift) & ((1u << bits) - 1u);
            atomicAdd(&s_hist[p][d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += RS_HTHREADS) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// exclusive scan of each pass's 256-bin histogram (in place)
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_warp[RS_RADIX / 32];
    uint32_t* h = hist + blockIdx.x * RS_RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    h[threadIdx.x] = off + inc - v;
}

template <int RS_THREADS, int RS_ITEMS>
struct RsSmem {
    static constexpr int RS_WARPS = RS_THREADS / 32;
    static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint32_t warp_hist[RS_WARPS][RS_RADIX];
    uint32_t digit_start[RS_RADIX];
    uint32_t global_off[RS_RADIX];
    uint32_t warp_tot[RS_RADIX / 32];
    uint32_t tile;
};

template <int RS_THREADS, int RS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
rs_onesweep_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                   const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, const uint32_t n,
                   const int shift, const uint32_t digit_mask, const uint32_t* __restrict__ digit_base,
                   volatile uint32_t* __restrict__ status /*[tiles][256]*/, uint32_t* __restrict__ ticket)
{
    constexpr int RS_WARPS = RS_THREADS / 32;
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem<RS_THREADS, RS_ITEMS>& S = *reinterpret_cast<RsSmem<RS_THREADS, RS_ITEMS>*>(rs_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const uint32_t base = tile * (uint32_t)RS_TILE;
    const uint32_t count = min((uint32_t)RS_TILE, n - base);

    // warp-striped load: item i of lane l in warp w = base + w*512 + i*32 + l (keeps stability)
    uint64_t key[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
    const uint32_t wbase = warp * (32 * RS_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        key[i] = (loc < count) ? keys_in[base + loc] : ~0ull;
    }

    // rank keys within the warp, item by item (match-any multisplit)
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        const bool valid = loc < count;
        const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & digit_mask) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = peers & lt_mask;
        uint32_t old = 0;
        if (valid && below == 0) {
            old = S.warp_hist[warp][d];
            S.warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[i] = (uint16_t)(old + __popc(below));
        __syncwarp();
    }
    __syncthreads();

    // values are fetched now so that their latency overlaps the look-back below
    uint32_t val[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        val[i] = (loc < count) ? vals_in[base + loc] : 0u;
    }

    // thread d: exclusive scan over the warps' counts of digit d; tile total; look-back
    if (tid < RS_RADIX) {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t t = S.warp_hist[w][d];
            S.warp_hist[w][d] = sum;
            sum += t;
        }
        volatile uint32_t* my_status = status + (size_t)tile * RS_RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            *my_status = (sum << 2) | FLAG_INC;
        } else {
            *my_status = (sum << 2) | FLAG_AGG;
            int64_t t = (int64_t)tile - 1;
            while (true) {
                uint32_t v;
                do {
                    v = status[(size_t)t * RS_RADIX + d];
                } while ((v & 3u) == 0u);
                excl += v >> 2;
                if (v & FLAG_INC) break;
                --t;
            }
            *my_status = ((excl + sum) << 2) | FLAG_INC;
        }
        // block exclusive scan of the tile's digit totals -> position of each digit inside the tile
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) S.warp_tot[warp] = inc;
        // barrier among the RS_RADIX digit threads only (the block may be larger)
        asm volatile("bar.sync 1, %0;" ::"n"(RS_RADIX) : "memory");
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += S.warp_tot[w];
        const uint32_t dstart = woff + inc - sum;
        S.digit_start[d] = dstart;
        S.global_off[d] = digit_base[d] + excl - dstart;  // wraps mod 2^32 on purpose
    }
    __syncthreads();

    // scatter keys into tile-sorted order in shared memory
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        if (loc < count) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
            const uint32_t pos = S.digit_start[d] + S.warp_hist[warp][d] + rank[i];
            S.keys[pos] = key[i];
            rank[i] = (uint16_t)pos;
        }
    }
    // values follow the same permutation
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        if (loc < count) S.vals[rank[i]] = val[i];
    }
    __syncthreads();
    // coalesced write-out: consecutive smem positions of one digit go to consecutive global slots
    for (uint32_t p = tid; p < count; p += RS_THREADS) {
        const uint64_t k = S.keys[p];
        const uint32_t d = (uint32_t)(k >> shift) & digit_mask;
        const uint32_t g = S.global_off[d] + p;
        keys_out[g] = k;
        vals_out[g] = S.vals[p];
    }
}

template <int T, int I>
static int launch_passes(uint32_t n, int end_bit, const uint64_t* keys_u, const uint32_t* vals_u, uint64_t* keys_a,
                         uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                         uint32_t* tickets, uint32_t status_tiles, cudaStream_t st)
{
    using Smem = RsSmem<T, I>;
    static bool attr_set = false;
    if (!attr_set) {
        GIGS_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        attr_set = true;
    }
    const int passes = (end_bit + 7) / 8;
    const uint32_t tiles = (n + T * I - 1) / (T * I);
    const uint64_t* kin = keys_u;
    const uint32_t* vin = vals_u;
    for (int p = 0; p < passes; ++p) {
        // choose outputs so that the last pass lands in (keys_a, vals_a)
        const bool to_a = ((passes - 1 - p) % 2) == 0;
        uint64_t* kout = to_a ? keys_a : keys_b;
        uint32_t* vout = to_a ? vals_a : vals_b;
        const int shift = p * 8;
        const int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        ProfScope ps(ST_SORT_PASS, st);
        rs_onesweep_kernel<T, I><<<tiles, T, sizeof(Smem), st>>>(kin, kout, vin, vout, n, shift, (1u << bits) - 1u,
                                                                 hist + p * RS_RADIX,
                                                                 status + (size_t)p * status_tiles * RS_RADIX,
                                                                 tickets + p);
        GIGS_LAUNCH_CHECK("rs_onesweep_kernel");
        kin = kout;
        vin = vout;
    }
    return 0;
}

constexpr int RS_MIN_TILE = 3072;  // smallest tile of any configuration: sizes the look-back status array

// Sort R pairs on key bits [0, end_bit). Input in (keys_u, vals_u) (left intact); result in
// (keys_a, vals_a). keys_b/vals_b are the ping-pong partners.
int launch_radix_sort(uint64_t R, int end_bit, const uint64_t* keys_u, const uint32_t* vals_u, uint64_t* keys_a,
                      uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                      uint32_t* tickets, uint64_t status_bytes_total, cudaStream_t st)
{
    if (R == 0) return 0;
    if (R >= (1ull << 30)) {
        set_error("radix sort: %llu instances exceeds the 2^30 limit", (unsigned long long)R);
        return -3;
    }
    const int passes = (end_bit + 7) / 8;
    const uint32_t n = (uint32_t)R;
    const uint32_t status_tiles = (n + RS_MIN_TILE - 1) / RS_MIN_TILE;
    // hist, tickets and status are contiguous in the scratch blob: one memset
    GIGS_CUDA(cudaMemsetAsync(hist, 0, status_bytes_total, st));
    const uint32_t hblocks = min((n + 4095u) / 4096u, 148u * 8u);
    rs_histogram_kernel<<<hblocks, RS_HTHREADS, 0, st>>>(keys_u, n, passes, end_bit, hist);
    GIGS_LAUNCH_CHECK("rs_histogram_kernel");
    rs_scan_hist_kernel<<<passes, RS_RADIX, 0, st>>>(hist);
    GIGS_LAUNCH_CHECK("rs_scan_hist_kernel");
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("GIGS_RS_CFG");
        cfg = e ? atoi(e) : 2;
    }
#define RS_GO(T, I) return launch_passes<T, I>(n, end_bit, keys_u, vals_u, keys_a, vals_a, keys_b, vals_b, hist, status, tickets, status_tiles, st)
    switch (cfg) {
        case 1: RS_GO(256, 12);
        case 2: RS_GO(512, 16);
        case 3: RS_GO(384, 16);
        case 4: RS_GO(512, 12);
        default: RS_GO(256, 16);
    }
#undef RS_GO
}

uint32_t radix_sort_tiles(uint64_t R) { return (uint32_t)((R + RS_MIN_TILE - 1) / RS_MIN_TILE); }

}  // namespace gigs
