// Onesweep LSD radix sort of (key, u32 value) pairs, keys u32 or u64, digits of 6/7/8 bits.
//
// The reference sorts R (tile << 32 | depth) keys with cub::DeviceRadixSort::SortPairs over 32 + log2(tiles) bits
// (cuda_rasterizer/rasterizer_impl.cu:612-617): 6 passes x 24 B per instance at 800x800. Here the same ordering is
// produced by two much smaller sorts (see preprocess.cu / api.cu):
//   1. the P Gaussians are sorted by their 32 depth bits (4 passes over P pairs of 8 B), instances are then EMITTED
//      in that order, and
//   2. the R instances are sorted by tile id only (12 bits at 800x800: 2 passes of 6-bit digits over 8-B pairs).
// Both sorts are stable, so the composition equals the stable sort on the full 44-bit key bit for bit (ties: same
// tile and same depth bits -> ascending Gaussian index, exactly what stability gives the reference).
//
// Structure of a pass (Adinets & Merrill, "Onesweep"): one histogram kernel computes the digit histograms of ALL
// passes in a single read of the keys; each pass is then ONE kernel that ranks a tile of keys in shared memory
// (match.any multisplit per warp), resolves its global digit offsets with a decoupled look-back over the preceding
// tiles (chained scan, tiles ordered by an atomic ticket so look-back never waits on a tile that has not started),
// and scatters key+value together through shared memory so global stores are coalesced runs.
#include <cstdlib>
#include "common.cuh"

namespace gigs {

constexpr int RS_HTHREADS = 256;   // histogram kernel
constexpr int RS_HU = 8;           // keys fetched per thread before any is counted
constexpr int RS_MAX_RADIX = 256;  // histogram / status rows are laid out for 8-bit digits
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_MIN_TILE = 2048;  // smallest tile of any configuration: sizes the look-back status array

constexpr uint32_t FLAG_AGG = 1u, FLAG_INC = 2u;
constexpr int LB_W = 8;  // look-back window: predecessor status words fetched per round


template <typename K>
__global__ void __launch_bounds__(RS_HTHREADS)
rs_histogram_kernel(const K* __restrict__ keys, const uint32_t n, const int passes, const int digit_bits, const int end_bit,
                    uint32_t* __restrict__ hist /*[passes][256]*/)
{
    pdl_enter();
    __shared__ uint32_t s_hist[RS_MAX_PASSES][RS_MAX_RADIX];
    for (int i = threadIdx.x; i < passes * RS_MAX_RADIX; i += RS_HTHREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    // RS_HU independent loads in flight per thread (the one-key-per-iteration loop was a chain of DRAM round trips:
    // 13.5 us for 300k keys, long-scoreboard bound)
    const uint32_t stride = gridDim.x * RS_HTHREADS;
    for (uint32_t i0 = blockIdx.x * RS_HTHREADS + threadIdx.x; i0 < n; i0 += stride * RS_HU) {
        K k[RS_HU];
#pragma unroll
        for (int u = 0; u < RS_HU; ++u) {
            const uint32_t i = i0 + u * stride;
            k[u] = (i < n) ? keys[i] : (K)0;
        }
#pragma unroll
        for (int u = 0; u < RS_HU; ++u) {
            if (i0 + u * stride >= n) break;
            for (int p = 0; p < passes; ++p) {
                const int shift = p * digit_bits;
                const int bits = min(digit_bits, end_bit - shift);
                const uint32_t d = (uint32_t)(k[u] >> shift) & ((1u << bits) - 1u);
                atomicAdd(&s_hist[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_MAX_RADIX; i += RS_HTHREADS) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---------------------------------------------------------------------------------------------
// Instance sort by tile id (<= RS_JOINT_BITS key bits): ONE histogram over the whole key gives everything the sort and
// the blend need — the digit histogram of every pass (sum of the joint bins sharing that digit) and the tile ranges
// (exclusive scan of the bins: in the sorted list tile t occupies [sum of bins < t, + bin t)), which the reference
// finds by comparing neighbours of the sorted keys in a launch of its own (identifyTileRanges,
// rasterizer_impl.cu:117-138).
// ---------------------------------------------------------------------------------------------
constexpr int RS_JOINT_BITS = 13;
constexpr int RS_JOINT_BINS = 1 << RS_JOINT_BITS;

__global__ void __launch_bounds__(512)
rs_joint_histogram_kernel(const uint32_t* __restrict__ keys, const uint32_t n, const int bins, uint32_t* __restrict__ joint)
{
    pdl_enter();
    extern __shared__ uint32_t s_joint[];
    for (int i = threadIdx.x; i < bins; i += 512) s_joint[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * 512 * 4;
    for (uint32_t i = (blockIdx.x * 512 + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n) {
            const uint4 k = *reinterpret_cast<const uint4*>(keys + i);
            atomicAdd(&s_joint[k.x], 1u); atomicAdd(&s_joint[k.y], 1u); atomicAdd(&s_joint[k.z], 1u); atomicAdd(&s_joint[k.w], 1u);
        } else {
            for (uint32_t j = i; j < n; ++j) atomicAdd(&s_joint[keys[j]], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += 512) {
        const uint32_t v = s_joint[i];
        if (v) atomicAdd(&joint[i], v);
    }
}

// one CTA: joint -> per-pass digit bases (exclusive scans, written to hist[p][*]) and tile ranges
__global__ void __launch_bounds__(1024)
rs_joint_scan_kernel(const uint32_t* __restrict__ joint, const int bins, const int passes, const int digit_bits,
                     const int end_bit, uint32_t* __restrict__ hist, uint2* __restrict__ ranges, const uint32_t num_tiles)
{
    pdl_enter();
    __shared__ uint32_t s_hist[RS_MAX_PASSES][RS_MAX_RADIX];
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < passes * RS_MAX_RADIX; i += 1024) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    // each thread owns a run of consecutive bins: ranges need the running sum in bin order
    const int per = (bins + 1023) / 1024;
    const int b0 = tid * per;
    uint32_t local = 0;
    for (int k = 0; k < per; ++k) {
        const int b = b0 + k;
        if (b < bins) {
            const uint32_t c = joint[b];
            local += c;
            if (c) {
                for (int p = 0; p < passes; ++p) {
                    const int shift = p * digit_bits;
                    const int bits = min(digit_bits, end_bit - shift);
                    atomicAdd(&s_hist[p][(b >> shift) & ((1 << bits) - 1)], c);
                }
            }
        }
    }
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t run = inc - local;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (int k = 0; k < per; ++k) {
        const int b = b0 + k;
        if (b < bins) {
            const uint32_t c = joint[b];
            // an empty tile keeps the (0, 0) the reference's zero-initialised ranges hold
            if ((uint32_t)b < num_tiles) ranges[b] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
            run += c;
        }
    }
    __syncthreads();
    // exclusive scan of every pass's digit histogram: warp w handles pass w
    if (warp < passes) {
        uint32_t carry = 0;
        for (int c0 = 0; c0 < RS_MAX_RADIX; c0 += 32) {
            const uint32_t v = s_hist[warp][c0 + lane];
            uint32_t in2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, in2, o);
                if (lane >= o) in2 += t;
            }
            hist[warp * RS_MAX_RADIX + c0 + lane] = carry + in2 - v;
            carry += __shfl_sync(0xffffffffu, in2, 31);
        }
    }
}

template <typename K, int BITS, int THREADS, int ITEMS>
struct RsSmem {
    static constexpr int RADIX = 1 << BITS;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    K keys[TILE];
    uint32_t vals[TILE];
    uint32_t warp_hist[WARPS][RADIX];
    uint32_t digit_start[RADIX];
    uint32_t global_off[RADIX];
    uint32_t warp_tot[RADIX / 32];
    uint32_t warp_tot_h[RADIX / 32];
    uint32_t tile;
};

// vals_in == nullptr: the value of element i is i (first pass of an argsort; saves materialising an iota)
template <typename K, int BITS, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
rs_onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out, const uint32_t* __restrict__ vals_in,
                   uint32_t* __restrict__ vals_out, const uint32_t n, const int shift, const uint32_t digit_mask,
                   const uint32_t* __restrict__ digit_base, volatile uint32_t* __restrict__ status /*[tiles][RADIX]*/,
                   uint32_t* __restrict__ ticket, const bool raw_hist)
{
    pdl_enter();
    using Smem = RsSmem<K, BITS, THREADS, ITEMS>;
    constexpr int RADIX = Smem::RADIX, WARPS = Smem::WARPS, TILE = Smem::TILE;
    static_assert(RADIX % 32 == 0 && RADIX <= THREADS, "digit threads must be whole warps of the block");
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(rs_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const uint32_t base = tile * (uint32_t)TILE;
    const uint32_t count = min((uint32_t)TILE, n - base);

    // warp-striped load: item i of lane l in warp w = base + w*32*ITEMS + i*32 + l (keeps stability)
    K key[ITEMS];
    uint16_t rank[ITEMS];
    const uint32_t wbase = warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        key[i] = (loc < count) ? keys_in[base + loc] : (K)~(K)0;
    }

    // rank keys within the warp, item by item (multisplit). The set of lanes holding the same digit is built from
    // one ballot per digit bit (VOTE is a cheap ALU-side op; MATCH.ANY measured ~100 cycles of issue per warp
    // instruction per SM here and bounded the whole pass). All lanes of a digit read its running count (a shared
    // memory broadcast), then the lowest lane of the set adds the set size.
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        const bool valid = loc < count;
        const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
        uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int bit = 0; bit < BITS; ++bit) {
            const bool on = (d >> bit) & 1u;
            const uint32_t m = __ballot_sync(0xffffffffu, on);
            peers &= on ? m : ~m;
        }
        const uint32_t below = peers & lt_mask;
        uint32_t old = 0;
        if (valid) old = S.warp_hist[warp][d];
        __syncwarp();
        if (valid && below == 0) S.warp_hist[warp][d] = old + __popc(peers);
        __syncwarp();
        rank[i] = (uint16_t)(old + __popc(below));
    }
    __syncthreads();

    // values are fetched now so that their latency overlaps the look-back below
    uint32_t val[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        val[i] = (loc < count) ? (vals_in ? vals_in[base + loc] : base + loc) : 0u;
    }

    // thread d: exclusive scan over the warps' counts of digit d; tile total; look-back
    if (tid < RADIX) {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t t = S.warp_hist[w][d];
            S.warp_hist[w][d] = sum;
            sum += t;
        }
        volatile uint32_t* my_status = status + (size_t)tile * RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            *my_status = (sum << 2) | FLAG_INC;
        } else {
            *my_status = (sum << 2) | FLAG_AGG;
            // Look-back with LB_W predecessors in flight per round: the loads are independent, so a round costs one
            // L2 round trip instead of LB_W (measured serial depth at 300k / 800x800: 16 tiles on average, 35 max).
            int64_t t = (int64_t)tile - 1;
            bool found = false;
            while (!found) {
                uint32_t v[LB_W];
#pragma unroll
                for (int k = 0; k < LB_W; ++k)
                    v[k] = (t - k >= 0) ? status[(size_t)(t - k) * RADIX + d] : (uint32_t)2u /* FLAG_INC: before tile 0 the prefix is 0 */;
#pragma unroll
                for (int k = 0; k < LB_W; ++k) {
                    if (!found) {
                        while ((v[k] & 3u) == 0u) v[k] = status[(size_t)(t - k) * RADIX + d];
                        excl += v[k] >> 2;
                        found = (v[k] & FLAG_INC) != 0u;
                    }
                }
                t -= LB_W;
            }
            *my_status = ((excl + sum) << 2) | FLAG_INC;
        }
        // exclusive scan of the tile's digit totals -> position of each digit inside the tile; with raw_hist the
        // same shuffles also scan the global digit counts (digit_base then holds the histogram itself, not its
        // exclusive scan: one launch fewer in front of the passes)
        const uint32_t hv = digit_base[d];
        uint32_t inc = sum, hinc = hv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            uint32_t th = __shfl_up_sync(0xffffffffu, hinc, o);
            if (lane >= o) {
                inc += t;
                hinc += th;
            }
        }
        if (lane == 31) {
            S.warp_tot[warp] = inc;
            S.warp_tot_h[warp] = hinc;
        }
        // barrier among the RADIX digit threads only (the block is larger)
        asm volatile("bar.sync 1, %0;" ::"n"(RADIX) : "memory");
        uint32_t woff = 0, hoff = 0;
        for (int w = 0; w < warp; ++w) {
            woff += S.warp_tot[w];
            hoff += S.warp_tot_h[w];
        }
        const uint32_t dstart = woff + inc - sum;
        const uint32_t dbase = raw_hist ? (hoff + hinc - hv) : hv;
        S.digit_start[d] = dstart;
        S.global_off[d] = dbase + excl - dstart;  // wraps mod 2^32 on purpose
    }
    __syncthreads();

    // scatter keys into tile-sorted order in shared memory
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
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
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t loc = wbase + i * 32;
        if (loc < count) S.vals[rank[i]] = val[i];
    }
    __syncthreads();
    // coalesced write-out: consecutive smem positions of one digit go to consecutive global slots
    for (uint32_t p = tid; p < count; p += THREADS) {
        const K k = S.keys[p];
        const uint32_t d = (uint32_t)(k >> shift) & digit_mask;
        const uint32_t g = S.global_off[d] + p;
        keys_out[g] = k;
        vals_out[g] = S.vals[p];
    }
}

template <typename K, int BITS, int T, int I>
static int launch_passes(uint32_t n, int end_bit, const K* keys_u, const uint32_t* vals_u, K* keys_a, uint32_t* vals_a,
                         K* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status, uint32_t* tickets,
                         uint32_t status_tiles, int pass_stage, bool raw_hist, cudaStream_t st)
{
    using Smem = RsSmem<K, BITS, T, I>;
    static_assert(T * I <= 65536 && T * I >= RS_MIN_TILE, "tile must fit the u16 ranks and the status allocation");
    GIGS_SMEM_ATTR((rs_onesweep_kernel<K, BITS, T, I>), sizeof(Smem));
    const int passes = (end_bit + BITS - 1) / BITS;
    const uint32_t tiles = (n + T * I - 1) / (T * I);
    const K* kin = keys_u;
    const uint32_t* vin = vals_u;
    for (int p = 0; p < passes; ++p) {
        // choose outputs so that the last pass lands in (keys_a, vals_a)
        const bool to_a = ((passes - 1 - p) % 2) == 0;
        K* kout = to_a ? keys_a : keys_b;
        uint32_t* vout = to_a ? vals_a : vals_b;
        const int shift = p * BITS;
        const int bits = (end_bit - shift) < BITS ? (end_bit - shift) : BITS;
        ProfScope ps(pass_stage, st);
        GIGS_CUDA(launch_k(rs_onesweep_kernel<K, BITS, T, I>, dim3(tiles), dim3(T), (size_t)(sizeof(Smem)), st, 
            kin, kout, vin, vout, n, shift, (1u << bits) - 1u, hist + p * RS_MAX_RADIX,
            status + (size_t)p * status_tiles * RS_MAX_RADIX, tickets + p, raw_hist));
        GIGS_LAUNCH_CHECK("rs_onesweep_kernel");
        kin = kout;
        vin = vout;
    }
    return 0;
}

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// digit width for a key of `bits` significant bits: the fewest passes, then the narrowest digit that still covers
// the key (narrow digits = small per-warp histograms and short look-back rows)
int radix_digit_bits(int bits)
{
    const int passes = (bits + 7) / 8;
    int d = (bits + passes - 1) / passes;
    if (d < 6) d = 6;
    return d;
}
int radix_sort_passes(int bits) { return (bits + 7) / 8; }

template <typename K>
static int sort_pairs(uint64_t R, int end_bit, const K* keys_u, const uint32_t* vals_u, K* keys_a, uint32_t* vals_a,
                      K* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status, uint32_t* tickets,
                      uint64_t zero_bytes, int pass_stage, cudaStream_t st, uint32_t* joint = nullptr,
                      uint2* ranges = nullptr, uint32_t num_tiles = 0)
{
    if (R == 0) return 0;
    if (R >= (1ull << 30)) {
        set_error("radix sort: %llu pairs exceeds the 2^30 limit", (unsigned long long)R);
        return -3;
    }
    const uint32_t n = (uint32_t)R;
    const int digit_bits = radix_digit_bits(end_bit);
    const int passes = (end_bit + digit_bits - 1) / digit_bits;
    const uint32_t status_tiles = (n + RS_MIN_TILE - 1) / RS_MIN_TILE;
    bool raw_hist = false;
    // hist, tickets and status are contiguous in the scratch blob: one memset (zero_bytes == 0: the caller's previous
    // kernel already cleared them)
    if (zero_bytes) GIGS_CUDA(cudaMemsetAsync(hist, 0, zero_bytes, st));
    if (joint != nullptr && sizeof(K) == 4 && end_bit <= RS_JOINT_BITS) {
        const int bins = 1 << end_bit;
        const uint32_t hblocks = min((n + 16383u) / 16384u, 148u * 2u);
        GIGS_SMEM_ATTR(rs_joint_histogram_kernel, RS_JOINT_BINS * 4);
        GIGS_CUDA(launch_k(rs_joint_histogram_kernel, dim3(hblocks), dim3(512), (size_t)((size_t)bins * 4), st, (const uint32_t*)keys_u, n, bins, joint));
        GIGS_LAUNCH_CHECK("rs_joint_histogram_kernel");
        GIGS_CUDA(launch_k(rs_joint_scan_kernel, dim3(1), dim3(1024), (size_t)(0), st, joint, bins, passes, digit_bits, end_bit, hist, ranges, num_tiles));
        GIGS_LAUNCH_CHECK("rs_joint_scan_kernel");
    } else {
        const uint32_t hblocks = min((n + 2047u) / 2048u, 148u * 8u);
        GIGS_CUDA(launch_k(rs_histogram_kernel<K>, dim3(hblocks), dim3(RS_HTHREADS), (size_t)(0), st, keys_u, n, passes, digit_bits, end_bit, hist));
        GIGS_LAUNCH_CHECK("rs_histogram_kernel");
        raw_hist = true;   // the passes scan the digit counts themselves (no scan launch in front of them)
    }
    static const int cfg = env_int("GIGS_RS_CFG", 0);
#define RS_GO(B, T, I) \
    return launch_passes<K, B, T, I>(n, end_bit, keys_u, vals_u, keys_a, vals_a, keys_b, vals_b, hist, status, tickets, status_tiles, pass_stage, raw_hist, st)
    // small inputs: small tiles so that every SM gets work; large inputs: big tiles for long coalesced runs
    const bool small = n < 148u * 8192u;
    if (digit_bits == 6) {
        if (cfg == 1) RS_GO(6, 256, 8);
        if (cfg == 2) RS_GO(6, 512, 16);
        if (cfg == 3) RS_GO(6, 256, 16);
        if (cfg == 4) RS_GO(6, 1024, 4);
        if (cfg == 5) RS_GO(6, 512, 4);
        if (small) RS_GO(6, 256, 8);
        RS_GO(6, 1024, 4);  // measured best at R = 4.2M (B200): 40.6 us / pass vs 43.3 for 512 x 8
    } else if (digit_bits == 7) {
        if (small) RS_GO(7, 256, 8);
        RS_GO(7, 512, 8);
    } else {
        if (cfg == 1) RS_GO(8, 256, 8);
        if (cfg == 2) RS_GO(8, 512, 16);
        if (cfg == 3) RS_GO(8, 256, 16);
        if (cfg == 4) RS_GO(8, 512, 4);
        if (cfg == 5) RS_GO(8, 1024, 2);
        if (small) RS_GO(8, 512, 4);  // measured best for the 300k-key depth argsort: 60 us vs 70 for 256 x 8
        RS_GO(8, 512, 16);
    }
#undef RS_GO
}

// Sort R pairs on key bits [0, end_bit). Input in (keys_u, vals_u) (left intact; vals_u == nullptr means value i = i);
// result in (keys_a, vals_a). keys_b/vals_b are the ping-pong partners. hist / tickets / status are contiguous and
// zero_bytes long (see radix_scratch_bytes).
int launch_radix_sort(uint64_t R, int end_bit, const uint64_t* keys_u, const uint32_t* vals_u, uint64_t* keys_a,
                      uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                      uint32_t* tickets, uint64_t zero_bytes, cudaStream_t st)
{
    return sort_pairs<uint64_t>(R, end_bit, keys_u, vals_u, keys_a, vals_a, keys_b, vals_b, hist, status, tickets, zero_bytes, -1, st);
}
int launch_radix_sort32(uint64_t R, int end_bit, const uint32_t* keys_u, const uint32_t* vals_u, uint32_t* keys_a,
                        uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                        uint32_t* tickets, uint64_t zero_bytes, int pass_stage, cudaStream_t st)
{
    return sort_pairs<uint32_t>(R, end_bit, keys_u, vals_u, keys_a, vals_a, keys_b, vals_b, hist, status, tickets, zero_bytes, pass_stage, st);
}
// the instance sort: keys are tile ids. joint = RS_JOINT_BINS zeroed words inside the zeroed scratch region. Returns 1
// in *ranges_done when the tile ranges were produced from the histogram (keys of at most RS_JOINT_BITS bits).
int launch_tile_sort(uint64_t R, int end_bit, const uint32_t* keys_u, const uint32_t* vals_u, uint32_t* keys_a,
                     uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                     uint32_t* tickets, uint64_t zero_bytes, int pass_stage, uint32_t* joint, uint2* ranges,
                     uint32_t num_tiles, int* ranges_done, cudaStream_t st)
{
    *ranges_done = (R > 0 && joint != nullptr && end_bit <= RS_JOINT_BITS) ? 1 : 0;
    return sort_pairs<uint32_t>(R, end_bit, keys_u, vals_u, keys_a, vals_a, keys_b, vals_b, hist, status, tickets, zero_bytes,
                                pass_stage, st, joint, ranges, num_tiles);
}
uint64_t radix_joint_bytes() { return (uint64_t)RS_JOINT_BINS * 4; }


uint32_t radix_sort_tiles(uint64_t R) { return (uint32_t)((R + RS_MIN_TILE - 1) / RS_MIN_TILE); }

}  // namespace gigs
