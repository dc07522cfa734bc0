// distCUDA2: mean of the 3 smallest squared distances to other points. Replaces SimpleKNN::knn
// (/root/reference/submodules/simple-knn/simple_knn.cu:165-207; kernels :57,71,130).
//
// Same exact-result contract as the reference (the Morton/box structure only prunes; each d^2 is
// formed as dx*dx + dy*dy + dz*dz in that order and the result is (b0+b1+b2)/3), but no thrust
// allocations, no host round trips for the bounding box, the caller's stream, and our own
// onesweep radix sort for the 30-bit Morton keys.
#include <cfloat>
#include "common.cuh"

namespace gigs {

int launch_radix_sort(uint64_t R, int end_bit, const uint64_t* keys_u, const uint32_t* vals_u, uint64_t* keys_a,
                      uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                      uint32_t* tickets, uint64_t status_bytes_total, cudaStream_t st);
uint32_t radix_sort_tiles(uint64_t R);

constexpr int KNN_BOX = 1024;

struct MinMax {
    float3 minn, maxx;
};

// order-preserving float <-> uint mapping for atomicMin/Max
__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o)
{
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

__global__ void knn_init_bounds(uint32_t* b)
{
    // the reference reduces with init (0,0,0) for both min and max (simple_knn.cu:171-181)
    if (threadIdx.x < 6) b[threadIdx.x] = f2ord(0.f);
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, uint32_t* __restrict__ b)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&b[k], f2ord(mn[k]));
            atomicMax(&b[3 + k], f2ord(mx[k]));
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ b, uint64_t* __restrict__ keys,
                  uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float3 mn = {ord2f(b[0]), ord2f(b[1]), ord2f(b[2])};
    const float3 mx = {ord2f(b[3]), ord2f(b[4]), ord2f(b[5])};
    const float3 c = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
    const uint32_t x = prep_morton((uint32_t)(((c.x - mn.x) / (mx.x - mn.x)) * ((1 << 10) - 1)));
    const uint32_t y = prep_morton((uint32_t)(((c.y - mn.y) / (mx.y - mn.y)) * ((1 << 10) - 1)));
    const uint32_t z = prep_morton((uint32_t)(((c.z - mn.z) / (mx.z - mn.z)) * ((1 << 10) - 1)));
    keys[i] = (uint64_t)(x | (y << 1) | (z << 2));
    vals[i] = (uint32_t)i;
}

// sorted positions + box bounds (one CTA per 1024 consecutive sorted points)
__global__ void __launch_bounds__(KNN_BOX)
knn_box_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, float4* __restrict__ sorted,
               MinMax* __restrict__ boxes)
{
    __shared__ float s_red[6][KNN_BOX / 32];
    const int idx = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (idx < P) {
        const uint32_t o = order[idx];
        const float3 c = {pts[3 * (size_t)o], pts[3 * (size_t)o + 1], pts[3 * (size_t)o + 2]};
        sorted[idx] = make_float4(c.x, c.y, c.z, __uint_as_float(o));
        mn[0] = mx[0] = c.x; mn[1] = mx[1] = c.y; mn[2] = mx[2] = c.z;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_red[k][threadIdx.x >> 5] = mn[k];
            s_red[3 + k][threadIdx.x >> 5] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MinMax m;
        float r[6];
        for (int k = 0; k < 6; ++k) {
            r[k] = s_red[k][0];
            for (int w = 1; w < KNN_BOX / 32; ++w) r[k] = (k < 3) ? fminf(r[k], s_red[k][w]) : fmaxf(r[k], s_red[k][w]);
        }
        m.minn = make_float3(r[0], r[1], r[2]);
        m.maxx = make_float3(r[3], r[4], r[5]);
        boxes[blockIdx.x] = m;
    }
}

__device__ __forceinline__ float dist_box_point(const MinMax& box, const float3& p)
{
    float3 diff = {0, 0, 0};
    if (p.x < box.minn.x || p.x > box.maxx.x) diff.x = fminf(fabsf(p.x - box.minn.x), fabsf(p.x - box.maxx.x));
    if (p.y < box.minn.y || p.y > box.maxx.y) diff.y = fminf(fabsf(p.y - box.minn.y), fabsf(p.y - box.maxx.y));
    if (p.z < box.minn.z || p.z > box.maxx.z) diff.z = fminf(fabsf(p.z - box.minn.z), fabsf(p.z - box.maxx.z));
    return diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
}

__device__ __forceinline__ void update_k_best(const float3& ref, const float3& point, float* knn)
{
    const float3 d = {point.x - ref.x, point.y - ref.y, point.z - ref.z};
    float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            const float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}

// One CTA per box of query points; candidate boxes are staged through shared memory so the whole
// CTA scans them together (the reference has every thread stream every surviving box from global).
__global__ void __launch_bounds__(256)
knn_mean_dist_kernel(int P, int num_boxes, const float4* __restrict__ sorted, const MinMax* __restrict__ boxes,
                     float* __restrict__ dists)
{
    __shared__ float4 s_pts[KNN_BOX];
    const int qbox = blockIdx.x;
    // each thread owns 4 query points of this box
    float3 q[4];
    float best[4][3];
    float reject[4];
    int qi[4];
    uint32_t qo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        qi[k] = qbox * KNN_BOX + k * 256 + threadIdx.x;
        if (qi[k] < P) {
            const float4 t = sorted[qi[k]];
            q[k] = make_float3(t.x, t.y, t.z);
            qo[k] = __float_as_uint(t.w);
            float b3[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
            for (int i = max(0, qi[k] - 3); i <= min(P - 1, qi[k] + 3); i++) {
                if (i == qi[k]) continue;
                const float4 o = sorted[i];
                update_k_best(q[k], make_float3(o.x, o.y, o.z), b3);
            }
            reject[k] = b3[2];
        } else {
            q[k] = make_float3(0.f, 0.f, 0.f);
            qo[k] = 0;
            reject[k] = -1.f;
        }
        best[k][0] = best[k][1] = best[k][2] = FLT_MAX;
    }
    for (int b = 0; b < num_boxes; ++b) {
        const MinMax box = boxes[b];
        bool want[4];
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dist = dist_box_point(box, q[k]);
            want[k] = (qi[k] < P) && !(dist > reject[k] || dist > best[k][2]);
            any |= want[k];
        }
        // the vote is also the barrier that protects s_pts from the previous iteration's readers
        if (!__syncthreads_or(any ? 1 : 0)) continue;
        const int base = b * KNN_BOX;
        const int cnt = min(KNN_BOX, P - base);
        for (int i = threadIdx.x; i < cnt; i += 256) s_pts[i] = sorted[base + i];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!want[k]) continue;
            for (int i = 0; i < cnt; ++i) {
                if (base + i == qi[k]) continue;
                const float4 o = s_pts[i];
                update_k_best(q[k], make_float3(o.x, o.y, o.z), best[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (qi[k] < P) dists[qo[k]] = (best[k][0] + best[k][1] + best[k][2]) / 3.0f;
}

}  // namespace gigs

using namespace gigs;

extern "C" int gigs_dist2(int32_t P, const float* points, float* mean_dist2, void* scratch, uint64_t* scratch_bytes,
                          void* stream)
{
    if (P < 0 || !scratch_bytes) { set_error("gigs_dist2: bad arguments"); return -1; }
    const uint64_t n = (uint64_t)P;
    const uint32_t tiles = radix_sort_tiles(n);
    const int passes = 4;  // 30-bit Morton keys
    const uint32_t num_boxes = (uint32_t)((n + KNN_BOX - 1) / KNN_BOX);
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { o = align_up(o, 128); uint64_t r = o; o += bytes; return r; };
    const uint64_t o_bounds = take(32);
    const uint64_t o_keys_u = take(n * 8), o_vals_u = take(n * 4);
    const uint64_t o_keys_a = take(n * 8), o_vals_a = take(n * 4);
    const uint64_t o_keys_b = take(n * 8), o_vals_b = take(n * 4);
    const uint64_t o_hist = take(8 * 256 * 4), o_ticket = take(128);
    const uint64_t o_status = take((uint64_t)passes * tiles * 256 * 4);
    const uint64_t o_sorted = take(n * 16);
    const uint64_t o_boxes = take((uint64_t)num_boxes * sizeof(MinMax));
    const uint64_t need = align_up(o, 128) + 128;
    if (!scratch) { *scratch_bytes = need; return 0; }
    if (*scratch_bytes < need) { set_error("gigs_dist2: scratch too small"); return -2; }
    if (P == 0) return 0;
    if (!points || !mean_dist2) { set_error("gigs_dist2: bad arguments"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    char* s = (char*)scratch;
    uint32_t* bounds = (uint32_t*)(s + o_bounds);
    ProfScope ps(ST_DIST2, st);
    knn_init_bounds<<<1, 32, 0, st>>>(bounds);
    knn_bounds_kernel<<<min((P + 255) / 256, 148 * 8), 256, 0, st>>>(P, points, bounds);
    GIGS_LAUNCH_CHECK("knn_bounds_kernel");
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, points, bounds, (uint64_t*)(s + o_keys_u), (uint32_t*)(s + o_vals_u));
    GIGS_LAUNCH_CHECK("knn_morton_kernel");
    const uint64_t status_total = (o_status - o_hist) + (uint64_t)passes * tiles * 256 * 4;
    if (int e = launch_radix_sort(n, 30, (const uint64_t*)(s + o_keys_u), (const uint32_t*)(s + o_vals_u),
                                  (uint64_t*)(s + o_keys_a), (uint32_t*)(s + o_vals_a), (uint64_t*)(s + o_keys_b),
                                  (uint32_t*)(s + o_vals_b), (uint32_t*)(s + o_hist), (uint32_t*)(s + o_status),
                                  (uint32_t*)(s + o_ticket), status_total, st))
        return e;
    knn_box_kernel<<<num_boxes, KNN_BOX, 0, st>>>(P, points, (const uint32_t*)(s + o_vals_a), (float4*)(s + o_sorted),
                                                  (MinMax*)(s + o_boxes));
    GIGS_LAUNCH_CHECK("knn_box_kernel");
    knn_mean_dist_kernel<<<num_boxes, 256, 0, st>>>(P, (int)num_boxes, (const float4*)(s + o_sorted),
                                                    (const MinMax*)(s + o_boxes), mean_dist2);
    GIGS_LAUNCH_CHECK("knn_mean_dist_kernel");
    return 0;
}
