// C-ABI entry points for the rasterizer (include/gigs_b200.h): workspace layout, forward
// orchestration (reference cuda_rasterizer/rasterizer_impl.cu:486-672) and backward orchestration
// (:676-803). No torch types; the caller owns every buffer.
#include <cstdarg>
#include <cstring>
#include <map>
#include <atomic>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include "common.cuh"

namespace gigs {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

static std::mutex g_attr_mutex;
static std::map<std::pair<int, const void*>, int> g_attr_set;  // (device, kernel) -> bytes already granted
int ensure_dynamic_smem(const void* kernel, int bytes)
{
    int dev = 0;
    GIGS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    auto key = std::make_pair(dev, kernel);
    auto it = g_attr_set.find(key);
    if (it != g_attr_set.end() && it->second >= bytes) return 0;
    GIGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_attr_set[key] = bytes;
    return 0;
}

// ---- programmatic dependent launch (common.cuh: launch_k) -------------------------------------------
static int g_pdl = -1;   // -1: not decided yet (env GIGS_PDL, default on)
static std::atomic<uint64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled()
{
    if (g_pdl < 0) {
        const char* e = getenv("GIGS_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl != 0;
}

// ---- optional per-stage event timing -------------------------------------------------------------
struct ProfRec { int stage; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
int prof_begin(int stage, cudaStream_t st)
{
    if (!g_prof_on || stage < 0) return -1;
    ProfRec r;
    r.stage = stage;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int token, cudaStream_t st)
{
    if (token < 0 || token >= (int)g_prof.size()) return;
    cudaEventRecord(g_prof[token].e1, st);
}

// reference rasterizer_impl.cu:35-50 (number of key bits needed for the tile id)
uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb)
            msb += step;
        else
            msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

uint32_t radix_sort_tiles(uint64_t R);

Layout make_layout(int P, int W, int H, uint64_t R)
{
    Layout L;
    memset(&L, 0, sizeof(L));
    const uint64_t Pn = (uint64_t)(P > 0 ? P : 0);
    const uint64_t N = (uint64_t)W * (uint64_t)H;
    L.tiles_x = (W + TILE_X - 1) / TILE_X;
    L.tiles_y = (H + TILE_Y - 1) / TILE_Y;
    L.num_tiles = L.tiles_x * L.tiles_y;
    L.num_blocks = (uint32_t)((Pn + 255) / 256);

    uint64_t o = 0;
    auto take = [&](uint64_t bytes) {
        o = align_up(o, 128);
        uint64_t r = o;
        o += bytes;
        return r;
    };
    // geom
    L.off.g_record = take(Pn * REC_FLOATS * 4);
    L.off.g_cov3D = take(Pn * 6 * 4);
    L.off.g_clamped = take(Pn * 4);
    L.off.g_tiles_touched = take(Pn * 4);
    L.off.g_depth_keys = take(Pn * 4);
    L.off.g_order = take(Pn * 4);
    L.off.g_block_sums = take(((uint64_t)L.num_blocks + 1) * 4);
    L.g_block_sums2 = take(((uint64_t)L.num_blocks + 1) * 4);
    L.off.g_num_rendered = take(16);
    {   // depth argsort of the Gaussians: 32-bit keys, 4 passes
        const uint32_t ptiles = radix_sort_tiles(Pn);
        L.p_keys_a = take(Pn * 4);
        L.p_keys_b = take(Pn * 4);
        L.p_vals_b = take(Pn * 4);
        L.p_hist = take(8 * 256 * 4);
        L.p_ticket = take(128);
        L.p_status = take((uint64_t)4 * ptiles * 256 * 4);
        L.p_zero_bytes = (L.p_status - L.p_hist) + (uint64_t)4 * ptiles * 256 * 4;
    }
    L.size.geom_bytes = align_up(o, 128) + 128;
    // img
    o = 0;
    L.off.i_final_T = take(N * 4);
    L.off.i_n_contrib = take(N * 4);
    L.off.i_ranges = take((uint64_t)L.num_tiles * 8);
    L.size.img_bytes = align_up(o, 128) + 128;
    // binning (kept for backward): sorted Gaussian ids + the forward's footprint-test masks
    o = 0;
    L.off.b_point_list = take(R * 4);
    L.b_warp_masks = take(((R >> 5) + (uint64_t)L.num_tiles + 1) * WARP_MASK_WORDS * 4);
    L.size.binning_bytes = align_up(o, 128) + 128;
    // instance sort scratch: (tile id, Gaussian id) pairs, tile-id bits only
    o = 0;
    L.sort_bits = higher_msb(L.num_tiles);
    L.sort_passes = (uint32_t)radix_sort_passes((int)L.sort_bits);
    L.sort_tiles = radix_sort_tiles(R);
    L.off.s_tiles_unsorted = take(R * 4);
    L.off.s_vals_unsorted = take(R * 4);
    L.s_keys_a = take(R * 4);
    L.s_keys_b = take(R * 4);
    L.s_vals_b = take(R * 4);
    L.s_hist = take(8 * 256 * 4);
    L.s_joint = take(radix_joint_bytes());     // joint tile-id histogram (digit histograms + tile ranges come from it)
    L.s_ticket = take(128);
    L.s_status = take((uint64_t)L.sort_passes * L.sort_tiles * 256 * 4);
    L.s_zero_bytes = (L.s_status - L.s_hist) + (uint64_t)L.sort_passes * L.sort_tiles * 256 * 4;
    L.off.s_tiles_sorted = L.s_keys_a;
    L.size.sort_bytes = align_up(o, 128) + 128;
    return L;
}

// launchers defined in the kernel files
int launch_preprocess(const GigsRasterFwd* a, const Layout& L, cudaStream_t st, const float* sh_rest);
int launch_depth_argsort(const GigsRasterFwd* a, const Layout& L, cudaStream_t st);
int launch_emit_keys(const GigsRasterFwd* a, const Layout& L, uint32_t* keys, uint32_t* vals, void* clear,
                     uint64_t clear_bytes, int* cleared, cudaStream_t st);
int launch_tile_ranges(uint64_t R, const uint32_t* tiles_sorted, uint2* ranges, uint32_t num_tiles, cudaStream_t st);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t st);
int launch_tile_sort(uint64_t R, int end_bit, const uint32_t* keys_u, const uint32_t* vals_u, uint32_t* keys_a,
                     uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                     uint32_t* tickets, uint64_t zero_bytes, int pass_stage, uint32_t* joint, uint2* ranges,
                     uint32_t num_tiles, int* ranges_done, cudaStream_t st);
int launch_radix_sort32(uint64_t R, int end_bit, const uint32_t* keys_u, const uint32_t* vals_u, uint32_t* keys_a,
                        uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                        uint32_t* tickets, uint64_t zero_bytes, int pass_stage, cudaStream_t st);
int launch_blend_forward(const GigsRasterFwd* a, const Layout& L, bool lite, cudaStream_t st);
int launch_blend_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st);
int launch_gaussian_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st);

static int check_common(int P, const GigsCamera& c)
{
    if (P < 0) { set_error("P must be >= 0"); return -1; }
    if (c.width <= 0 || c.height <= 0) { set_error("image size must be positive"); return -1; }
    if (!c.viewmatrix || !c.projmatrix || !c.campos || !c.bg) { set_error("camera pointers must not be NULL"); return -1; }
    return 0;
}

// The page-locked word num_rendered is delivered in: the caller's (GigsRasterFwd.pinned_num_rendered) or one of this
// thread's own. dev = its device alias when the word is mapped into the device's address space (page-locked memory
// is, under unified addressing): the preprocess scan kernel then stores the total straight into host memory.
int host_total_slot(const GigsRasterFwd* a, HostSlot* s)
{
    static thread_local uint32_t* pinned = nullptr;
    if (!a->pinned_num_rendered && !pinned) GIGS_CUDA(cudaMallocHost((void**)&pinned, 64));
    s->host = a->pinned_num_rendered ? a->pinned_num_rendered : pinned;
    void* d = nullptr;
    static const bool no_map = getenv("GIGS_NO_MAPPED_READBACK") != nullptr;
    if (no_map || cudaHostGetDevicePointer(&d, s->host, 0) != cudaSuccess) {
        cudaGetLastError();
        d = nullptr;
    }
    s->dev = (uint32_t*)d;
    return 0;
}

// num_rendered to the host. The depth argsort of the Gaussians (which does not depend on num_rendered) is queued
// first, so the GPU keeps working while the host sizes the binning blob. Mapped word: the host polls it (the scan
// kernel stored the total there; launch_preprocess armed it with NUM_RENDERED_PENDING) - no copy or event node in
// the stream. Otherwise: a 4-byte copy and an event, waiting for THAT copy only.
int read_back_num_rendered(GigsRasterFwd* a, const Layout& L, cudaStream_t st)
{
    HostSlot slot;
    if (int e = host_total_slot(a, &slot)) return e;
    if (slot.dev) {
        {
            ProfScope ps(ST_DEPTH_SORT, st);
            if (int e = launch_depth_argsort(a, L, st)) return e;
        }
        volatile uint32_t* w = slot.host;
        uint32_t v = *w;
        for (uint32_t spins = 1; v == NUM_RENDERED_PENDING; ++spins) {
            if ((spins & 0x3fffu) == 0) {   // a failed launch / sticky error must not spin for ever
                const cudaError_t q = cudaStreamQuery(st);
                if (q != cudaSuccess && q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery (num_rendered)");
                if (q == cudaSuccess && *w == NUM_RENDERED_PENDING) {
                    set_error("num_rendered was not delivered although the stream is idle");
                    return -1;
                }
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            v = *w;
        }
        a->num_rendered = (int64_t)v;
        return 0;
    }
    // an event belongs to the device that was current when it was created: one per (thread, device)
    constexpr int MAX_DEV = 64;
    static thread_local cudaEvent_t evs[MAX_DEV] = {};
    int dev = 0;
    GIGS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) { set_error("device index %d out of range", dev); return -1; }
    if (!evs[dev]) GIGS_CUDA(cudaEventCreateWithFlags(&evs[dev], cudaEventDisableTiming));
    cudaEvent_t ev = evs[dev];
    GIGS_CUDA(cudaMemcpyAsync(slot.host, (char*)a->geom + L.off.g_num_rendered, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GIGS_CUDA(cudaEventRecord(ev, st));
    {
        ProfScope ps(ST_DEPTH_SORT, st);
        if (int e = launch_depth_argsort(a, L, st)) return e;
    }
    GIGS_CUDA(cudaEventSynchronize(ev));
    a->num_rendered = (int64_t)*slot.host;
    return 0;
}

int forward_finish_impl(GigsRasterFwd* a, bool lite)
{
    if (!a) { set_error("null args"); return -1; }
    if (int e = check_common(a->P, a->cam)) return e;
    cudaStream_t st = (cudaStream_t)a->stream;
    if (a->P == 0) {
        // the reference short-circuits to its zero-filled outputs (rasterize_points.cu:191)
        const size_t N = (size_t)a->cam.width * a->cam.height * sizeof(float);
        float* three[] = {a->out_color, lite ? nullptr : a->out_normal, lite ? nullptr : a->out_normal_view,
                          lite ? nullptr : a->out_pos, lite ? nullptr : a->out_albedo};
        float* one[] = {a->out_opacity, a->out_depth, lite ? nullptr : a->out_roughness,
                        lite ? nullptr : a->out_metallic};
        for (float* p : three)
            if (p) GIGS_CUDA(cudaMemsetAsync(p, 0, 3 * N, st));
        for (float* p : one)
            if (p) GIGS_CUDA(cudaMemsetAsync(p, 0, N, st));
        return 0;
    }
    const uint64_t R = (uint64_t)a->num_rendered;
    const Layout L = make_layout(a->P, a->cam.width, a->cam.height, R);
    if (a->geom_bytes < L.size.geom_bytes || a->img_bytes < L.size.img_bytes ||
        a->binning_bytes < L.size.binning_bytes || a->sort_bytes < L.size.sort_bytes) {
        set_error("workspace too small (geom %llu/%llu img %llu/%llu binning %llu/%llu sort %llu/%llu)",
                  (unsigned long long)a->geom_bytes, (unsigned long long)L.size.geom_bytes,
                  (unsigned long long)a->img_bytes, (unsigned long long)L.size.img_bytes,
                  (unsigned long long)a->binning_bytes, (unsigned long long)L.size.binning_bytes,
                  (unsigned long long)a->sort_bytes, (unsigned long long)L.size.sort_bytes);
        return -2;
    }
    char* sc = (char*)a->sort;
    char* im = (char*)a->img;
    char* bn = (char*)a->binning;
    uint32_t* keys_u = (uint32_t*)(sc + L.off.s_tiles_unsorted);
    uint32_t* vals_u = (uint32_t*)(sc + L.off.s_vals_unsorted);
    int ranges_done = 0, sort_cleared = 0;
    if (a->P > 0 && R > 0) {
        {
            ProfScope ps(ST_EMIT_KEYS, st);
            if (int e = launch_emit_keys(a, L, keys_u, vals_u, sc + L.s_hist, L.s_zero_bytes, &sort_cleared, st)) return e;
        }
        ProfScope ps(ST_SORT, st);
        if (int e = launch_tile_sort(R, (int)L.sort_bits, keys_u, vals_u, (uint32_t*)(sc + L.s_keys_a),
                                     (uint32_t*)(bn + L.off.b_point_list), (uint32_t*)(sc + L.s_keys_b),
                                     (uint32_t*)(sc + L.s_vals_b), (uint32_t*)(sc + L.s_hist),
                                     (uint32_t*)(sc + L.s_status), (uint32_t*)(sc + L.s_ticket),
                                     sort_cleared ? 0 : L.s_zero_bytes,   // 0: the emit kernel cleared them
                                     ST_SORT_PASS, (uint32_t*)(sc + L.s_joint), (uint2*)(im + L.off.i_ranges), L.num_tiles,
                                     &ranges_done, st))
            return e;
    }
    if (!ranges_done) {
        ProfScope ps(ST_RANGES, st);
        if (int e = launch_tile_ranges(R, (const uint32_t*)(sc + L.s_keys_a), (uint2*)(im + L.off.i_ranges), L.num_tiles, st))
            return e;
    }
    {
        ProfScope ps(ST_BLEND_FWD, st);
        if (int e = launch_blend_forward(a, L, lite, st)) return e;
    }
    if (a->cam.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_abi_version(void) { return GIGS_ABI_VERSION; }

uint64_t gigs_launch_count(void) { return gigs::g_launches.load(std::memory_order_relaxed); }

int gigs_set_dependent_launch(int32_t on)
{
    const int prev = gigs::pdl_enabled() ? 1 : 0;
    if (on >= 0) gigs::g_pdl = on ? 1 : 0;
    return prev;
}

int gigs_sizeof(int32_t which)
{
    switch (which) {
        case 0: return (int)sizeof(GigsCamera);
        case 1: return (int)sizeof(GigsSizes);
        case 2: return (int)sizeof(GigsLayout);
        case 3: return (int)sizeof(GigsRasterFwd);
        case 4: return (int)sizeof(GigsRasterBwd);
        case 5: return (int)sizeof(GigsShade);
        case 6: return (int)sizeof(GigsFrameLayout);
        case 7: return (int)sizeof(GigsFrame);
        case 8: return (int)sizeof(GigsLightLayout);
        case 9: return (int)sizeof(GigsAdamGroup);
        case 10: return (int)sizeof(GigsDensifyGroup);
        case 11: return (int)sizeof(GigsStage1Layout);
        case 12: return (int)sizeof(GigsStage1);
        default: return -1;
    }
}
const char* gigs_last_error(void) { return g_last_error.c_str(); }

int gigs_profile_enable(int32_t on)
{
    g_prof_on = on != 0;
    return 0;
}

int gigs_profile_read(int32_t* stages, float* ms, int32_t cap)
{
    int n = 0;
    for (auto& r : g_prof) {
        float t = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            if (n < cap && stages && ms) {
                stages[n] = r.stage;
                ms[n] = t;
                ++n;
            }
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    return n;
}

int gigs_raster_sizes(int32_t P, int32_t W, int32_t H, uint64_t R, GigsSizes* out)
{
    if (!out || P < 0 || W <= 0 || H <= 0) { set_error("gigs_raster_sizes: bad arguments"); return -1; }
    *out = make_layout(P, W, H, R).size;
    return 0;
}

int gigs_raster_layout(int32_t P, int32_t W, int32_t H, uint64_t R, GigsLayout* out)
{
    if (!out || P < 0 || W <= 0 || H <= 0) { set_error("gigs_raster_layout: bad arguments"); return -1; }
    *out = make_layout(P, W, H, R).off;
    return 0;
}

int gigs_raster_forward_begin(GigsRasterFwd* a)
{
    if (!a) { set_error("null args"); return -1; }
    if (int e = check_common(a->P, a->cam)) return e;
    a->num_rendered = 0;
    if (a->P == 0) return 0;
    if (!a->means3D || !a->opacities || !a->radii) { set_error("means3D/opacities/radii must not be NULL"); return -1; }
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr)) {
        set_error("provide exactly one of shs / colors_precomp");
        return -1;
    }
    if (a->cov3D_precomp == nullptr && (a->scales == nullptr || a->rotations == nullptr)) {
        set_error("provide scales+rotations or cov3D_precomp");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)a->stream;
    const Layout L = make_layout(a->P, a->cam.width, a->cam.height, 0);
    if (a->geom_bytes < L.size.geom_bytes || a->img_bytes < L.size.img_bytes) {
        set_error("geom/img workspace too small");
        return -2;
    }
    {
        ProfScope ps(ST_PREPROCESS, st);
        if (int e = launch_preprocess(a, L, st, nullptr)) return e;
    }
    if (int e = read_back_num_rendered(a, L, st)) return e;
    return 0;
}

int gigs_raster_forward_finish(GigsRasterFwd* a) { return forward_finish_impl(a, false); }
int gigs_lite_forward_finish(GigsRasterFwd* a) { return forward_finish_impl(a, true); }

int gigs_raster_backward(GigsRasterBwd* a)
{
    if (!a) { set_error("null args"); return -1; }
    if (int e = check_common(a->P, a->cam)) return e;
    if (a->P == 0) return 0;
    if (!a->accum || !a->dL_dmean2D || !a->dL_dopacity || !a->dL_dcolor || !a->dL_dnormal || !a->dL_dalbedo ||
        !a->dL_droughness || !a->dL_dmetallic || !a->dL_dmean3D || !a->dL_dcov3D) {
        set_error("backward: required output pointer is NULL");
        return -1;
    }
    if ((a->shs != nullptr) && !a->dL_dsh) { set_error("backward: dL_dsh required with shs"); return -1; }
    if ((a->scales != nullptr) && (!a->dL_dscale || !a->dL_drot || !a->rotations)) {
        set_error("backward: dL_dscale/dL_drot required with scales");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)a->stream;
    const Layout L = make_layout(a->P, a->cam.width, a->cam.height, (uint64_t)a->num_rendered);
    {
        ProfScope ps(ST_BLEND_BWD, st);
        GIGS_CUDA(cudaMemsetAsync(a->accum, 0, (size_t)a->P * ACC_FLOATS * sizeof(float), st));
        if (a->num_rendered > 0) {
            if (int e = launch_blend_backward(a, L, st)) return e;
        }
    }
    {
        ProfScope ps(ST_GAUSS_BWD, st);
        if (int e = launch_gaussian_backward(a, L, st)) return e;
    }
    if (a->cam.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int gigs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present, void* stream)
{
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("gigs_mark_visible: bad arguments"); return -1; }
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

}  // extern "C"
