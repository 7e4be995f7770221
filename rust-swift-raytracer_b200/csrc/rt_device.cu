// rt_device.cu — device management and render orchestration (host side of the CUDA path).
//
//   World  --(pack, once)-->  scene blob  --(one H2D copy per device, once)-->  HBM
//   ray_trace_into():  [memset 16 B counters] -> ONE persistent render kernel -> [D2H frame]
//
// No CPU fallback: every failure becomes a std::runtime_error carrying the CUDA error text.
#include "rt_host.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

namespace rt {

// defined in rt_kernels_exact.cu / rt_kernels_fast.cu
cudaError_t launch_render_exact(const RtFrameParams&, const RtSceneView&, int grid, size_t smem_limit, cudaStream_t);
cudaError_t launch_render_fast(const RtFrameParams&, const RtSceneView&, int grid, size_t smem_limit, cudaStream_t);
cudaError_t occupancy_exact(size_t hot_bytes, size_t smem_limit, int* blocks_per_sm, int* block_size);
cudaError_t occupancy_fast(size_t hot_bytes, size_t smem_limit, int* blocks_per_sm, int* block_size);
cudaError_t launch_selftest_division(unsigned long long n_per_thread, uint32_t seed, int grid, int block,
                                     unsigned long long* d_mismatches, cudaStream_t);
cudaError_t launch_ffma_peak(float* out, int iters, int grid, int block, cudaStream_t);
double      ffma_peak_flops_per_launch(int iters, int grid, int block);

namespace {

[[noreturn]] void fail(const char* what, cudaError_t e)
{
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define RT_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) fail(#expr, e_); } while (0)

constexpr int kCounterSlots = 64;

struct CounterSlot {            // 16 B, zeroed by one memset per launch
    unsigned long long rays;
    unsigned int       work;
    unsigned int       pad;
};

struct DeviceContext {
    int          device   = -1;
    int          num_sms  = 0;
    size_t       smem_optin = 0;
    cudaStream_t stream   = nullptr;
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    CounterSlot* d_slots  = nullptr;
    CounterSlot* h_slot   = nullptr;   // pinned
    int          next_slot = 0;
    uint32_t*    d_out    = nullptr;   size_t d_out_cap = 0;     // internal frame buffer (pixels)
    unsigned char* h_stage = nullptr;  size_t h_stage_cap = 0;   // pinned staging for pageable destinations
    std::map<std::pair<size_t, int>, std::pair<int, int>> occupancy;   // (hot_bytes, fast) -> (CTAs per SM, CTA width)
};

std::mutex                    g_mutex;          // one render at a time per process (lib.rs is single-threaded)
std::map<int, DeviceContext*> g_contexts;

DeviceContext& context_for(int device)
{
    auto it = g_contexts.find(device);
    if (it != g_contexts.end()) return *it->second;
    auto* c = new DeviceContext();
    c->device = device;
    RT_CUDA(cudaSetDevice(device));
    int v = 0;
    RT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    c->num_sms = v;
    RT_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    c->smem_optin = (size_t)v;
    RT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RT_CUDA(cudaEventCreate(&c->ev0));
    RT_CUDA(cudaEventCreate(&c->ev1));
    RT_CUDA(cudaMalloc(&c->d_slots, kCounterSlots * sizeof(CounterSlot)));
    RT_CUDA(cudaMallocHost(&c->h_slot, sizeof(CounterSlot)));
    g_contexts[device] = c;
    return *c;
}

int resolve_device(int requested)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw std::runtime_error(std::string("no CUDA device available (this library has no CPU render path): ") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    int dev = requested;
    if (dev < 0) RT_CUDA(cudaGetDevice(&dev));
    if (dev >= n) throw std::runtime_error("CUDA device ordinal out of range");
    return dev;
}

}   // namespace

// Per-device copy of the packed scene.
struct DeviceScene {
    int            device = -1;
    unsigned char* blob   = nullptr;
    size_t         bytes  = 0;
    RtSceneView    view{};
    ~DeviceScene()
    {
        if (blob) {
            int cur = -1;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            cudaFree(blob);
            if (cur >= 0) cudaSetDevice(cur);
        }
    }
};

World::World()  = default;
World::~World() = default;

void World::invalidate_device()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    packed_.reset();
    device_.clear();
}

namespace {

// Scene upload (K1): one cudaMemcpy of the SoA blob, cached per (world, device).
const DeviceScene& device_scene(const World& w, DeviceContext& ctx)
{
    for (auto& s : w.device_)
        if (s->device == ctx.device) return *s;
    const World::Packed& p = w.packed();
    auto s    = std::make_unique<DeviceScene>();
    s->device = ctx.device;
    s->bytes  = p.blob.size();
    RT_CUDA(cudaMalloc(&s->blob, s->bytes));
    RT_CUDA(cudaMemcpyAsync(s->blob, p.blob.data(), s->bytes, cudaMemcpyHostToDevice, ctx.stream));
    RT_CUDA(cudaStreamSynchronize(ctx.stream));
    s->view = p.view(s->blob);
    w.device_.push_back(std::move(s));
    return *w.device_.back();
}

bool is_pinned_or_device_accessible(const void* p)
{
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

}   // namespace

void ray_trace_into(const World& world, const Camera& camera, size_t width, size_t height,
                    const Options& opt, ColorU8* host_pixels, void* device_pixels, void* device_accum,
                    void* user_stream)
{
    const auto t_begin = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lock(g_mutex);

    if (width < 1 || height < 1) throw std::runtime_error("framebuffer must be at least 1x1");
    if (width > 0x3fffffffu || height > 0x3fffffffu) throw std::runtime_error("framebuffer too large");
    if (opt.tile_rows < 4 || (opt.tile_rows & 3u)) throw std::runtime_error("tile_rows must be a positive multiple of 4");
    if (opt.shard_count < 1 || opt.shard_index >= opt.shard_count) throw std::runtime_error("bad shard index/count");
    if ((opt.accum_in || opt.accum_out) && !device_accum) throw std::runtime_error("accumulator requested but device_accum is null");

    const int dev = resolve_device(opt.device);
    RT_CUDA(cudaSetDevice(dev));
    DeviceContext&     ctx   = context_for(dev);
    const DeviceScene& scene = device_scene(world, ctx);
    cudaStream_t       stream = user_stream ? static_cast<cudaStream_t>(user_stream) : ctx.stream;

    const uint32_t W = (uint32_t)width, H = (uint32_t)height;
    const uint32_t n_tiles = shard_tile_count(H, opt.tile_rows, opt.shard_index, opt.shard_count);
    const uint64_t slots   = (uint64_t)n_tiles * ((W + 7u) / 8u) * (opt.tile_rows / 4u) * 32u;
    if (slots >= 0xffffff00ull) throw std::runtime_error("frame shard exceeds 2^32 pixel slots; use more shards");
    const bool   compact    = opt.shard_count > 1;
    const size_t out_pixels = compact ? (size_t)n_tiles * opt.tile_rows * W : (size_t)W * H;

    uint32_t* d_out = static_cast<uint32_t*>(device_pixels);
    if (!d_out && !opt.no_resolve) {
        if (ctx.d_out_cap < out_pixels) {
            if (ctx.d_out) RT_CUDA(cudaFree(ctx.d_out));
            ctx.d_out = nullptr; ctx.d_out_cap = 0;
            RT_CUDA(cudaMalloc(&ctx.d_out, out_pixels * sizeof(uint32_t)));
            ctx.d_out_cap = out_pixels;
        }
        d_out = ctx.d_out;
    }

    CounterSlot* slot = ctx.d_slots + (ctx.next_slot++ % kCounterSlots);

    RtFrameParams P{};
    P.camera       = camera.d;
    P.wm1          = (float)(W - 1u);      // `(width - 1) as f32`, common.rs:335
    P.hm1          = (float)(H - 1u);
    P.width        = W;
    P.height       = H;
    P.spp          = opt.samples_per_pixel;
    P.depth        = opt.max_ray_bounces;
    P.sample_begin = opt.sample_begin;
    P.resolve_spp  = opt.resolve_spp ? opt.resolve_spp : opt.sample_begin + opt.samples_per_pixel;
    P.seed         = opt.seed;
    P.flags        = (opt.fixed_jitter ? RT_FLAG_FIXED_JITTER : 0u) | (opt.accum_in ? RT_FLAG_ACCUM_IN : 0u) |
              (opt.accum_out ? RT_FLAG_ACCUM_OUT : 0u) | (opt.no_resolve ? RT_FLAG_NO_RESOLVE : 0u) |
              (compact ? RT_FLAG_COMPACT_OUT : 0u);
    P.tile_rows    = opt.tile_rows;
    P.tile_first   = opt.shard_index;
    P.tile_stride  = opt.shard_count;
    P.n_tiles      = n_tiles;
    P.out          = d_out;
    P.accum        = static_cast<RtFloat4*>(device_accum);
    P.ray_counter  = &slot->rays;
    P.work_counter = &slot->work;

    // launch geometry: persistent CTAs, resident-CTA count from the occupancy API
    const size_t hot_bytes  = (size_t)(scene.view.n_sph_pad + scene.view.n_tri_pad) * sizeof(RtFloat4);
    const size_t smem_limit = ctx.smem_optin > 1024 ? ctx.smem_optin - 1024 : 0;   // static smem: the mbarrier
    auto&        occ        = ctx.occupancy[{hot_bytes, opt.fast_math ? 1 : 0}];
    if (occ.first == 0) {
        RT_CUDA(opt.fast_math ? occupancy_fast(hot_bytes, smem_limit, &occ.first, &occ.second)
                              : occupancy_exact(hot_bytes, smem_limit, &occ.first, &occ.second));
        if (occ.first < 1) throw std::runtime_error("render kernel does not fit on this device");
    }
    const int      per_sm = occ.first, block = occ.second;
    const uint64_t want_ctas = (slots + (uint64_t)block - 1) / (uint64_t)block;
    int grid = (int)std::min<uint64_t>((uint64_t)per_sm * ctx.num_sms, std::max<uint64_t>(want_ctas, 1));
    // Work-queue granularity: a warp takes `reserve` pixel slots per atomicAdd.  Aim for >= 64
    // slabs per warp so that the last slab of the slowest warp is a small part of the frame.
    const uint64_t warps   = (uint64_t)grid * (uint64_t)(block / 32);
    uint64_t       reserve = slots / (warps * 64u) / 32u * 32u;
    P.reserve = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(reserve, 32u), 256u);

    if (n_tiles > 0) {
        RT_CUDA(cudaMemsetAsync(slot, 0, sizeof(CounterSlot), stream));
        if (opt.stats) RT_CUDA(cudaEventRecord(ctx.ev0, stream));
        RT_CUDA(opt.fast_math ? launch_render_fast(P, scene.view, grid, smem_limit, stream)
                              : launch_render_exact(P, scene.view, grid, smem_limit, stream));
        if (opt.stats) RT_CUDA(cudaEventRecord(ctx.ev1, stream));
    }

    if (host_pixels && !opt.no_resolve && n_tiles > 0) {
        // D2H of the finished RGBA8 rows.  Every tile is one contiguous byte range of the frame
        // (image.rs:27 row-major), so a shard copies tile by tile and a full frame in one piece.
        const bool   direct = is_pinned_or_device_accessible(host_pixels);
        const size_t tile_px = (size_t)opt.tile_rows * W;
        auto copy_range = [&](size_t dst_px, size_t src_px, size_t count) {
            if (direct) {
                RT_CUDA(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(host_pixels) + dst_px, d_out + src_px,
                                        count * 4, cudaMemcpyDeviceToHost, stream));
            } else {
                RT_CUDA(cudaMemcpyAsync(ctx.h_stage + src_px * 4, d_out + src_px, count * 4, cudaMemcpyDeviceToHost, stream));
            }
        };
        if (!direct && ctx.h_stage_cap < out_pixels * 4) {
            if (ctx.h_stage) RT_CUDA(cudaFreeHost(ctx.h_stage));
            ctx.h_stage = nullptr; ctx.h_stage_cap = 0;
            RT_CUDA(cudaMallocHost(&ctx.h_stage, out_pixels * 4));
            ctx.h_stage_cap = out_pixels * 4;
        }
        if (!compact) {
            copy_range(0, 0, (size_t)W * H);
            RT_CUDA(cudaStreamSynchronize(stream));
            if (!direct) std::memcpy(host_pixels, ctx.h_stage, (size_t)W * H * 4);
        } else {
            for (uint32_t j = 0; j < n_tiles; ++j) {
                const size_t tile  = (size_t)opt.shard_index + (size_t)j * opt.shard_count;
                const size_t first = tile * tile_px;
                const size_t count = std::min(tile_px, (size_t)W * H - first);
                copy_range(first, (size_t)j * tile_px, count);
            }
            RT_CUDA(cudaStreamSynchronize(stream));
            if (!direct)
                for (uint32_t j = 0; j < n_tiles; ++j) {
                    const size_t tile  = (size_t)opt.shard_index + (size_t)j * opt.shard_count;
                    const size_t first = tile * tile_px;
                    const size_t count = std::min(tile_px, (size_t)W * H - first);
                    std::memcpy(reinterpret_cast<uint32_t*>(host_pixels) + first, ctx.h_stage + (size_t)j * tile_px * 4, count * 4);
                }
        }
    } else if (!user_stream) {
        RT_CUDA(cudaStreamSynchronize(stream));
    }

    if (opt.stats) {
        RenderStats& st = *opt.stats;
        st = RenderStats{};
        st.grid       = (uint32_t)grid;
        st.block      = (uint32_t)block;
        st.resident   = hot_bytes <= smem_limit ? 1u : 0u;
        st.smem_bytes = st.resident ? (uint32_t)hot_bytes : 0u;
        if (n_tiles > 0) {
            RT_CUDA(cudaMemcpyAsync(ctx.h_slot, slot, sizeof(CounterSlot), cudaMemcpyDeviceToHost, stream));
            RT_CUDA(cudaStreamSynchronize(stream));
            RT_CUDA(cudaEventElapsedTime(&st.kernel_ms, ctx.ev0, ctx.ev1));
            st.rays     = ctx.h_slot->rays;
            st.launches = 1;
            // samples actually traced by this shard
            uint64_t px = 0;
            for (uint32_t j = 0; j < n_tiles; ++j) {
                const uint64_t tile = (uint64_t)opt.shard_index + (uint64_t)j * opt.shard_count;
                const uint64_t r0   = tile * opt.tile_rows;
                const uint64_t r1   = std::min<uint64_t>(r0 + opt.tile_rows, H);
                px += (r1 - r0) * W;
            }
            st.samples = (opt.samples_per_pixel > 0 && opt.max_ray_bounces > 0) ? px * (uint64_t)opt.samples_per_pixel : 0;
        }
        st.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
}

Framebuffer ray_trace(const World& world, const Camera& camera, Framebuffer framebuffer, Options& options)
{
    if (framebuffer.pixels.size() != framebuffer.width * framebuffer.height)
        framebuffer.pixels.resize(framebuffer.width * framebuffer.height, ColorU8{0, 0, 0, 0});
    Options o     = options;
    o.shard_index = 0;
    o.shard_count = 1;
    o.accum_in = o.accum_out = o.no_resolve = false;
    ray_trace_into(world, camera, framebuffer.width, framebuffer.height, o, framebuffer.pixels.data(), nullptr,
                   nullptr, nullptr);
    return framebuffer;
}

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

double measure_fp32_peak_tflops(int device, float* sm_clock_mhz_out)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    const int dev = resolve_device(device);
    RT_CUDA(cudaSetDevice(dev));
    DeviceContext& ctx = context_for(dev);
    const int block = 256, grid = ctx.num_sms * 8, iters = 16384;
    float* d = nullptr;
    RT_CUDA(cudaMalloc(&d, (size_t)grid * block * sizeof(float)));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        RT_CUDA(cudaEventRecord(ctx.ev0, ctx.stream));
        RT_CUDA(launch_ffma_peak(d, iters, grid, block, ctx.stream));
        RT_CUDA(cudaEventRecord(ctx.ev1, ctx.stream));
        RT_CUDA(cudaStreamSynchronize(ctx.stream));
        float ms = 0.f;
        RT_CUDA(cudaEventElapsedTime(&ms, ctx.ev0, ctx.ev1));
        if (rep > 0) best = std::max(best, ffma_peak_flops_per_launch(iters, grid, block) / (ms * 1e-3) / 1e12);
    }
    RT_CUDA(cudaFree(d));
    if (sm_clock_mhz_out) {
        int khz = 0;
        RT_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
        *sm_clock_mhz_out = khz / 1000.0f;
    }
    return best;
}

long long selftest_division(int device, unsigned long long operand_sets, uint32_t seed)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    const int dev = resolve_device(device);
    RT_CUDA(cudaSetDevice(dev));
    DeviceContext& ctx = context_for(dev);
    const int block = 256, grid = ctx.num_sms * 4;
    const unsigned long long per_thread = (operand_sets + (unsigned long long)grid * block - 1) / ((unsigned long long)grid * block);
    unsigned long long* d = nullptr;
    RT_CUDA(cudaMalloc(&d, sizeof *d));
    RT_CUDA(cudaMemsetAsync(d, 0, sizeof *d, ctx.stream));
    RT_CUDA(launch_selftest_division(per_thread, seed, grid, block, d, ctx.stream));
    unsigned long long h = 0;
    RT_CUDA(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, ctx.stream));
    RT_CUDA(cudaStreamSynchronize(ctx.stream));
    RT_CUDA(cudaFree(d));
    return (long long)h;
}

void* alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void free_pinned(void* p) { if (p) cudaFreeHost(p); }

}   // namespace rt
