// hostsim.cpp — TEST-ONLY: compiles the device header csrc/rt_trace.cuh (exact policy) with
// g++ and drives it with a plain per-pixel loop, so that the shading / intersection / RNG
// logic of the CUDA kernel can be checked against the oracle on a machine without a GPU.
// It is NOT part of the product and is never loaded by the package: the shipped library has
// no CPU render path.  Only tests/test_hostsim.py loads it.
#define RT_TU_EXACT 1
#include "../../rust-swift-raytracer_b200/csrc/rt_trace.cuh"
#include "../../rust-swift-raytracer_b200/csrc/rt_host.hpp"

#include <cstring>

namespace rt {
struct DeviceScene {};
World::World()  = default;
World::~World() = default;
void World::invalidate_device() { packed_.reset(); }
}   // namespace rt

extern "C" int hostsim_render(const char* scene_text, const float cam12[12], uint32_t W, uint32_t H, int32_t spp,
                              int32_t depth, uint32_t seed, uint32_t flags, int32_t sample_begin,
                              int32_t resolve_spp, uint8_t* out, uint64_t* rays_out)
{
    using namespace rt;
    ParseResult pr = parse_input(scene_text, std::strlen(scene_text));
    if (pr.error != ParseError::Ok) return (int)pr.error;
    const World::Packed& pk = pr.world->packed();
    RtSceneView G = pk.view(pk.blob.data());

    RtFrameParams P{};
    std::memcpy(&P.camera, cam12, sizeof(float) * 12);
    P.width = W; P.height = H; P.spp = spp; P.depth = depth; P.seed = seed; P.flags = flags & 0x3fffffffu;   // bits 31/30 select the sphere-walk variant (below)
    P.wm1 = (float)(W - 1u); P.hm1 = (float)(H - 1u);
    P.sample_begin = sample_begin;
    P.resolve_spp  = resolve_spp ? resolve_spp : sample_begin + spp;

    CullView cv{G.cull_bound, G.cull_sph, G.cull_r2, G.cull_orig, G.n_groups};
    if ((flags & 0x40000000u) && G.n_groups == 0) return 100;      // the world has no block C (< 64 spheres)
    uint64_t rays = 0;
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    const bool trace = spp > 0 && depth > 0;
    for (uint32_t image_row = 0; image_row < H; ++image_row)
        for (uint32_t column = 0; column < W; ++column) {
            Lane L{};
            begin_pixel(L, P, column, H - 1u - image_row, image_row * W + column);
            // the kernel's loop for one lane: one ray segment per iteration until the pixel is complete
            // the kernel's loop for one lane; bit 31 of `flags` selects the FILTER variant, bit 30 the CULL variant
            while (trace && L.sample < spp) {
                ++rays;
                if (flags & 0x40000000u)      trace_segment<false, RT_SPH_CULL, true>(L, P, G, G.sph_filter, G.sph_r2, cv, G.tri_plane);
                else if (flags & 0x80000000u) trace_segment<false, RT_SPH_FILTER, true>(L, P, G, G.sph_filter, G.sph_r2, cv, G.tri_plane);
                else if (G.n_tri_pad)         trace_segment<false, RT_SPH_DIRECT, true>(L, P, G, G.sph, nullptr, cv, G.tri_plane);
                else                          trace_segment<false, RT_SPH_DIRECT, false>(L, P, G, G.sph, nullptr, cv, G.tri_plane);
            }
            out32[L.out_index] = resolve_pixel<false>(L.acc_r, L.acc_g, L.acc_b, pixel_alpha(1.0f, spp > 0 ? spp : 0),
                                                      P.resolve_spp);
        }
    if (rays_out) *rays_out = rays;
    return 0;
}
