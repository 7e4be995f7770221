// hostsim.cpp — TEST-ONLY: compiles the device header csrc/rt_trace.cuh (exact policy) with
// g++ and drives it with a plain per-pixel loop, so that the shading / intersection / RNG
// logic of the CUDA kernel can be checked against the oracle on a machine without a GPU.
// It is NOT part of the product and is never loaded by the package: the shipped library has
// no CPU render path.  Only tests/test_hostsim.py loads it.
#define RT_TU_EXACT 1
#include "../../rust-swift-raytracer_b200/csrc/rt_trace.cuh"
#include "../../rust-swift-raytracer_b200/csrc/rt_host.hpp"

#include <cmath>
#include <cstring>
#include <vector>

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
    P.width = W; P.height = H; P.spp = spp; P.depth = depth; P.seed = seed; P.flags = flags & 0x1fffffffu;   // bits 31/30/29 select the sphere-walk variant (below)
    P.wm1 = (float)(W - 1u); P.hm1 = (float)(H - 1u);
    P.sample_begin = sample_begin;
    P.one = 1.0f;
    P.resolve_spp  = resolve_spp ? resolve_spp : sample_begin + spp;

    CullView cv{G.cull_bound, G.cull_sph, G.cull_r2, G.cull_orig, G.n_groups};
    if ((flags & 0x40000000u) && G.n_groups == 0) return 100;      // the world has no block C (< 64 spheres)
    uint64_t rays = 0;
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    const bool trace = spp > 0 && depth > 0;
    if (flags & 0x20000000u) {
        // the kernel's NP = 2 loop (rt_kernels.cuh step 2): a lane carries two pixels and walks the lists once
        // for both rays; when one pixel is complete its path idles with a NaN direction
        const uint32_t n_px = W * H;
        for (uint32_t i = 0; i < n_px; i += 2) {
            Lane L[2] = {};
            bool have[2];
            for (int p = 0; p < 2; ++p) {
                const uint32_t idx = i + (uint32_t)p;
                have[p] = idx < n_px;
                if (have[p]) begin_pixel(L[p], P, idx % W, H - 1u - idx / W, idx);
            }
            for (;;) {
                V3   o[2], d[2];
                Hit  h[2];
                bool live[2];
                for (int p = 0; p < 2; ++p) {
                    live[p] = have[p] && trace && L[p].sample < spp;
                    o[p] = mk(0.f, 0.f, 0.f);
                    d[p] = mk(NAN, NAN, NAN);
                    if (live[p]) { d[p] = segment_begin<false>(L[p], P); o[p] = L[p].o; ++rays; }
                }
                if (!live[0] && !live[1]) break;
                if (G.n_tri_pad) closest_hit_n<false, true, 2>(G.sph_filter, G.sph_r2, G.n_sph, G.n_sph_pad, G.tri_plane, G.tri_cull, G.tri_v, G.n_tri_pad, o, d, P.one, h);
                else             closest_hit_n<false, false, 2>(G.sph_filter, G.sph_r2, G.n_sph, G.n_sph_pad, G.tri_plane, G.tri_cull, G.tri_v, G.n_tri_pad, o, d, P.one, h);
                for (int p = 0; p < 2; ++p)
                    if (live[p]) segment_end<false, RT_SPH_FILTER, true>(L[p], G, G.sph_filter, d[p], h[p]);
            }
            for (int p = 0; p < 2; ++p)
                if (have[p])
                    out32[L[p].out_index] = resolve_pixel<false>(L[p].acc_r, L[p].acc_g, L[p].acc_b,
                                                                 pixel_alpha(1.0f, spp > 0 ? spp : 0), P.resolve_spp);
        }
        if (rays_out) *rays_out = rays;
        return 0;
    }
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

// TEST-ONLY: structural invariants of block C (rt_scene.cpp build_cull_block), independent of any render:
//  * every sphere of the list appears exactly once among the group members (cull_orig);
//  * a member's filter record and r*r are the list's own;
//  * the stored bound constants dominate  A = |c_k - cB| + r_k + B(|c_k| + r_k)  for every member,
//    where A is recovered from gB = 2AB(1+m)  (groups with wB = -inf always pass and are exempt).
// Returns 0 when all hold, otherwise a code; *n_groups_out / *n_always_out describe the block.
extern "C" int hostsim_check_cull(const char* scene_text, uint32_t* n_groups_out, uint32_t* n_always_out)
{
    using namespace rt;
    ParseResult pr = parse_input(scene_text, std::strlen(scene_text));
    if (pr.error != ParseError::Ok) return (int)pr.error;
    const World::Packed& pk = pr.world->packed();
    RtSceneView G = pk.view(pk.blob.data());
    if (n_groups_out) *n_groups_out = G.n_groups;
    if (G.n_groups == 0) return G.n_sph >= RT_FILTER_FROM ? 101 : 0;
    if (G.n_groups % 32u) return 102;
    std::vector<uint32_t> seen(G.n_sph, 0);
    uint32_t always = 0;
    for (uint32_t g = 0; g < G.n_groups; ++g) {
        const RtFloat4 b  = G.cull_bound[g];
        const float    gb = G.cull_sph[9u * g + 8u].x;
        const bool pass_always = std::isinf(b.w) && b.w < 0;
        if (pass_always) ++always;
        const double A = (double)gb / (2.0 * (double)RT_CULL_B * (1.0 + (double)RT_CULL_M));
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t idx = G.cull_orig[8u * g + k];
            const RtFloat4 s   = G.cull_sph[9u * g + k];
            if (idx == 0xffffffffu) { if (s.x == s.x) return 103; continue; }      // padding must be NaN
            if (idx >= G.n_sph) return 104;
            ++seen[idx];
            {   // block B stores the same records in pairs: {x0, x1, y0, y1} {z0, z1, -w0, -w1}
                const float* f = reinterpret_cast<const float*>(G.sph_filter + (idx & ~1u));
                const uint32_t h = idx & 1u;
                const RtFloat4 rec = {f[h], f[2 + h], f[4 + h], -f[6 + h]};
                if (std::memcmp(&s, &rec, sizeof s) != 0) return 105;
            }
            if (std::memcmp(&G.cull_r2[8u * g + k], &G.sph_r2[idx], sizeof(float)) != 0) return 106;
            if (pass_always) continue;
            const double c[3] = {s.x, s.y, s.z}, r = std::sqrt((double)G.sph_r2[idx]);
            const double dx = c[0] - b.x, dy = c[1] - b.y, dz = c[2] - b.z;
            const double need = std::sqrt(dx * dx + dy * dy + dz * dz) + r +
                                (double)RT_CULL_B * (std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) + r);
            if (!(A >= need * (1.0 - 1e-9))) return 107;
            // wB = cB.cB - A^2 - m(cB.cB + A^2), rounded down: must not exceed the exact value
            const double ccb = (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z;
            if (!((double)b.w <= ccb - A * A * (1.0 - 1e-6))) return 108;
        }
    }
    for (uint32_t i = 0; i < G.n_sph; ++i)
        if (seen[i] != 1) return 109;
    if (n_always_out) *n_always_out = always;
    return 0;
}

// rt_types.h rt_divisor / rt_div against the plain `/`: every divisor of `dlist` (plus the powers of two and their
// neighbours) with numerators around every multiple boundary it can reach, the extremes, and `n_random` random
// numerators each.  Returns the number of mismatches.
extern "C" uint64_t hostsim_divisor_mismatches(const uint32_t* dlist, uint32_t n_d, uint32_t n_random, uint32_t seed)
{
    std::vector<uint32_t> ds(dlist, dlist + n_d);
    for (uint32_t l = 0; l <= 31u; ++l)
        for (int64_t delta = -1; delta <= 1; ++delta) {
            const int64_t d = ((int64_t)1 << l) + delta;
            if (d >= 1 && d <= ((int64_t)1 << 31)) ds.push_back((uint32_t)d);
        }
    uint64_t bad = 0;
    uint32_t x = seed ? seed : 1u;
    auto next = [&x] { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; };
    for (uint32_t d : ds) {
        const RtDivisor k = rt_divisor(d);
        auto check = [&](uint32_t n) { bad += rt_div(n, k) != n / d; };
        for (uint32_t n : {0u, 1u, d - 1u, d, d + 1u, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu}) check(n);
        for (uint32_t i = 0; i < n_random; ++i) {
            const uint32_t n = next();
            check(n);
            const uint32_t m = n / d * d;                  // a multiple of d and its neighbours
            check(m); check(m - 1u); check(m + (d - 1u));
        }
    }
    return bad;
}

// The render kernel's work space (rt_trace.cuh decode_slot): for one launch geometry, decode EVERY slot of every shard
// exactly as the kernel does and check that the valid slots are a bijection onto (pixel, pass[, sample]) of the frame,
// that every shard only touches its own tiles (rt_shard_tile), and that out_index addresses the frame (or the shard's
// compact buffer).  Returns the number of violations.  items != 0: RT_FLAG_SAMPLE_ITEMS with spp samples.
extern "C" uint64_t hostsim_decode_violations(uint32_t W, uint32_t H, uint32_t tile_rows, uint32_t shard_count,
                                              uint32_t passes, int32_t spp, int items, int compact)
{
    using namespace rt;
    const uint32_t subtiles_x = (W + 7u) >> 3, chunks_per_strip = subtiles_x * (tile_rows >> 2);
    const uint32_t per_slot   = items ? (uint32_t)spp : 1u;
    const uint32_t slots_per_tile = chunks_per_strip * 32u * per_slot;
    std::vector<uint32_t> seen((size_t)W * H * passes * per_slot, 0u);
    uint64_t bad = 0;
    for (uint32_t shard = 0; shard < shard_count; ++shard) {
        RtFrameParams P{};
        P.width = W; P.height = H; P.spp = spp; P.passes = passes; P.tile_rows = tile_rows;
        P.tile_first = shard; P.tile_stride = shard_count;
        P.flags = (items ? RT_FLAG_SAMPLE_ITEMS : 0u) | (compact ? RT_FLAG_COMPACT_OUT : 0u);
        P.div_subtiles_x       = rt_divisor(subtiles_x);
        P.div_chunks_per_strip = rt_divisor(chunks_per_strip);
        const uint32_t q_tiles = shard_tile_count(H, tile_rows, shard, shard_count);
        const uint32_t slots_per_pass = q_tiles * slots_per_tile;
        std::vector<uint32_t> compact_seen(compact ? (size_t)q_tiles * tile_rows * W : 0, 0u);
        for (uint32_t slot = 0; slot < slots_per_pass * passes; ++slot) {
            uint32_t sample = 0, pass = 0;
            const PixelSlot s = decode_slot(P, slot, subtiles_x, chunks_per_strip, shard, q_tiles, slots_per_pass, sample, pass);
            if (!s.valid) continue;
            if (s.column >= W || s.image_row >= H || pass >= passes || sample >= per_slot) { ++bad; continue; }
            const uint32_t tile = s.image_row / tile_rows;
            bool mine = false;
            for (uint32_t j = 0; j < q_tiles; ++j) mine = mine || rt_shard_tile(shard, shard_count, j) == tile;
            bad += !mine;
            if (compact) {
                if (s.out_index >= compact_seen.size()) { ++bad; continue; }
                if (pass == 0 && sample == 0) ++compact_seen[s.out_index];
            } else {
                bad += s.out_index != s.image_row * W + s.column;
            }
            ++seen[(((size_t)pass * per_slot + sample) * H + s.image_row) * W + s.column];
        }
        uint64_t used = 0;
        for (uint32_t c : compact_seen) { bad += c > 1u; used += c; }
        if (compact) {                                           // every pixel of the shard's tiles, once
            uint64_t rows = 0;
            for (uint32_t j = 0; j < q_tiles; ++j) {
                const uint32_t first = rt_shard_tile(shard, shard_count, j) * tile_rows;
                rows += first < H ? (H - first < tile_rows ? H - first : tile_rows) : 0u;
            }
            bad += used != rows * W;
        }
    }
    for (uint32_t c : seen) bad += c != 1u;                      // every (pixel, pass, sample) exactly once over all shards
    return bad;
}
