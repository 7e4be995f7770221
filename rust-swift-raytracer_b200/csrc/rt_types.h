// rt_types.h — plain data shared by the host layer and the CUDA kernels.
//
// Layout in HBM (one contiguous, 16-byte aligned "scene blob" per device, built once per
// load_world / World::new and never touched by the render loop again):
//
//   [ sph       : float4 x S ]  {cx, cy, cz, r*r}      hot  — staged into shared memory
//   [ tri_plane : float4 x T ]  {n.x, n.y, n.z, n.v0}  hot  — staged into shared memory
//   [ tri_v     : float4 x 3T]  {v_k.xyz, stored_normal[k]}   warm (plane-stage survivors)
//   [ mat       : float4 x P ]  {r, g, b, fuzz|ir}     cold (one gather per hit), P = S+T
//   [ sph_r     : float  x S ]  radius                 cold
//   [ mat_type  : u32    x P ]  RT_MAT_*               cold
//
// This is the SoA split the reference author sketches in raytracer/TODO.txt:27-39: the
// closest-hit loop reads 16 B per primitive and nothing else.
#pragma once
#include <stddef.h>
#include <stdint.h>

// materials.rs:7-12
enum RtMaterialType : uint32_t {
    RT_MAT_DIFFUSE    = 0,
    RT_MAT_METAL      = 1,
    RT_MAT_DIELECTRIC = 2,
    RT_MAT_EMISSION   = 3,
};

struct RtVec3 { float x, y, z; };

// camera.rs:8-15
struct RtCameraData {
    RtVec3 origin;
    RtVec3 lower_left_corner;
    RtVec3 horizontal;
    RtVec3 vertical;
};

struct RtFloat4 { float x, y, z, w; };

// Device- or host-resident view of a packed scene (pointers into the blob).
struct RtSceneView {
    const RtFloat4* sph;        // [S]
    const RtFloat4* tri_plane;  // [T]
    const RtFloat4* tri_v;      // [3T]
    const RtFloat4* mat;        // [S+T]
    const float*    sph_r;      // [S]
    const uint32_t* mat_type;   // [S+T]
    uint32_t        n_sph;
    uint32_t        n_tri;
};

// Flags of RtFrameParams::flags
enum : uint32_t {
    RT_FLAG_FIXED_JITTER = 1u << 0,   // sub-pixel offset (0.5, 0.5); no jitter draws
    RT_FLAG_ACCUM_IN     = 1u << 1,   // continue from the float4 accumulator (progressive pass)
    RT_FLAG_ACCUM_OUT    = 1u << 2,   // write the float4 accumulator back
    RT_FLAG_NO_RESOLVE   = 1u << 3,   // skip the RGBA8 pack (intermediate progressive pass)
    RT_FLAG_COMPACT_OUT  = 1u << 4,   // out/accum hold only this shard's tiles, packed
};

// Everything one render launch needs (passed by value as a __grid_constant__).
struct RtFrameParams {
    RtCameraData camera;
    uint32_t width, height;
    int32_t  spp;            // samples traced by this launch (common.rs:334)
    int32_t  depth;          // max_ray_bounces (common.rs:267)
    int32_t  sample_begin;   // index of this launch's first sample (progressive passes)
    int32_t  resolve_spp;    // divisor of the resolve (common.rs:345-348)
    uint32_t seed;
    uint32_t flags;
    // Row-tile sharding: this launch renders image-row tiles
    //   tile_first, tile_first + tile_stride, ...   (n_tiles of them, tile_rows rows each).
    uint32_t tile_rows, tile_first, tile_stride, n_tiles;
    uint32_t* out;           // RGBA8 as u32, full frame or compact (RT_FLAG_COMPACT_OUT)
    RtFloat4* accum;         // optional float4 sums, same indexing as out
    unsigned long long* ray_counter;   // += number of World::hit calls
    unsigned int* work_counter;        // zeroed before launch
};
