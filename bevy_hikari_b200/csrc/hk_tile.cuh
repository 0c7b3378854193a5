// hk_tile.cuh — 2-D tiles of per-pixel planes staged into shared memory by the TMA unit (cp.async.bulk.tensor.2d, SASS UTMALDG), for
// the kernels whose threads gather from a bounded neighbourhood of their pixel: spatial reuse (neighbour rejection tests + the
// screen-space depth march) and the a-trous levels.  One thread issues one instruction per plane; the copy engine computes the
// addresses, zero-fills what lies outside the plane (= what an out-of-bounds textureLoad returns) and signals an mbarrier.
//
// The first element of a box must sit on a 16-byte boundary of its plane (measured on B200: a 4-byte plane read at an x that is not a
// multiple of 4 raises "illegal instruction"), so tiles start at tile_origin_x() — the wanted first column rounded DOWN to a multiple
// of 4 pixels of the plane — and are TILE_SLACK columns wider than the neighbourhood they must cover.
//
// A TileMap is the 128-byte CUtensorMap the driver encodes on the host (context.cu make_tile_map) for one plane and one box size;
// the kernel-logic emulation (tests/emu) stores a plain description in the same bytes and copies with memcpy.
#pragma once
#include "hk_pool.cuh"      // mbarrier helpers

namespace hkd {

struct alignas(64) TileMap { unsigned char bytes[128]; };

// the CTA's dynamic shared memory (TMA destinations must be 128-byte aligned)
#ifdef HK_EMU
#define HK_DYNAMIC_SMEM(name) alignas(128) static thread_local unsigned char name[128 * 1024]
#else
#define HK_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

constexpr int TILE_SLACK = 3;
__host__ __device__ constexpr int tile_box_width(int columns) { return (columns + TILE_SLACK + 3) & ~3; }   // columns needed -> box width
// plane column (allocation coordinates, may be negative) a tile that must start at or before `wanted` starts at
__device__ __forceinline__ int tile_origin_x(int wanted) { return (wanted >> 2) << 2; }

struct TileMapEmu {          // what the emulated build keeps in TileMap::bytes
    const unsigned char* base;
    uint32_t elem_bytes, width, height, pitch_bytes, box_w, box_h;   // in elements of elem_bytes
};

// Box (box_w x box_h elements, fixed when the map was encoded) whose first element is (x, y) of the plane -> smem_dst (128-byte
// aligned, row-major box); completion adds the box's byte count to `bar`.
__device__ __forceinline__ void tile_load_2d(void* smem_dst, const TileMap* map, int x, int y, uint64_t* bar) {
#ifdef HK_EMU
    const TileMapEmu& m = *reinterpret_cast<const TileMapEmu*>(map->bytes);
    unsigned char* dst = static_cast<unsigned char*>(smem_dst);
    for (uint32_t r = 0; r < m.box_h; ++r)
        for (uint32_t c = 0; c < m.box_w; ++c) {
            const long long sx = (long long)x + c, sy = (long long)y + r;
            unsigned char* d = dst + ((size_t)r * m.box_w + c) * m.elem_bytes;
            if (sx < 0 || sy < 0 || sx >= (long long)m.width || sy >= (long long)m.height) memset(d, 0, m.elem_bytes);
            else memcpy(d, m.base + (size_t)sy * m.pitch_bytes + (size_t)sx * m.elem_bytes, m.elem_bytes);
        }
    (void)bar;
#else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
#endif
}

}  // namespace hkd
