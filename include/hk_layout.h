// hk_layout.h — byte layouts of every record that crosses the drop-in boundary.
//
// These are the std430 storage-buffer / uniform layouts of the reference
// (src/shaders/mesh_material_types.wgsl:3-83, src/shaders/mesh_view_types.wgsl:3-25,
//  src/shaders/light.wgsl:35-43, mirrored host-side by src/mesh_material/mod.rs:60-299 and src/view.rs:105-123).
// A Rust host that already fills wgpu storage buffers can hand the same bytes to hk_scene_upload().
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hk_node {         /* mesh_material_types.wgsl:35-40 ; mod.rs:177-201 */
    float min[3];
    uint32_t entry_index;        /* child index, or 0x80000000|shape for a leaf record */
    float max[3];
    uint32_t exit_index;         /* next record on miss / after a leaf */
} hk_node;

typedef struct hk_primitive_vertex { /* mesh_material_types.wgsl:10-13 */
    float position[3];
    uint32_t index;
} hk_primitive_vertex;

typedef struct hk_primitive {    /* mesh_material_types.wgsl:15-17 */
    hk_primitive_vertex vertices[3];
} hk_primitive;

typedef struct hk_vertex {       /* mesh_material_types.wgsl:3-8 */
    float position[3];
    float u;
    float normal[3];
    float v;
} hk_vertex;

typedef struct hk_mesh_index {   /* mesh_material_types.wgsl:19-23 */
    uint32_t vertex;
    uint32_t primitive;
    uint32_t node_offset;
    uint32_t node_count;
} hk_mesh_index;

typedef struct hk_instance {     /* mesh_material_types.wgsl:25-33 ; mod.rs:147-156 */
    float min[3];
    uint32_t material;
    float max[3];
    uint32_t node_index;
    float model[16];                    /* column-major */
    float inverse_transpose_model[16];  /* column-major */
    hk_mesh_index mesh;
} hk_instance;

typedef struct hk_material {     /* mesh_material_types.wgsl:42-56 ; mod.rs:203-218 */
    float base_color[4];
    uint32_t base_color_texture;
    uint32_t _pad0[3];
    float emissive[4];
    uint32_t emissive_texture;
    float perceptual_roughness;
    float metallic;
    uint32_t metallic_roughness_texture;
    float reflectance;
    uint32_t normal_map_texture;
    uint32_t occlusion_texture;
    uint32_t _pad1;
} hk_material;

typedef struct hk_alias_entry {  /* mesh_material_types.wgsl:58-61 */
    float prob;
    uint32_t index;
} hk_alias_entry;

typedef struct hk_emissive {     /* mesh_material_types.wgsl:63-71 ; mod.rs:228-237 */
    float emissive[4];
    float position[3];
    float radius;
    uint32_t instance;
    uint32_t _pad0;
    uint32_t alias_table_offset;
    uint32_t alias_table_count;
    float surface_area;
    uint32_t node_index;
    uint32_t _pad1[2];
} hk_emissive;

typedef struct hk_packed_reservoir { /* light.wgsl:35-43 */
    uint32_t radiance[2];            /* 4 x f16 */
    uint32_t random[2];              /* 4 x unorm16 */
    float visible_position[4];       /* w = depth */
    float sample_position[4];        /* w = f32(visible_instance) */
    uint32_t visible_normal;         /* snorm8 xyz, w = lifetime/127 - 1 */
    uint32_t sample_normal;          /* snorm8 xyz, w = sample_position.w */
    uint32_t reservoir[2];           /* f16: count, w | w_sum, w2_sum */
} hk_packed_reservoir;

typedef struct hk_frame_uniform {  /* mesh_view_types.wgsl:3-20 ; view.rs:105-123 */
    float kernel[3][4];            /* mat3x3: 3 columns padded to 16 B */
    float halton[8][4];
    float clear_color[4];
    uint32_t number;
    uint32_t direct_validate_interval;
    uint32_t emissive_validate_interval;
    uint32_t indirect_bounces;
    uint32_t temporal_reuse;
    uint32_t emissive_spatial_reuse;
    uint32_t indirect_spatial_reuse;
    uint32_t max_temporal_reuse_count;
    uint32_t max_spatial_reuse_count;
    float max_reservoir_lifetime;
    float solar_angle;
    float max_indirect_luminance;
    float upscale_ratio;
    uint32_t _pad[3];
} hk_frame_uniform;

typedef struct hk_previous_view { /* mesh_view_types.wgsl:22-25 ; view.rs:31-35 */
    float view_proj[16];
    float inverse_view_proj[16];
} hk_previous_view;

/* The slice of bevy_pbr 0.9 `View` that the path reads (light.wgsl:714-727,1040; prepass.wgsl:45-71). */
typedef struct hk_view {
    float view_proj[16];
    float inverse_view_proj[16];
    float view[16];
    float inverse_view[16];
    float projection[16];
    float inverse_projection[16];
    float world_position[3];
    float _pad0;
    float viewport[4];
} hk_view;

/* The slice of bevy_pbr 0.9 `Lights` that the path reads (light.wgsl:611-613,832,852-855). */
typedef struct hk_lights {
    float directional_color[4];       /* lights.directional_lights[0].color */
    float direction_to_light[3];      /* lights.directional_lights[0].direction_to_light */
    float _pad0;
    float ambient_color[4];           /* lights.ambient_color */
} hk_lights;

#ifdef __cplusplus
}
static_assert(sizeof(hk_node) == 32 && offsetof(hk_node, entry_index) == 12 && offsetof(hk_node, exit_index) == 28, "Node");
static_assert(sizeof(hk_primitive) == 48, "Primitive");
static_assert(sizeof(hk_vertex) == 32 && offsetof(hk_vertex, normal) == 16, "Vertex");
static_assert(sizeof(hk_instance) == 176 && offsetof(hk_instance, model) == 32 &&
              offsetof(hk_instance, inverse_transpose_model) == 96 && offsetof(hk_instance, mesh) == 160, "Instance");
static_assert(sizeof(hk_material) == 80 && offsetof(hk_material, base_color_texture) == 16 &&
              offsetof(hk_material, emissive) == 32 && offsetof(hk_material, emissive_texture) == 48 &&
              offsetof(hk_material, reflectance) == 64 && offsetof(hk_material, occlusion_texture) == 72, "Material");
static_assert(sizeof(hk_alias_entry) == 8, "AliasEntry");
static_assert(sizeof(hk_emissive) == 64 && offsetof(hk_emissive, position) == 16 && offsetof(hk_emissive, instance) == 32 &&
              offsetof(hk_emissive, alias_table_offset) == 40 && offsetof(hk_emissive, surface_area) == 48, "Emissive");
static_assert(sizeof(hk_packed_reservoir) == 64 && offsetof(hk_packed_reservoir, visible_position) == 16 &&
              offsetof(hk_packed_reservoir, visible_normal) == 48 && offsetof(hk_packed_reservoir, reservoir) == 56, "PackedReservoir");
static_assert(sizeof(hk_frame_uniform) == 256 && offsetof(hk_frame_uniform, halton) == 48 &&
              offsetof(hk_frame_uniform, clear_color) == 176 && offsetof(hk_frame_uniform, number) == 192 &&
              offsetof(hk_frame_uniform, upscale_ratio) == 240, "Frame");
static_assert(sizeof(hk_previous_view) == 128, "PreviousView");
#endif
