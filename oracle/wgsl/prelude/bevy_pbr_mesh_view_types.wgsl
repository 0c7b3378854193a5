#define_import_path bevy_pbr::mesh_view_types
// PRELUDE — not the reference's text: bevy_pbr 0.9.1 `mesh_view_types.wgsl`, the two uniform structs the hikari shaders bind
// (mesh_view_bindings.wgsl:9,13), restated (SURVEY App. D).  The path reads view.{view_proj, inverse_view_proj, projection,
// world_position} and lights.{directional_lights[0].color, .direction_to_light, ambient_color} (light.wgsl:611-613,714-727,832).

struct View {
    view_proj: mat4x4<f32>,
    inverse_view_proj: mat4x4<f32>,
    view: mat4x4<f32>,
    inverse_view: mat4x4<f32>,
    projection: mat4x4<f32>,
    inverse_projection: mat4x4<f32>,
    world_position: vec3<f32>,
    viewport: vec4<f32>,
};

struct DirectionalLight {
    view_projection: mat4x4<f32>,
    color: vec4<f32>,
    direction_to_light: vec3<f32>,
    flags: u32,
    shadow_depth_bias: f32,
    shadow_normal_bias: f32,
};

struct Lights {
    directional_lights: array<DirectionalLight, 10u>,
    ambient_color: vec4<f32>,
    cluster_dimensions: vec4<u32>,
    cluster_factors: vec4<f32>,
    n_directional_lights: u32,
    spot_light_shadowmap_offset: i32,
};
