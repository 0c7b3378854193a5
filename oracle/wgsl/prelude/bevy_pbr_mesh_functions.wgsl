#define_import_path bevy_pbr::mesh_functions
// PRELUDE — not the reference's text: bevy_pbr 0.9.1 `mesh_functions.wgsl`, the two functions prepass.wgsl calls, restated (SURVEY App. D).
// `mesh` and `view` are the importing module's bindings, as in bevy.

fn mesh_position_local_to_world(model: mat4x4<f32>, vertex_position: vec4<f32>) -> vec4<f32> {
    return model * vertex_position;
}

fn mesh_normal_local_to_world(vertex_normal: vec3<f32>) -> vec3<f32> {
    // NOTE: The mikktspace method of normal mapping requires that the world normal is
    // re-normalized in the vertex shader to match the way mikktspace bakes vertex tangents
    // and normal maps so that the exported normal map can be used as-is.
    return normalize(
        mat3x3<f32>(
            mesh.inverse_transpose_model[0].xyz,
            mesh.inverse_transpose_model[1].xyz,
            mesh.inverse_transpose_model[2].xyz
        ) * vertex_normal
    );
}
