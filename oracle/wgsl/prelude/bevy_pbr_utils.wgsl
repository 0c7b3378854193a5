#define_import_path bevy_pbr::utils
// PRELUDE — not the reference's text: bevy_pbr 0.9.1 `utils.wgsl`, the part the hikari shaders use, restated (SURVEY App. D).

let PI: f32 = 3.141592653589793;

fn saturate(value: f32) -> f32 {
    return clamp(value, 0.0, 1.0);
}
