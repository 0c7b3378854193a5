// hikari_plugin.cpp — the part of the host mirror that CALLS the C ABI: the three render-graph nodes and HikariPlugin.
// Linked into libhikari_b200.so; everything else of the mirror (scene preparation, settings, frame uniforms) is in
// libhikari_host.so, which has no CUDA dependency at all (the CPU reference arm of bench.py loads only that one).
#include "hikari.hpp"

namespace hikari {

int PrepassNode::run(hk_context* ctx, const hk_frame_inputs& in) { return hk_prepass_run(ctx, &in); }
int LightNode::run(hk_context* ctx, const hk_frame_inputs& in) { return hk_light_run(ctx, &in); }
int PostProcessNode::run(hk_context* ctx, const hk_frame_inputs& in) { return hk_post_process_run(ctx, &in); }

HikariPlugin::~HikariPlugin() { if (ctx_) hk_context_destroy(ctx_); }
int HikariPlugin::build(int cuda_device, uint32_t width, uint32_t height, uint32_t row_begin, uint32_t row_end,
                        const uint8_t* noise, void* cuda_stream) {
    return build_tile(cuda_device, width, height, 0, width, row_begin, row_end, noise, cuda_stream);
}
int HikariPlugin::build_tile(int cuda_device, uint32_t width, uint32_t height, uint32_t col_begin, uint32_t col_end,
                             uint32_t row_begin, uint32_t row_end, const uint8_t* noise, void* cuda_stream) {
    if (ctx_) { hk_context_destroy(ctx_); ctx_ = nullptr; }
    int e = hk_context_create_tile(&ctx_, cuda_device, width, height, col_begin, col_end, row_begin, row_end, cuda_stream);
    if (e != HK_OK) return e;
    counter.value = 0;
    return hk_set_noise(ctx_, noise);
}
int HikariPlugin::upload_scene(const MeshMaterialWorld& world) {
    if (!ctx_) return HK_ERR_NOT_READY;
    hk_scene_desc d = world.scene_desc();
    return hk_scene_upload(ctx_, &d);
}
int HikariPlugin::update_instances(const MeshMaterialWorld& world) {
    if (!ctx_) return HK_ERR_NOT_READY;
    hk_scene_desc d = world.scene_desc();
    return hk_scene_update_instances(ctx_, &d);
}
int HikariPlugin::update_transforms(MeshMaterialWorld& world, bool* used_device_path) {
    if (!ctx_) return HK_ERR_NOT_READY;
    if (world.prepare_instance_transforms()) {
        int rc = hk_scene_update_transforms(ctx_, world.transform_models.data(), world.transform_previous.data(), world.transform_aabbs.data(),
                                            (uint32_t)(world.transform_models.size() / 16));
        if (rc != HK_ERR_UNSUPPORTED) {      // (a TLAS that is not in bvh 0.7.1's layout keeps the host path)
            if (used_device_path) *used_device_path = true;
            return rc;
        }
    }
    if (used_device_path) *used_device_path = false;
    world.prepare_instances();
    return update_instances(world);
}
int HikariPlugin::run_frame(const HikariSettings& settings, const ViewInputs& view) {
    if (!ctx_) return HK_ERR_NOT_READY;
    counter.value += 1;
    hk_frame_inputs in = make_frame_inputs(settings, counter, view);
    // The reference runs smaa_tu4x / taa_jasmine / FSR1 whenever the settings select them (post_process.rs:1236-1308), so that
    // is the default here too.  A caller that ends the path at the tone-mapped image (the sharded benchmark) opts out; the
    // prepass jitter only makes sense with the passes that resolve it, so it is switched off with them.
    in.temporal_upscalers = temporal_upscalers ? 1u : 0u;
    if (!temporal_upscalers) in.taa_jitter = 0u;
    return hk_render_frame(ctx_, &in);
}
std::string HikariPlugin::last_error() const { return hk_last_error(ctx_); }

}  // namespace hikari
