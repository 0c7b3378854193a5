/* box.c — the drop-in boundary used from plain C: no Python, no torch, only include/hikari_host.h + include/hikari_b200.h.
 *
 * What a Bevy app does around bevy-hikari's render-graph nodes, restated in ~150 lines of C: spawn meshes / materials /
 * instances (MeshMaterialPlugin's world, src/mesh_material), add a camera (perspective, infinite reverse-Z like bevy's
 * default), run HikariPlugin for a few frames with HikariSettings::default() overrides, read the tone-mapped image back.
 * Scene: a floor, a box and a small emissive quad above them (a poor man's cornell).
 *
 *   gcc -O2 -Iinclude examples/c/box.c -Lbevy_hikari_b200 -lhikari_b200 -lhikari_host -lm -Wl,-rpath,$PWD/bevy_hikari_b200 -o box
 *   ./box data/noise_rgba8_64x64x16.bin          (needs a CUDA device: there is no CPU fallback)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hikari_host.h"

typedef struct { float m[16]; } mat4;   /* column-major, like glam / the headers */

static mat4 mul(mat4 a, mat4 b) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int w = 0; w < 4; ++w) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += a.m[k * 4 + w] * b.m[c * 4 + k];
            r.m[c * 4 + w] = s;
        }
    return r;
}
static mat4 translation_scale(float x, float y, float z, float sx, float sy, float sz) {
    mat4 r; memset(&r, 0, sizeof r);
    r.m[0] = sx; r.m[5] = sy; r.m[10] = sz; r.m[15] = 1.0f; r.m[12] = x; r.m[13] = y; r.m[14] = z;
    return r;
}
/* camera world transform looking from `eye` to `target` (right-handed, -Z forward, +Y up) and its inverse (rigid) */
static void look_at(const float eye[3], const float target[3], mat4* world, mat4* view) {
    float f[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
    float fl = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int i = 0; i < 3; ++i) f[i] /= fl;
    float up[3] = {0, 1, 0};
    float r[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    float rl = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    for (int i = 0; i < 3; ++i) r[i] /= rl;
    float u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
    memset(world, 0, sizeof *world); memset(view, 0, sizeof *view);
    for (int i = 0; i < 3; ++i) {
        world->m[0 + i] = r[i]; world->m[4 + i] = u[i]; world->m[8 + i] = -f[i]; world->m[12 + i] = eye[i];
        view->m[i * 4 + 0] = r[i]; view->m[i * 4 + 1] = u[i]; view->m[i * 4 + 2] = -f[i];
    }
    world->m[15] = view->m[15] = 1.0f;
    view->m[12] = -(r[0] * eye[0] + r[1] * eye[1] + r[2] * eye[2]);
    view->m[13] = -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]);
    view->m[14] = (f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2]);
}
/* Mat4::perspective_infinite_reverse_rh(fov_y, aspect, near) and its inverse */
static void perspective(float fov, float aspect, float z_near, mat4* p, mat4* inv) {
    float f = 1.0f / tanf(0.5f * fov);
    memset(p, 0, sizeof *p); memset(inv, 0, sizeof *inv);
    p->m[0] = f / aspect; p->m[5] = f; p->m[11] = -1.0f; p->m[14] = z_near;
    inv->m[0] = aspect / f; inv->m[5] = 1.0f / f; inv->m[11] = 1.0f / z_near; inv->m[14] = -1.0f;
}

static uint32_t add_quad(hikari_world* w, float half) {   /* a quad in the xz plane, normal +y (bevy shape::Plane) */
    const float pos[12] = {half, 0, -half, half, 0, half, -half, 0, half, -half, 0, -half};
    const float nrm[12] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
    const float uv[8] = {1, 0, 1, 1, 0, 1, 0, 0};
    const uint32_t idx[6] = {0, 2, 1, 0, 3, 2};
    return hikari_world_add_mesh(w, pos, nrm, uv, 4, idx, 6, 0);
}
static uint32_t add_box(hikari_world* w, float h) {       /* 6 faces x 4 vertices */
    static const float n[6][3] = {{0, 0, 1}, {0, 0, -1}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    float pos[72], nrm[72], uv[48]; uint32_t idx[36];
    for (int f = 0; f < 6; ++f) {
        float a[3] = {n[f][1], n[f][2], n[f][0]}, b[3] = {n[f][1] * a[2] - n[f][2] * a[1], n[f][2] * a[0] - n[f][0] * a[2], n[f][0] * a[1] - n[f][1] * a[0]};
        const float s[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int v = 0; v < 4; ++v)
            for (int k = 0; k < 3; ++k) {
                pos[(f * 4 + v) * 3 + k] = h * (n[f][k] + s[v][0] * a[k] + s[v][1] * b[k]);
                nrm[(f * 4 + v) * 3 + k] = n[f][k];
            }
        for (int v = 0; v < 4; ++v) { uv[(f * 4 + v) * 2] = 0.5f * (s[v][0] + 1); uv[(f * 4 + v) * 2 + 1] = 0.5f * (s[v][1] + 1); }
        const uint32_t q[6] = {0, 1, 2, 0, 2, 3};
        for (int k = 0; k < 6; ++k) idx[f * 6 + k] = (uint32_t)(f * 4) + q[k];
    }
    return hikari_world_add_mesh(w, pos, nrm, uv, 24, idx, 36, 0);
}
static hk_material material(float r, float g, float b, float er, float eg, float eb, float ea) {
    hk_material m; memset(&m, 0, sizeof m);
    m.base_color[0] = r; m.base_color[1] = g; m.base_color[2] = b; m.base_color[3] = 1.0f;
    m.emissive[0] = er; m.emissive[1] = eg; m.emissive[2] = eb; m.emissive[3] = ea;
    m.perceptual_roughness = 0.6f; m.metallic = 0.01f; m.reflectance = 0.5f;
    m.base_color_texture = m.emissive_texture = m.metallic_roughness_texture = m.normal_map_texture = m.occlusion_texture = 0xFFFFFFFFu;
    return m;
}

int main(int argc, char** argv) {
    enum { W = 160, H = 120, FRAMES = 8 };
    static uint8_t noise[16 * 64 * 64 * 4];
    FILE* fp = fopen(argc > 1 ? argv[1] : "data/noise_rgba8_64x64x16.bin", "rb");
    if (!fp || fread(noise, 1, sizeof noise, fp) != sizeof noise) { fprintf(stderr, "cannot read the blue-noise textures\n"); return 2; }
    fclose(fp);

    hikari_world* world = hikari_world_create();
    uint32_t quad = add_quad(world, 0.5f), box = add_box(world, 0.5f);
    hk_material grey = material(0.8f, 0.8f, 0.8f, 0, 0, 0, 1), red = material(0.8f, 0.2f, 0.15f, 0, 0, 0, 1), lamp = material(1, 1, 1, 1.0f, 0.9f, 0.7f, 0.2f);
    uint32_t m_grey = hikari_world_add_material(world, &grey), m_red = hikari_world_add_material(world, &red), m_lamp = hikari_world_add_material(world, &lamp);
    mat4 floor_xf = translation_scale(0, 0, 0, 6, 1, 6), box_xf = translation_scale(0.3f, 0.5f, -0.2f, 1, 1, 1);
    mat4 lamp_xf = translation_scale(0, 2.2f, 0, 1, -1, 1);      /* mirrored: faces down */
    hikari_world_add_instance(world, quad, m_grey, floor_xf.m, 1);
    hikari_world_add_instance(world, box, m_red, box_xf.m, 1);
    hikari_world_add_instance(world, quad, m_lamp, lamp_xf.m, 1);
    hikari_world_prepare(world);

    hikari_plugin* plugin = hikari_plugin_create();
    int rc = hikari_plugin_build(plugin, 0, W, H, 0, H, noise, NULL);
    if (rc != HK_OK) { fprintf(stderr, "hikari_plugin_build: %d (%s)\n", rc, hk_last_error(NULL)); return 3; }   /* e.g. no CUDA device */
    if ((rc = hikari_plugin_upload_scene(plugin, world)) != HK_OK) { fprintf(stderr, "upload: %s\n", hk_last_error(hikari_plugin_context(plugin))); return 4; }

    hikari_settings settings; hikari_settings_default(&settings);
    settings.taa = HIKARI_TAA_NONE; settings.upscale_kind = HIKARI_UPSCALE_SMAA_TU4X; settings.upscale_ratio = 1.0f;
    settings.indirect_bounces = 2; settings.emissive_spatial_reuse = 1;

    hk_view view; hk_previous_view previous; hk_lights lights;
    memset(&view, 0, sizeof view); memset(&lights, 0, sizeof lights);
    const float eye[3] = {2.5f, 1.8f, 3.5f}, target[3] = {0, 0.6f, 0};
    mat4 cam, v, p, ip; look_at(eye, target, &cam, &v); perspective(0.785398163f, (float)W / H, 0.1f, &p, &ip);
    mat4 vp = mul(p, v), ivp = mul(cam, ip);
    memcpy(view.view_proj, vp.m, 64); memcpy(view.inverse_view_proj, ivp.m, 64); memcpy(view.view, cam.m, 64); memcpy(view.inverse_view, v.m, 64);
    memcpy(view.projection, p.m, 64); memcpy(view.inverse_projection, ip.m, 64); memcpy(view.world_position, eye, 12);
    view.viewport[2] = W; view.viewport[3] = H;
    memcpy(previous.view_proj, vp.m, 64); memcpy(previous.inverse_view_proj, ivp.m, 64);          /* static camera */
    lights.ambient_color[0] = lights.ambient_color[1] = lights.ambient_color[2] = 0.05f; lights.ambient_color[3] = 1.0f;

    for (int f = 0; f < FRAMES; ++f)
        if ((rc = hikari_plugin_run_frame(plugin, &settings, &view, &previous, &lights)) != HK_OK) {
            fprintf(stderr, "frame %d: %s\n", f, hk_last_error(hikari_plugin_context(plugin))); return 5;
        }
    hk_context* ctx = hikari_plugin_context(plugin);
    uint32_t ow = 0, oh = 0; hk_output_extent(ctx, HK_OUT_GBUFFER_POSITION, &ow, &oh);
    float* position = malloc((size_t)ow * oh * 16);
    uint16_t* image = malloc((size_t)W * H * 8);
    if (hk_readback(ctx, HK_OUT_GBUFFER_POSITION, position, (size_t)ow * oh * 16) != HK_OK || hk_readback(ctx, HK_OUT_TONE_MAPPED, image, (size_t)W * H * 8) != HK_OK) {
        fprintf(stderr, "readback: %s\n", hk_last_error(ctx)); return 6;
    }
    size_t covered = 0, lit = 0;
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const int hit = position[4 * i + 3] > 0.0f;
        covered += hit;
        lit += hit && (image[4 * i] & 0x7FFF) > 0x2000;   /* f16 red channel above ~0.008 on a surface */
    }
    hk_frame_stats st; hk_get_stats(ctx, &st);
    printf("%s  %ux%u  frames=%d  covered=%.3f  lit=%.3f  kernel_launches/frame=%u\n", hk_version(), W, H, FRAMES, (double)covered / (W * H),
           (double)lit / (double)(covered ? covered : 1), st.kernel_launches);
    int ok = covered > (size_t)W * H / 4 && lit > covered / 2 && st.kernel_launches >= 10;
    free(position); free(image);
    hikari_plugin_destroy(plugin); hikari_world_destroy(world);
    return ok ? 0 : 1;
}
