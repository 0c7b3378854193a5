"""Python face of the host mirror (bevy_hikari_b200/host/hikari.hpp): the names a bevy-hikari user knows —
HikariPlugin, HikariSettings, Taa, Upscale, graph NAME — driving libhikari_b200.so through ctypes.  No arithmetic of
the path happens here; this file only moves bytes across the C ABI."""
import ctypes as C
import os

import numpy as np

from . import _ffi
from . import layout as L
from ._ffi import Settings, check, host_lib, lib

TAA_JASMINE, TAA_NONE = 0, 1
TUNE_POOLED_INDIRECT, TUNE_TILED_SPATIAL, TUNE_TILED_DENOISE, TUNE_WIDE_TRAVERSAL = 1, 2, 3, 4
UPSCALE_FSR1, UPSCALE_SMAA_TU4X = 0, 1
NOISE_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "noise_rgba8_64x64x16.bin")


def graph_name():
    return host_lib().hikari_graph_name().decode()


def HikariSettings(**overrides):
    """HikariSettings::default() (src/lib.rs:435-455) with keyword overrides, as a ctypes struct."""
    s = Settings()
    host_lib().hikari_settings_default(C.byref(s))
    for k, v in overrides.items():
        if k == "clear_color":
            s.clear_color[:] = list(v)
        elif not hasattr(s, k):
            raise AttributeError(f"HikariSettings has no field {k}")
        else:
            setattr(s, k, v)
    return s


def make_frame_inputs(settings, frame_counter, view, previous_view, lights, temporal_upscalers=False):
    out = L.FrameInputs()
    host_lib().hikari_make_frame_inputs(C.byref(settings), int(frame_counter), C.byref(view), C.byref(previous_view),
                                   C.byref(lights), C.byref(out))
    out.temporal_upscalers = 1 if temporal_upscalers else 0
    return out


def load_noise():
    a = np.fromfile(NOISE_PATH, np.uint8)
    assert a.size == 16 * 64 * 64 * 4, "data/noise_rgba8_64x64x16.bin is damaged"
    return a


class World:
    """MeshMaterialPlugin's render-world state: meshes, materials, textures, instances -> the nine GPU buffers."""

    def __init__(self):
        self._w = host_lib().hikari_world_create()
        self._keep = []

    def __del__(self):
        if getattr(self, "_w", None):
            host_lib().hikari_world_destroy(self._w)
            self._w = None

    def add_mesh(self, positions, normals, uvs, indices=None, topology=0):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float32) for a in (positions, normals, uvs)]
        n = len(arrs[0]) if arrs[0] is not None else 0
        idx = None if indices is None else np.ascontiguousarray(indices, np.uint32).reshape(-1)
        ptr = lambda a: None if a is None else a.ctypes.data
        return host_lib().hikari_world_add_mesh(self._w, ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), n, ptr(idx),
                                           0 if idx is None else idx.size, topology)

    def add_material(self, material_record):
        m = np.zeros(1, L.MATERIAL)
        m[0] = material_record
        return host_lib().hikari_world_add_material(self._w, m.ctypes.data)

    def set_material(self, material_id, material_record):
        m = np.zeros(1, L.MATERIAL)
        m[0] = material_record
        host_lib().hikari_world_set_material(self._w, material_id, m.ctypes.data)

    def prepare_materials(self):
        host_lib().hikari_world_prepare_materials(self._w)

    def add_texture(self, rgba, address_mode_u=0, address_mode_v=0, filter_linear=1, srgb=1):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        t = L.TextureDesc(rgba.ctypes.data, rgba.shape[1], rgba.shape[0], address_mode_u, address_mode_v, filter_linear, srgb)
        return host_lib().hikari_world_add_texture(self._w, C.byref(t))

    def add_instance(self, mesh, material, transform16, visible=True):
        t = np.ascontiguousarray(transform16, np.float32).reshape(16)
        return host_lib().hikari_world_add_instance(self._w, mesh, material, t.ctypes.data, 1 if visible else 0)

    def prepare(self):
        host_lib().hikari_world_prepare(self._w)

    # animated instances (transform.rs:31-44, instance.rs:352-437)
    def set_instance_transform(self, instance, transform16):
        t = np.ascontiguousarray(transform16, np.float32).reshape(16)
        host_lib().hikari_world_set_instance_transform(self._w, instance, t.ctypes.data)

    def set_instance_visible(self, instance, visible):
        host_lib().hikari_world_set_instance_visible(self._w, instance, 1 if visible else 0)

    def previous_transform_system(self):
        host_lib().hikari_world_previous_transform_system(self._w)

    def prepare_instances(self):
        host_lib().hikari_world_prepare_instances(self._w)

    def prepare_instance_transforms(self):
        """The transforms-only form of prepare_instances (hikari_world_prepare_instance_transforms): (models, previous_models,
        mesh_aabbs) as float32 arrays of 16 / 16 / 6 columns, or None when the device path does not apply."""
        m, pm, ab, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        if not host_lib().hikari_world_prepare_instance_transforms(self._w, C.byref(m), C.byref(pm), C.byref(ab), C.byref(n)):
            return None
        k = n.value

        def arr(ptr, cols):
            return np.frombuffer(C.string_at(ptr.value, k * cols * 4), np.float32).reshape(k, cols).copy() if k else np.zeros((0, cols), np.float32)
        return arr(m, 16), arr(pm, 16), arr(ab, 6)

    def previous_models(self):
        d = self.scene_desc()
        if not d.previous_instance_models or d.instance_count == 0:
            return np.zeros((0, 16), np.float32)
        return np.frombuffer(C.string_at(d.previous_instance_models, d.instance_count * 64), np.float32).reshape(-1, 16).copy()

    def mesh_error(self, mesh):
        return host_lib().hikari_world_mesh_error(self._w, mesh)

    def scene_desc(self):
        d = L.SceneDesc()
        host_lib().hikari_world_scene_desc(self._w, C.byref(d))
        return d

    def buffers(self):
        """Copies of the nine buffers as numpy structured arrays (for comparison with the oracle's builder)."""
        d = self.scene_desc()
        out = {}
        counts = {"vertices": d.vertex_count, "primitives": d.primitive_count, "asset_nodes": d.asset_node_count,
                  "alias_table": d.alias_count, "instances": d.instance_count, "instance_nodes": d.instance_node_count,
                  "materials": d.material_count, "emissive_nodes": d.emissive_node_count, "emissives": d.emissive_count}
        for name, dt in L.SCENE_BUFFERS:
            n = counts[name]
            ptr = getattr(d, name)
            if n == 0 or not ptr:
                out[name] = np.zeros(0, dt)
            else:
                out[name] = np.frombuffer(C.string_at(ptr, n * dt.itemsize), dt).copy()
        return out


def scene_desc_from_buffers(buffers, textures=()):
    """hk_scene_desc over numpy arrays (keeps them alive on the returned object)."""
    d = L.SceneDesc()
    keep = []
    names = {"vertices": "vertex_count", "primitives": "primitive_count", "asset_nodes": "asset_node_count",
             "alias_table": "alias_count", "instances": "instance_count", "instance_nodes": "instance_node_count",
             "materials": "material_count", "emissive_nodes": "emissive_node_count", "emissives": "emissive_count"}
    for name, dt in L.SCENE_BUFFERS:
        a = np.ascontiguousarray(buffers[name], dt)
        keep.append(a)
        setattr(d, name, a.ctypes.data if a.size else None)
        setattr(d, names[name], a.size)
    if textures:
        arr = (L.TextureDesc * len(textures))()
        for i, t in enumerate(textures):
            rgba = np.ascontiguousarray(t["rgba"], np.uint8)
            keep.append(rgba)
            arr[i] = L.TextureDesc(rgba.ctypes.data, rgba.shape[1], rgba.shape[0], int(t.get("address_mode_u", 0)),
                                   int(t.get("address_mode_v", 0)), int(t.get("filter_linear", 1)), int(t.get("srgb", 1)))
        keep.append(arr)
        d.textures = C.cast(arr, C.c_void_p)
        d.texture_count = len(textures)
    d._keep = keep
    return d


class HikariPlugin:
    """HikariPlugin + one camera with `CameraRenderGraph::new(graph::NAME)` (src/lib.rs:95-370)."""

    def __init__(self, width, height, cuda_device=0, row_begin=0, row_end=None, cuda_stream=None, col_begin=0, col_end=None, flavor=None):
        self._lib = lib(flavor)        # "product" | "exact" | None = _ffi.DEFAULT_FLAVOR
        self.width, self.height = width, height
        self.row_begin, self.row_end = row_begin, height if row_end is None else row_end
        self.col_begin, self.col_end = col_begin, width if col_end is None else col_end
        self._p = self._lib.hikari_plugin_create()
        noise = load_noise()
        rc = self._lib.hikari_plugin_build_tile(self._p, cuda_device, width, height, self.col_begin, self.col_end, self.row_begin,
                                            self.row_end, noise.ctypes.data, cuda_stream)
        if rc != _ffi.HK_OK:
            msg = self._lib.hk_last_error(None).decode()
            self._lib.hikari_plugin_destroy(self._p)
            self._p = None
            raise _ffi.HikariError(f"HikariPlugin.build failed ({rc}): {msg}")
        self.ctx = self._lib.hikari_plugin_context(self._p)

    def close(self):
        if getattr(self, "_p", None):
            self._lib.hikari_plugin_destroy(self._p)
            self._p = None
            self.ctx = None

    __del__ = close

    @property
    def owned_rows(self):
        return self.row_end - self.row_begin

    @property
    def owned_cols(self):
        return self.col_end - self.col_begin

    def upload_scene(self, world):
        check(self._lib.hikari_plugin_upload_scene(self._p, world._w), self.ctx, self._lib)

    def upload_scene_desc(self, desc):
        check(self._lib.hk_scene_upload(self.ctx, C.byref(desc)), self.ctx, self._lib)

    def update_instances(self, world):
        check(self._lib.hikari_plugin_update_instances(self._p, world._w), self.ctx, self._lib)

    def update_instances_desc(self, desc):
        check(self._lib.hk_scene_update_instances(self.ctx, C.byref(desc)), self.ctx, self._lib)

    def update_transforms(self, world):
        """Animated instances, rebuilt on the device when only transforms changed (hk_scene_update_transforms), else through
        prepare_instances + hk_scene_update_instances.  Returns True when the device path ran."""
        used = C.c_int(0)
        check(self._lib.hikari_plugin_update_transforms(self._p, world._w, C.byref(used)), self.ctx, self._lib)
        return bool(used.value)

    def update_transforms_arrays(self, models, mesh_aabbs, previous_models=None):
        """hk_scene_update_transforms on explicit arrays (n x 16, n x 6, optional n x 16)"""
        models = np.ascontiguousarray(models, np.float32); mesh_aabbs = np.ascontiguousarray(mesh_aabbs, np.float32)
        prev = None if previous_models is None else np.ascontiguousarray(previous_models, np.float32)
        check(self._lib.hk_scene_update_transforms(self.ctx, models.ctypes.data, prev.ctypes.data if prev is not None else None,
                                                   mesh_aabbs.ctypes.data, len(models)), self.ctx, self._lib)

    def scene_readback(self, which):
        """the per-frame scene buffers as they are on the device (hk_scene_readback), as numpy structured arrays"""
        dt = {L.SCENE_INSTANCES: L.INSTANCE, L.SCENE_INSTANCE_NODES: L.NODE, L.SCENE_EMISSIVES: L.EMISSIVE, L.SCENE_EMISSIVE_NODES: L.NODE,
              L.SCENE_PREVIOUS_MODELS: np.dtype((np.float32, 16)), L.SCENE_INSTANCE_MOVED: np.dtype(np.uint32)}[which]
        nbytes = C.c_size_t(0)
        check(self._lib.hk_scene_buffer_bytes(self.ctx, int(which), C.byref(nbytes)), self.ctx, self._lib)
        out = np.zeros(nbytes.value // dt.itemsize, dt)
        check(self._lib.hk_scene_readback(self.ctx, int(which), out.ctypes.data, out.nbytes), self.ctx, self._lib)
        return out

    @property
    def frame_counter(self):
        return self._lib.hikari_plugin_frame_counter(self._p)

    @frame_counter.setter
    def frame_counter(self, v):
        self._lib.hikari_plugin_set_frame_counter(self._p, int(v))

    def set_temporal_upscalers(self, enabled):
        self._lib.hikari_plugin_set_temporal_upscalers(self._p, 1 if enabled else 0)

    def run_frame(self, settings, view, previous_view, lights):
        check(self._lib.hikari_plugin_run_frame(self._p, C.byref(settings), C.byref(view), C.byref(previous_view), C.byref(lights)),
              self.ctx, self._lib)

    def import_gbuffer(self, pointers_and_pitches):
        """hk_import_gbuffer: five (device pointer, row pitch in bytes) pairs in the order position, normal, depth_gradient,
        instance_material, velocity_uv"""
        flat = []
        for ptr, pitch in pointers_and_pitches:
            flat += [ptr, pitch]
        desc = (C.c_size_t * 10)(*flat)
        check(self._lib.hk_import_gbuffer(self.ctx, desc), self.ctx, self._lib)

    # individual nodes / raw C ABI
    def prepass(self, inputs): check(self._lib.hk_prepass_run(self.ctx, C.byref(inputs)), self.ctx, self._lib)
    def light(self, inputs): check(self._lib.hk_light_run(self.ctx, C.byref(inputs)), self.ctx, self._lib)
    def run_pass(self, inputs, which, arg=0): check(self._lib.hk_run_pass(self.ctx, C.byref(inputs), int(which), int(arg)), self.ctx, self._lib)
    def post_process(self, inputs): check(self._lib.hk_post_process_run(self.ctx, C.byref(inputs)), self.ctx, self._lib)
    def render_frame(self, inputs): check(self._lib.hk_render_frame(self.ctx, C.byref(inputs)), self.ctx, self._lib)
    def sync(self): check(self._lib.hk_sync(self.ctx), self.ctx, self._lib)
    def reset_temporal_state(self): check(self._lib.hk_reset_temporal_state(self.ctx), self.ctx, self._lib)
    def set_profiling(self, count_rays, time_passes): check(self._lib.hk_set_profiling(self.ctx, int(count_rays), int(time_passes)), self.ctx, self._lib)
    def set_tuning(self, key, value): check(self._lib.hk_set_tuning(self.ctx, int(key), int(value)), self.ctx, self._lib)
    def set_profiling_kernel(self, kernel): check(self._lib.hk_set_profiling_kernel(self.ctx, int(kernel)), self.ctx, self._lib)
    def set_keep_intermediates(self, keep): check(self._lib.hk_set_keep_intermediates(self.ctx, int(keep)), self.ctx, self._lib)

    def stats(self):
        s = L.FrameStats()
        check(self._lib.hk_get_stats(self.ctx, C.byref(s)), self.ctx, self._lib)
        return s

    def output_extent(self, which):
        w, h = C.c_uint32(), C.c_uint32()
        check(self._lib.hk_output_extent(self.ctx, which, C.byref(w), C.byref(h)), self.ctx, self._lib)
        return w.value, h.value

    def readback(self, which, out=None):
        bpp, dt, comps = L.OUT_FORMATS[which]
        w, h = self.output_extent(which)
        n = w * h
        if out is None:
            out = np.empty(n * bpp, np.uint8)
        check(self._lib.hk_readback(self.ctx, which, out.ctypes.data, n * bpp), self.ctx, self._lib)
        return view_plane(out[:n * bpp], which, h, w)

    def readback_into(self, which, host_ptr, nbytes):
        check(self._lib.hk_readback(self.ctx, which, host_ptr, nbytes), self.ctx, self._lib)

    def readback_async(self, which, pinned_host_ptr, nbytes):
        check(self._lib.hk_readback_async(self.ctx, which, pinned_host_ptr, nbytes), self.ctx, self._lib)

    def readback_wait(self):
        check(self._lib.hk_readback_wait(self.ctx), self.ctx, self._lib)

    # exact tiling under camera motion
    def set_motion_margin(self, pixels):
        check(self._lib.hk_context_set_motion_margin(self.ctx, int(pixels)), self.ctx, self._lib)

    def enable_tile_upscalers(self, enabled=True):
        check(self._lib.hk_context_enable_tile_upscalers(self.ctx, 1 if enabled else 0), self.ctx, self._lib)

    def halo_pull(self, source):
        check(self._lib.hk_halo_pull(self.ctx, source.ctx), self.ctx, self._lib)

    HALO_DESCRIPTOR_BYTES = 44 * 64 + 11 * 4

    def halo_export(self):
        """bytes of an hk_halo_descriptor (CUDA IPC handles of the reservoir planes + tile rectangles) for another process"""
        buf = (C.c_uint8 * self.HALO_DESCRIPTOR_BYTES)()
        check(self._lib.hk_halo_export(self.ctx, buf), self.ctx, self._lib)
        return bytes(buf)

    def halo_import(self, descriptor):
        buf = (C.c_uint8 * self.HALO_DESCRIPTOR_BYTES).from_buffer_copy(descriptor)
        peer = C.c_void_p()
        check(self._lib.hk_halo_import(self.ctx, buf, C.byref(peer)), self.ctx, self._lib)
        return peer

    def halo_pull_peer(self, peer):
        check(self._lib.hk_halo_pull_peer(self.ctx, peer), self.ctx, self._lib)

    # frame assembly across tiles / GPUs (hk_set_frame_target)
    def frame_alloc(self):
        p = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        check(self._lib.hk_frame_alloc(self.ctx, C.byref(p), handle), self.ctx, self._lib)
        return p.value, bytes(handle)

    def frame_open(self, handle):
        p = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        check(self._lib.hk_frame_open(self.ctx, buf, C.byref(p)), self.ctx, self._lib)
        return p.value

    def set_frame_target(self, device_ptr, pitch_pixels=None):
        check(self._lib.hk_set_frame_target(self.ctx, device_ptr, self.width if pitch_pixels is None else pitch_pixels), self.ctx, self._lib)

    def frame_read(self, device_ptr):
        out = np.empty(self.width * self.height * 8, np.uint8)
        check(self._lib.hk_frame_read(self.ctx, device_ptr, out.ctypes.data, out.size), self.ctx, self._lib)
        return view_plane(out, L.OUT_TONE_MAPPED, self.height, self.width)

    def upload_state(self, which, array):
        a = np.ascontiguousarray(array)
        check(self._lib.hk_upload_state(self.ctx, which, a.ctypes.data, a.nbytes), self.ctx, self._lib)

    def output_device_pointer(self, which=L.OUT_TONE_MAPPED):
        p, b = C.c_void_p(), C.c_size_t()
        check(self._lib.hk_get_output(self.ctx, which, C.byref(p), C.byref(b)), self.ctx, self._lib)
        return p.value, b.value

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, L.RAY)
        hits = np.zeros(len(rays), L.HIT)
        check(self._lib.hk_trace_rays(self.ctx, rays.ctypes.data, len(rays), hits.ctypes.data), self.ctx, self._lib)
        return hits


def view_plane(raw, which, rows, width):
    bpp, dt, comps = L.OUT_FORMATS[which]
    a = np.frombuffer(raw, dt) if not isinstance(dt, np.dtype) or dt.names is None else np.frombuffer(raw, dt)
    if comps == 1:
        return a.reshape(rows, width)
    return a.reshape(rows, width, comps)
