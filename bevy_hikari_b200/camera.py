"""What Bevy hands to the render graph each frame for one camera: the `View` / `PreviousView` / `Lights` uniforms
(bevy_render 0.9 `ViewUniform`, bevy_pbr 0.9 `GpuLights`; src/view.rs:47-73) restated as float32 numpy, column-major
like glam.  These are INPUTS to the hot path (bind group 0), not part of it."""
import ctypes as C
import math

import numpy as np

from . import layout as L

F = np.float32


def perspective_infinite_reverse_rh(fov_y, aspect, z_near):
    """glam Mat4::perspective_infinite_reverse_rh; returned as [col][row]."""
    f = F(1.0) / F(math.tan(0.5 * fov_y))
    m = np.zeros((4, 4), F)
    m[0, 0] = f / F(aspect)
    m[1, 1] = f
    m[2, 3] = F(-1.0)
    m[3, 2] = F(z_near)
    return m


def orthographic_reverse_rh(half_height, aspect, z_near, z_far):
    """bevy OrthographicProjection::get_projection_matrix: glam Mat4::orthographic_rh(left, right, bottom, top, far, near)
    (near and far swapped for reverse Z); returned as [col][row].  projection[3][3] == 1 is what light.wgsl:1040 tests."""
    hw = half_height * aspect
    left, right, bottom, top = -hw, hw, -half_height, half_height
    near, far = z_far, z_near                       # the swapped arguments
    rw, rh, r = 1.0 / (right - left), 1.0 / (top - bottom), 1.0 / (near - far)
    m = np.zeros((4, 4), F)
    m[0, 0] = F(rw + rw)
    m[1, 1] = F(rh + rh)
    m[2, 2] = F(r)
    m[3, 0] = F(-(left + right) * rw)
    m[3, 1] = F(-(top + bottom) * rh)
    m[3, 2] = F(r * near)
    m[3, 3] = F(1.0)
    return m


def look_at(eye, target, up=(0.0, 1.0, 0.0)):
    """Transform::from_translation(eye).looking_at(target, up) -> camera world matrix [col][row]."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    forward = eye - target
    forward /= np.linalg.norm(forward)
    right = np.cross(up, forward)
    right /= np.linalg.norm(right)
    upv = np.cross(forward, right)
    m = np.zeros((4, 4), F)
    m[0, :3], m[1, :3], m[2, :3], m[3, :3] = right, upv, forward, eye
    m[3, 3] = 1
    return m


def _mul(a, b):
    # (a*b) for [col][row] storage: as ordinary matrices A = a.T, B = b.T, so (A@B).T = b@a in stored form
    return (b.astype(np.float64) @ a.astype(np.float64)).astype(F)


def _inv(a):
    return np.linalg.inv(a.astype(np.float64).T).T.astype(F)


def make_view(camera_world, projection, width, height):
    v = L.View()
    inverse_view = _inv(camera_world)
    view_proj = _mul(projection, inverse_view)
    for name, m in (("view_proj", view_proj), ("inverse_view_proj", _mul(camera_world, _inv(projection))),
                    ("view", camera_world), ("inverse_view", inverse_view), ("projection", projection),
                    ("inverse_projection", _inv(projection))):
        getattr(v, name)[:] = [float(x) for x in m.reshape(16)]
    v.world_position[:] = [float(x) for x in camera_world[3, :3]]
    v.viewport[:] = [0.0, 0.0, float(width), float(height)]
    return v


def make_previous_view(view):
    p = L.PreviousView()
    C.memmove(p.view_proj, view.view_proj, 64)
    C.memmove(p.inverse_view_proj, view.inverse_view_proj, 64)
    return p


def make_lights(sun_illuminance=None, sun_color=(1.0, 1.0, 1.0), sun_direction_to_light=(0.0, 0.0, 0.0),
                ambient_color=(1.0, 1.0, 1.0), ambient_brightness=0.05):
    """bevy_pbr 0.9 prepare_lights: directional colour = linear rgb * illuminance * exposure with
    exposure = 1 / (2^EV100 * 1.2), EV100 for aperture 4, shutter 1/250, ISO 100; ambient = colour * brightness."""
    l = L.Lights()
    if sun_illuminance is not None:
        ev100 = math.log2(4.0 * 4.0 / (1.0 / 250.0) * 100.0 / 100.0)
        exposure = 1.0 / (2.0 ** ev100 * 1.2)
        k = sun_illuminance * exposure
        l.directional_color[:] = [sun_color[0] * k, sun_color[1] * k, sun_color[2] * k, k]
        d = np.asarray(sun_direction_to_light, np.float64)
        d = d / np.linalg.norm(d)
        l.direction_to_light[:] = [float(x) for x in d.astype(F)]
    l.ambient_color[:] = [ambient_color[0] * ambient_brightness, ambient_color[1] * ambient_brightness,
                          ambient_color[2] * ambient_brightness, ambient_brightness]
    return l


def euler_xyz_back(x, y, z):
    """direction_to_light = Transform::from_rotation(Quat::from_euler(XYZ, x, y, z)).back() (= rotation * +Z)."""
    cx, sx, cy, sy, cz, sz = math.cos(x), math.sin(x), math.cos(y), math.sin(y), math.cos(z), math.sin(z)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (rx @ ry @ rz) @ np.array([0.0, 0.0, 1.0])
