"""The configurations BASELINE.json names, as concrete inputs:
 C1 caustic_glass  = docs/code/caustic_glass.jl      C2 shadows  = docs/code/spheres.jl
 C4 caustic_moving = docs/code/caustic_moving.jl     C3 / C5     = synthetic tessellated scenes ("tess-1M", "tess-10M")
Each builder returns (scene, camera, integrator_kwargs)."""
import os

import numpy as np

from .geometry import Bounds2, Point2f, Transformation, coordinate_system, look_at, normalize, translate, _v3, f32
from .render import Film, LanczosSincFilter, PerspectiveCamera
from .scene import (BVHAccel, ConstantTexture, GeometricPrimitive, GlassMaterial, MatteMaterial, MirrorMaterial,
                    PlasticMaterial, PointLight, PrimitiveBatch, RGBSpectrum, Scene, ShapeCore, Sphere, SpotLight,
                    TriangleMesh, TriangleSet, create_triangle_mesh, load_triangle_mesh)

ASSET_PLY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets", "caustic-glass.ply")


def _film(res_x, res_y=None, filename=None):
    res_y = res_x if res_y is None else res_y
    return Film(Point2f(res_x, res_y), Bounds2(Point2f(0.0), Point2f(1.0)), LanczosSincFilter(Point2f(1.0), 3.0),
                1.0, 1.0, filename)


def shadows(resolution=1024, filename=None):
    """docs/code/spheres.jl:4-103 (C2).  SPPMIntegrator(camera, 0.025, 5, 100)."""
    red = MatteMaterial(ConstantTexture(RGBSpectrum(0.796, 0.235, 0.2)), ConstantTexture(0.0))
    blue = MatteMaterial(ConstantTexture(RGBSpectrum(0.251, 0.388, 0.847)), ConstantTexture(0.0))
    white = MatteMaterial(ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(0.0))
    mirror = MirrorMaterial(ConstantTexture(RGBSpectrum(1.0)))
    glass = GlassMaterial(ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(0.0),
                          ConstantTexture(0.0), ConstantTexture(1.5), True)
    prims = []
    for pos, r, mat in (((0.3, 0.11, -2.2), 0.1, glass), ((0.2, 0.11, -2.6), 0.1, blue), ((0.7, 0.31, -2.8), 0.3, mirror),
                        ((0.7, 0.11, -2.3), 0.1, red)):
        prims.append(GeometricPrimitive(Sphere(ShapeCore(translate(_v3(pos)), False), r, 360.0), mat))
    tris = create_triangle_mesh(
        ShapeCore(translate(_v3((0, 0, -2))), False), 4, [1, 2, 3, 1, 4, 3, 2, 3, 5, 6, 5, 3], 6,
        [(0, 0, 0), (0, 0, -1), (1, 0, -1), (1, 0, 0), (0, 1, -1), (1, 1, -1)],
        [(0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 0, 1), (0, 0, 1)])
    for t, mat in zip(tris, (mirror, mirror, white, white)):
        prims.append(GeometricPrimitive(t, mat))
    bvh = BVHAccel(prims, 1)
    lights = [PointLight(translate(_v3((-1, 1, 0))), RGBSpectrum(25.0))]
    scene = Scene(lights, bvh)
    film = _film(resolution, filename=filename)
    camera = PerspectiveCamera(look_at(_v3((0, 15, 50)), _v3((0, 0, -2)), _v3((0, 1, 0))),
                               Bounds2(Point2f(-1.0), Point2f(1.0)), 0.0, 1.0, 0.0, 1e6, 90.0, film)
    return scene, camera, dict(initial_search_radius=0.025, max_depth=5, n_iterations=100)


def _spot_light_to_world(frm):
    """docs/code/caustic_glass.jl:49-64"""
    frm, to = _v3(frm), _v3((-5, 0, 5))
    direction = normalize(to - frm)
    direction, du, dv = coordinate_system(direction)
    m = np.eye(4, dtype=np.float32)
    m[0, :3], m[1, :3], m[2, :3] = du, dv, direction
    dir_to_z = Transformation(m)
    return translate(_v3((4.5, 0, -101))) * translate(frm) * dir_to_z.inv()


def _caustic_bvh(eta, ply=ASSET_PLY, builder="reference", max_node_primitives=1):
    glass = GlassMaterial(ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(0.0),
                          ConstantTexture(0.0), ConstantTexture(eta), True)
    plastic = PlasticMaterial(ConstantTexture(RGBSpectrum(0.6399999857)), ConstantTexture(RGBSpectrum(0.1000000015)),
                              ConstantTexture(0.010408001), True)
    _, triangles = load_triangle_mesh(ply, ShapeCore(translate(_v3((5, -1.49, -100))), False))
    floor = create_triangle_mesh(ShapeCore(translate(_v3((-10, 0, -87))), False), 2, [1, 2, 3, 1, 4, 3], 4,
                                 [(0, 0, 0), (0, 0, -30), (30, 0, -30), (30, 0, 0)], [(0, 1, 0)] * 4)
    prims = [PrimitiveBatch(triangles, glass)] + [GeometricPrimitive(t, plastic) for t in floor]
    return BVHAccel(prims, max_node_primitives, builder=builder)


def caustic_glass(resolution=256, max_depth=5, filename=None, ply=ASSET_PLY, builder="reference", max_node_primitives=1):
    """docs/code/caustic_glass.jl:6-95 (C1).  SPPMIntegrator(camera, 0.075, ray_depth, 100, -1)."""
    bvh = _caustic_bvh(1.25, ply, builder, max_node_primitives)
    lights = [SpotLight(_spot_light_to_world((0, 2, 0)), RGBSpectrum(60.0), 30.0, 20.0)]
    scene = Scene(lights, bvh)
    film = _film(resolution, filename=filename)
    camera = PerspectiveCamera(look_at(_v3((0, 150, 150)), _v3((-3, 0, -91)), _v3((0, 1, 0))),
                               Bounds2(Point2f(-1.0), Point2f(1.0)), 0.0, 1.0, 0.0, 1e6, 90.0, film)
    return scene, camera, dict(initial_search_radius=0.075, max_depth=max_depth, n_iterations=100)


def caustic_moving(shift=0.0, resolution=1024, filename=None, ply=ASSET_PLY, bvh=None, builder="reference",
                   max_node_primitives=1):
    """One frame of docs/code/caustic_moving.jl:5-103 (C4).  SPPMIntegrator(camera, 0.055, 5, 25, 1_250_000)."""
    bvh = bvh or _caustic_bvh(1.2, ply, builder, max_node_primitives)
    lights = [PointLight(translate(_v3((2.5, 10, -100))), RGBSpectrum(1.0) * 20.0),
              SpotLight(_spot_light_to_world((0, 0.5 + shift, 0)), RGBSpectrum(0.988235, 0.972549, 0.57647) * 60.0, 30.0, 20.0)]
    scene = Scene(lights, bvh)
    film = _film(resolution, filename=filename)
    camera = PerspectiveCamera(look_at(_v3((0, 150, 150)), _v3((-3, 0, -91)), _v3((0, 1, 0))),
                               Bounds2(Point2f(-1.0), Point2f(1.0)), 0.0, 1.0, 0.0, 1e6, 90.0, film)
    return scene, camera, dict(initial_search_radius=0.055, max_depth=5, n_iterations=25, photons_per_iteration=1_250_000)


def _uv_sphere(center, radius, stacks, slices):
    """Closed UV sphere: 2 * slices * (stacks - 1) triangles, outward vertex normals, 1-based indices."""
    th = (np.arange(stacks + 1, dtype=np.float64) / stacks) * np.pi
    ph = (np.arange(slices, dtype=np.float64) / slices) * 2.0 * np.pi
    T, P = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], axis=-1)
    verts = (np.asarray(center, dtype=np.float64) + radius * n).reshape(-1, 3).astype(np.float32)
    normals = n.reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(stacks), np.arange(slices), indexing="ij")
    a = i * slices + j
    b = i * slices + (j + 1) % slices
    c = (i + 1) * slices + j
    d = (i + 1) * slices + (j + 1) % slices
    t1 = np.stack([a, c, b], axis=-1)[1:]            # skip the degenerate fan triangles at the north pole
    t2 = np.stack([b, c, d], axis=-1)[:-1]           # ... and at the south pole
    idx = np.concatenate([t1.reshape(-1, 3), t2.reshape(-1, 3)], axis=0)
    return verts, normals, (idx + 1).astype(np.uint32)


def _heightfield(cells, x0=-10.0, x1=10.0, z0=-30.0, z1=-10.0):
    xs = np.linspace(x0, x1, cells + 1)
    zs = np.linspace(z0, z1, cells + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = 0.15 * np.sin(1.7 * X) * np.cos(1.3 * Z)
    dYdx = 0.15 * 1.7 * np.cos(1.7 * X) * np.cos(1.3 * Z)
    dYdz = -0.15 * 1.3 * np.sin(1.7 * X) * np.sin(1.3 * Z)
    n = np.stack([-dYdx, np.ones_like(Y), -dYdz], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    normals = n.reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    a = i * (cells + 1) + j
    b = a + 1
    c = a + (cells + 1)
    d = c + 1
    idx = np.concatenate([np.stack([a, b, c], axis=-1).reshape(-1, 3), np.stack([b, d, c], axis=-1).reshape(-1, 3)], axis=0)
    return verts, normals, (idx + 1).astype(np.uint32)


def tessellated(cells=600, stacks=266, slices=264, res=(1920, 1080), window=((-50.0, -28.125), (50.0, 28.125)),
                filename=None, builder="reference", max_node_primitives=1):
    """Synthetic "tess-1M" (C3, defaults: 999 840 triangles) / "tess-10M" (C5: cells=1900, stacks=835, slices=834,
    res=(4096, 4096), window=((-50,-50),(50,50)))  — SURVEY.md §8d.  WhittedIntegrator depth 5 / 8."""
    ident = ShapeCore(Transformation(), False)
    matte = MatteMaterial(ConstantTexture(RGBSpectrum(0.8)), ConstantTexture(0.0))
    glass = GlassMaterial(ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(RGBSpectrum(1.0)), ConstantTexture(0.0),
                          ConstantTexture(0.0), ConstantTexture(1.5), True)
    mirror = MirrorMaterial(ConstantTexture(RGBSpectrum(0.9)))
    prims = []
    v, n, idx = _heightfield(cells)
    prims.append(PrimitiveBatch(TriangleSet(ident, TriangleMesh(ident.object_to_world, len(idx), idx.reshape(-1), len(v), v, n)), matte))
    for center, mat in (((-2.5, 2.0, -20.0), glass), ((2.5, 2.0, -20.0), mirror)):
        v, n, idx = _uv_sphere(center, 2.0, stacks, slices)
        prims.append(PrimitiveBatch(TriangleSet(ident, TriangleMesh(ident.object_to_world, len(idx), idx.reshape(-1), len(v), v, n)), mat))
    bvh = BVHAccel(prims, max_node_primitives, builder=builder)
    scene = Scene([PointLight(translate(_v3((0, 12, -10))), RGBSpectrum(400.0))], bvh)
    film = _film(res[0], res[1], filename=filename)
    camera = PerspectiveCamera(look_at(_v3((0, 14, 15)), _v3((-16, -18, -20)), _v3((0, 1, 0))),
                               Bounds2(Point2f(*window[0]), Point2f(*window[1])), 0.0, 1.0, 0.0, 1e6, 90.0, film)
    return scene, camera, dict(max_depth=5)
