"""Host-side value types mirroring Trace.jl's L0 layer (src/transformations.jl, src/bounds.jl), float32 and in the
reference's association order.  These run on the host exactly as they do in the reference (camera / light / shape
matrices are built once per scene); the GPU receives the finished matrices (include/trace_cuda.h: trace_camera).

Quirks reproduced on purpose: Q1 `t1*t2` multiplies the inverse matrices in the SAME order
(transformations.jl:20-22); Q2 `perspective` fills its matrix without the transpose every other constructor uses
(transformations.jl:119-130).
"""
import math

import numpy as np

f32 = np.float32


def _v3(x):
    a = np.asarray(x, dtype=np.float32).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, 3)
    assert a.size == 3
    return a


Point3f = _v3
Vec3f = _v3
Normal3f = _v3


def Point2f(*x):
    a = np.asarray(x if len(x) > 1 else x[0], dtype=np.float32).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, 2)
    assert a.size == 2
    return a


def dot(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def cross(a, b):
    return np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]),
                     f32(a[0] * b[1]) - f32(a[1] * b[0])], dtype=np.float32)


def norm(a):
    return f32(np.sqrt(f32(f32(f32(a[0] * a[0]) + f32(a[1] * a[1])) + f32(a[2] * a[2]))))


def normalize(a):
    a = np.asarray(a, dtype=np.float32)
    return (f32(1.0) / norm(a)) * a


def coordinate_system(v1):
    """src/Trace.jl:139-146"""
    v1 = _v3(v1)
    if abs(v1[0]) > abs(v1[1]):
        v2 = np.array([-v1[2], 0, v1[0]], dtype=np.float32) / f32(np.sqrt(f32(f32(v1[0] * v1[0]) + f32(v1[2] * v1[2]))))
    else:
        v2 = np.array([0, v1[2], -v1[1]], dtype=np.float32) / f32(np.sqrt(f32(f32(v1[1] * v1[1]) + f32(v1[2] * v1[2]))))
    return v1, v2, cross(v1, v2)


def _matmul4(a, b):
    """StaticArrays 4x4 product: each element is a left-associated float32 sum of products."""
    out = np.zeros((4, 4), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            s = f32(a[i, 0] * b[0, j])
            for k in range(1, 4):
                s = f32(s + f32(a[i, k] * b[k, j]))
            out[i, j] = s
    return out


def _inv4(m):
    # StaticArrays' closed-form 4x4 inverse is not restated (dependency absent): float64 inverse rounded to float32.
    return np.linalg.inv(m.astype(np.float64)).astype(np.float32)


class Transformation:
    """src/transformations.jl:1-22.  `m` and `inv_m` are row-major 4x4 float32 (element [r, c])."""

    __slots__ = ("m", "inv_m")

    def __init__(self, m=None, inv_m=None):
        if m is None:
            m = np.eye(4, dtype=np.float32)
            inv_m = np.eye(4, dtype=np.float32)
        m = np.asarray(m, dtype=np.float32).reshape(4, 4)
        self.m = m
        self.inv_m = _inv4(m) if inv_m is None else np.asarray(inv_m, dtype=np.float32).reshape(4, 4)

    def __mul__(self, other):                       # :20-22 — inverse multiplied in the same order (Q1)
        return Transformation(_matmul4(self.m, other.m), _matmul4(self.inv_m, other.inv_m))

    def inv(self):                                   # :12
        return Transformation(self.inv_m, self.m)

    def transpose(self):
        return Transformation(self.m.T.copy(), self.inv_m.T.copy())

    def __eq__(self, o):
        return bool(np.all(self.m == o.m) and np.all(self.inv_m == o.inv_m))

    # application, :132-144
    def point(self, p):
        p = _v3(p)
        m = self.m
        out = np.zeros(4, dtype=np.float32)
        for i in range(4):
            out[i] = f32(f32(f32(f32(m[i, 0] * p[0]) + f32(m[i, 1] * p[1])) + f32(m[i, 2] * p[2])) + f32(m[i, 3] * f32(1)))
        if out[3] == 1:
            return out[:3].copy()
        return (out[:3] / out[3]).astype(np.float32)

    def vector(self, v):
        v = _v3(v)
        m = self.m
        return np.array([f32(f32(f32(m[i, 0] * v[0]) + f32(m[i, 1] * v[1])) + f32(m[i, 2] * v[2])) for i in range(3)],
                        dtype=np.float32)

    def normal(self, n):
        n = _v3(n)
        im = self.inv_m
        return np.array([f32(f32(f32(im[0, i] * n[0]) + f32(im[1, i] * n[1])) + f32(im[2, i] * n[2])) for i in range(3)],
                        dtype=np.float32)

    def points(self, pts):
        """Vectorised point transform with the same association order (used for whole meshes)."""
        p = np.asarray(pts, dtype=np.float32).reshape(-1, 3)
        m = self.m
        cols = []
        for i in range(4):
            s = (m[i, 0] * p[:, 0]).astype(np.float32)
            s = (s + (m[i, 1] * p[:, 1]).astype(np.float32)).astype(np.float32)
            s = (s + (m[i, 2] * p[:, 2]).astype(np.float32)).astype(np.float32)
            s = (s + f32(m[i, 3] * f32(1))).astype(np.float32)
            cols.append(s)
        w = cols[3]
        out = np.stack(cols[:3], axis=1)
        need = w != 1
        if np.any(need):
            out[need] = (out[need] / w[need, None]).astype(np.float32)
        return out

    def bounds(self, b):                              # :141-143: union of the 8 transformed corners
        r = Bounds3()
        for c in range(8):
            corner = np.array([b.p_max[0] if c & 1 else b.p_min[0], b.p_max[1] if c & 2 else b.p_min[1],
                               b.p_max[2] if c & 4 else b.p_min[2]], dtype=np.float32)
            r = r.union(Bounds3(self.point(corner)))
        return r

    def swaps_handedness(self):                       # :163-165
        return float(np.linalg.det(self.m[:3, :3].astype(np.float64))) < 0

    def __call__(self, x):
        if isinstance(x, Bounds3):
            return self.bounds(x)
        return self.point(x)


def translate(delta):
    d = _v3(delta)
    m = np.eye(4, dtype=np.float32)
    mi = np.eye(4, dtype=np.float32)
    m[:3, 3] = d
    mi[:3, 3] = -d
    return Transformation(m, mi)


def scale(x, y, z):
    m = np.diag(np.array([x, y, z, 1], dtype=np.float32))
    mi = np.diag(np.array([f32(1) / f32(x), f32(1) / f32(y), f32(1) / f32(z), 1], dtype=np.float32))
    return Transformation(m, mi)


def _deg2rad(t):
    return f32(f32(t) * f32(f32(np.pi) / f32(180)))


def rotate_x(theta):
    s, c = f32(np.sin(_deg2rad(theta))), f32(np.cos(_deg2rad(theta)))
    m = np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    return Transformation(m, m.T.copy())


def rotate_y(theta):
    s, c = f32(np.sin(_deg2rad(theta))), f32(np.cos(_deg2rad(theta)))
    m = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    return Transformation(m, m.T.copy())


def rotate_z(theta):
    s, c = f32(np.sin(_deg2rad(theta))), f32(np.cos(_deg2rad(theta)))
    m = np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    return Transformation(m, m.T.copy())


def look_at(position, target, up=(0, 1, 0)):
    """src/transformations.jl:105-117"""
    position, target, up = _v3(position), _v3(target), _v3(up)
    z_axis = normalize(position - target)
    x_axis = normalize(cross(up, z_axis))
    y_axis = cross(z_axis, x_axis)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0] = x_axis
    m[:3, 1] = y_axis
    m[:3, 2] = z_axis
    return translate(position) * Transformation(m, m.T.copy())


def perspective(fov, near, far):
    """src/transformations.jl:119-130.  Mat4f(...) is filled column-major and NOT transposed (Q2)."""
    fov, near, far = f32(fov), f32(near), f32(far)
    a = f32(far / f32(far - near))
    b = f32(f32(f32(-far) * near) / f32(far - near))
    flat = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, a, b, 0, 0, 1, 0], dtype=np.float32)
    p = flat.reshape(4, 4).T.copy()          # column-major fill => element [r, c] = flat[4 * c + r]
    inv_tan = f32(f32(1) / f32(np.tan(f32(_deg2rad(fov) / f32(2)))))
    return scale(inv_tan, inv_tan, f32(1)) * Transformation(p)


class Bounds2:
    __slots__ = ("p_min", "p_max")

    def __init__(self, p_min=None, p_max=None):
        if p_min is None:
            p_min, p_max = Point2f(np.inf), Point2f(-np.inf)
        self.p_min = Point2f(p_min)
        self.p_max = Point2f(p_min if p_max is None else p_max)

    def __eq__(self, o):
        return bool(np.all(self.p_min == o.p_min) and np.all(self.p_max == o.p_max))

    def __len__(self):                                 # bounds.jl:34-37
        d = np.ceil(self.p_max - self.p_min + f32(1))
        return int(d[0] * d[1])

    def __iter__(self):                                # bounds.jl:39-47
        d = self.p_max - self.p_min + f32(1)
        for j in range(len(self)):
            yield self.p_min + np.array([j % d[0], j // d[0]], dtype=np.float32)

    def intersect(self, o):
        return Bounds2(np.maximum(self.p_min, o.p_min), np.minimum(self.p_max, o.p_max))

    def diagonal(self):
        return self.p_max - self.p_min

    def area(self):                                    # bounds.jl:87-90
        d = self.p_max - self.p_min
        return f32(d[0] * d[1])

    def inclusive_sides(self):                         # bounds.jl:96-98
        return [abs(b1 - (b0 - f32(1))) for b1, b0 in zip(self.p_max, self.p_min)]

    def __repr__(self):
        return f"Bounds2({self.p_min.tolist()}, {self.p_max.tolist()})"


class Bounds3:
    __slots__ = ("p_min", "p_max")

    def __init__(self, p_min=None, p_max=None):
        if p_min is None:
            self.p_min, self.p_max = _v3(np.inf), _v3(-np.inf)
        else:
            self.p_min = _v3(p_min)
            self.p_max = _v3(p_min if p_max is None else p_max)

    def union(self, o):
        return Bounds3(np.minimum(self.p_min, o.p_min), np.maximum(self.p_max, o.p_max))

    def approx(self, o, rtol=math.sqrt(np.finfo(np.float32).eps)):
        return bool(np.allclose(self.p_min, o.p_min, rtol=rtol, atol=0) and np.allclose(self.p_max, o.p_max, rtol=rtol, atol=0))

    def __eq__(self, o):
        return bool(np.all(self.p_min == o.p_min) and np.all(self.p_max == o.p_max))

    def as6(self):
        return np.concatenate([self.p_min, self.p_max]).astype(np.float32)

    def __repr__(self):
        return f"Bounds3({self.p_min.tolist()}, {self.p_max.tolist()})"
