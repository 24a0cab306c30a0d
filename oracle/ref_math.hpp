// ref_math.hpp — float32 value types of the CPU oracle (TEST INFRASTRUCTURE ONLY).
// Restates the arithmetic the reference gets from GeometryBasics/StaticArrays, in the association
// order those packages generate (left-assoc sums, no FMA; compile with -ffp-contract=off):
//   dot   = (a1*b1 + a2*b2) + a3*b3            norm = sqrt((x*x + y*y) + z*z)
//   normalize(v) = v * (1/norm(v))             cross = textbook
//   M*v  row sums left-assoc                   min/max with Julia's -0/+0 and NaN rules
// Reference call sites: src/transformations.jl:132-144, src/bounds.jl:59-61, src/Trace.jl:98.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace ref {

static const float PI_F = 3.1415927f;       // Float32(pi)
static const float INF_F = std::numeric_limits<float>::infinity();

struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit V3(float a) : x(a), y(a), z(a) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float norm(V3 a) { return sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z); }
inline V3 normalize(V3 a) { float inv = 1.0f / norm(a); return V3(inv * a.x, inv * a.y, inv * a.z); }
inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// Julia's min/max for floats (NaN-propagating, -0.0 < +0.0).
inline float jl_min(float x, float y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    if (y < x) return y;
    if (x == y && std::signbit(y) && !std::signbit(x)) return y;
    return x;
}
inline float jl_max(float x, float y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    if (y > x) return y;
    if (x == y && !std::signbit(y) && std::signbit(x)) return y;
    return x;
}
inline float jl_clamp(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }  // Base.clamp

struct B3 {
    V3 lo, hi;
    B3() : lo(INF_F, INF_F, INF_F), hi(-INF_F, -INF_F, -INF_F) {}   // Bounds3() invalid config, bounds.jl:12
    B3(V3 a, V3 b) : lo(a), hi(b) {}
    explicit B3(V3 p) : lo(p), hi(p) {}
};
inline B3 unite(const B3& a, const B3& b) {                           // bounds.jl:59-61
    return B3(V3(jl_min(a.lo.x, b.lo.x), jl_min(a.lo.y, b.lo.y), jl_min(a.lo.z, b.lo.z)),
              V3(jl_max(a.hi.x, b.hi.x), jl_max(a.hi.y, b.hi.y), jl_max(a.hi.z, b.hi.z)));
}
inline V3 diagonal(const B3& b) { return b.hi - b.lo; }              // bounds.jl:80
inline float surface_area(const B3& b) {                              // bounds.jl:82-85
    V3 d = diagonal(b);
    return 2.0f * ((d.x * d.y + d.x * d.z) + d.y * d.z);
}
inline int maximum_extent(const B3& b) {                              // bounds.jl:112-120 (0-based here)
    V3 d = diagonal(b);
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
}
inline bool is_valid(const B3& b) {                                   // bounds.jl:30-32
    return b.lo.x != INF_F && b.lo.y != INF_F && b.lo.z != INF_F &&
           b.hi.x != -INF_F && b.hi.y != -INF_F && b.hi.z != -INF_F;
}
inline V3 offset(const B3& b, V3 p) {                                 // bounds.jl:134-143
    V3 o = p - b.lo;
    bool gx = b.hi.x > b.lo.x, gy = b.hi.y > b.lo.y, gz = b.hi.z > b.lo.z;
    if (!(gx || gy || gz)) return o;
    return V3(o.x / (gx ? b.hi.x - b.lo.x : 1.0f), o.y / (gy ? b.hi.y - b.lo.y : 1.0f),
              o.z / (gz ? b.hi.z - b.lo.z : 1.0f));
}

// Row-major 4x4; element (r,c) = m[4*r+c]. Transformation application, transformations.jl:132-144.
struct M4 { float m[16]; };
inline V3 xf_point(const float* m, V3 p) {
    float x = ((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3] * 1.0f;
    float y = ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7] * 1.0f;
    float z = ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11] * 1.0f;
    float w = ((m[12] * p.x + m[13] * p.y) + m[14] * p.z) + m[15] * 1.0f;
    if (w == 1.0f) return V3(x, y, z);
    return V3(x / w, y / w, z / w);
}
inline V3 xf_vector(const float* m, V3 v) {
    return V3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z,
              (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}
// normals use transpose(inv_m[1:3,1:3]) * n
inline V3 xf_normal(const float* inv_m, V3 n) {
    return V3((inv_m[0] * n.x + inv_m[4] * n.y) + inv_m[8] * n.z, (inv_m[1] * n.x + inv_m[5] * n.y) + inv_m[9] * n.z,
              (inv_m[2] * n.x + inv_m[6] * n.y) + inv_m[10] * n.z);
}

inline V3 face_forward(V3 n, V3 v) { return dot(n, v) < 0.0f ? -n : n; }          // Trace.jl:168
inline void coordinate_system(V3 v1, V3& v2, V3& v3) {                            // Trace.jl:139-146
    if (fabsf(v1.x) > fabsf(v1.y)) v2 = V3(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else v2 = V3(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v1, v2);
}

struct RGB {
    float r, g, b;
    RGB() : r(0), g(0), b(0) {}
    explicit RGB(float v) : r(v), g(v), b(v) {}
    RGB(float a, float b_, float c) : r(a), g(b_), b(c) {}
};
inline RGB operator+(RGB a, RGB b) { return RGB(a.r + b.r, a.g + b.g, a.b + b.b); }
inline RGB operator-(RGB a, float f) { return RGB(a.r - f, a.g - f, a.b - f); }
inline RGB operator*(RGB a, RGB b) { return RGB(a.r * b.r, a.g * b.g, a.b * b.b); }
inline RGB operator*(RGB a, float f) { return RGB(a.r * f, a.g * f, a.b * f); }
inline RGB operator*(float f, RGB a) { return RGB(a.r * f, a.g * f, a.b * f); }
inline RGB operator/(RGB a, float f) { return RGB(a.r / f, a.g / f, a.b / f); }
inline bool is_black(RGB a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }   // spectrum.jl:55
inline bool has_nan(RGB a) { return std::isnan(a.r) || std::isnan(a.g) || std::isnan(a.b); }
inline float to_Y(RGB s) { return (0.212671f * s.r + 0.715160f * s.g) + 0.072169f * s.b; }   // spectrum.jl:64-66
inline RGB clamp0(RGB a) { return RGB(jl_clamp(a.r, 0.0f, INF_F), jl_clamp(a.g, 0.0f, INF_F), jl_clamp(a.b, 0.0f, INF_F)); }
inline void rgb_to_xyz(RGB c, float out[3]) {                                      // spectrum.jl:8-14
    out[0] = (0.412453f * c.r + 0.357580f * c.g) + 0.180423f * c.b;
    out[1] = (0.212671f * c.r + 0.715160f * c.g) + 0.072169f * c.b;
    out[2] = (0.019334f * c.r + 0.119193f * c.g) + 0.950227f * c.b;
}

// Counter-based RNG that stands in for the reference's rand() (DESIGN.md "RNG"): splitmix64 finaliser
// over (seed, a, b, c); 24-bit mantissa in [0,1).
inline float rng_uniform(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) {
    uint64_t z = seed ^ ((uint64_t)a * 0x9E3779B97F4A7C15ull) ^ ((((uint64_t)b << 32) | (uint64_t)c) * 0xD1B54A32D192ED03ull);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * 5.9604644775390625e-8f;   // 2^-24
}

}  // namespace ref
