// ref_shading.hpp — CPU ORACLE (test infrastructure only): BxDFs, BSDF container, materials, lights, warps.
// Follows src/reflection/{bxdf,specular,lambertian,microfacet}.jl, src/materials/{bsdf,material}.jl,
// src/lights/{light,point,spot}.jl, src/Trace.jl:48-126. Quirks kept: Q6 (no eta^2 scaling),
// Q18 (sin_phi -> 1 at the pole, Oren-Nayar abs-of-Bool, Float64 2pi), Q24 (gating on ng), Q27 (microfacet pdf).
#pragma once
#include "ref_internal.hpp"

namespace ref {

enum : uint8_t { BSDF_NONE = 0, BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8,
                 BSDF_SPECULAR = 16, BSDF_ALL = 31 };                                  // bxdf.jl:1-7
enum LobeKind { L_LAMBERT, L_SPEC_REFL, L_SPEC_TRANS, L_FRESNEL_SPEC, L_MICRO_REFL, L_MICRO_TRANS, L_OREN_NAYAR };

struct Lobe {
    int kind;
    uint8_t type;
    RGB r, t;
    float eta_a, eta_b;        // transmission lobes / FresnelSpecular
    int fresnel;               // 0 FresnelNoOp, 1 FresnelDielectric(fi, ft)
    float fi, ft;
    float ax, ay;              // TrowbridgeReitz alphas (already clamped to >= 1e-3)
    float on_a, on_b;          // Oren-Nayar
};

struct BSDF {                  // materials/bsdf.jl:6-51
    float eta;
    V3 ng, ns, ss, ts;
    int n;
    Lobe lobes[8];
};

inline bool matches(const Lobe& l, uint8_t flags) { return (l.type & flags) == l.type; }   // bxdf.jl:9-11

inline float pow2(float x) { return x * x; }
inline float pow4(float x) { double d = (double)x; d = d * d; return (float)(d * d); }    // Float32^4 goes through Float64

// ---- Trace.jl:109-121
inline float cos_t(V3 w) { return w.z; }
inline float sin_t2(V3 w) { return jl_max(0.0f, 1.0f - cos_t(w) * cos_t(w)); }
inline float sin_t(V3 w) { return sqrtf(sin_t2(w)); }
inline float tan_t(V3 w) { return sin_t(w) / cos_t(w); }
inline float cos_p(V3 w) { float s = sin_t(w); return s == 0.0f ? 1.0f : jl_clamp(w.x / s, -1.0f, 1.0f); }
inline float sin_p(V3 w) { float s = sin_t(w); return s == 0.0f ? 1.0f : jl_clamp(w.y / s, -1.0f, 1.0f); }   // Q18
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }

// ---- warps, Trace.jl:48-96
inline void concentric_sample_disk(float u0, float u1, float& ox, float& oy) {
    float x = 2.0f * u0 - 1.0f, y = 2.0f * u1 - 1.0f;
    if (x == 0.0f && y == 0.0f) { ox = 0.0f; oy = 0.0f; return; }
    float r, th;
    if (fabsf(x) > fabsf(y)) { r = x; th = (y / x) * PI_F / 4.0f; }
    else { r = y; th = PI_F / 2.0f - (x / y) * PI_F / 4.0f; }
    ox = r * cosf(th); oy = r * sinf(th);
}
inline V3 cosine_sample_hemisphere(float u0, float u1) {
    float dx, dy;
    concentric_sample_disk(u0, u1, dx, dy);
    float z = sqrtf(jl_max(0.0f, 1.0f - dx * dx - dy * dy));
    return V3(dx, dy, z);
}
inline V3 uniform_sample_sphere(float u0, float u1) {
    float z = 1.0f - 2.0f * u0;
    float r = sqrtf(jl_max(0.0f, 1.0f - z * z));
    float phi = 2.0f * PI_F * u1;
    return V3(r * cosf(phi), r * sinf(phi), z);
}
inline V3 uniform_sample_cone(float u0, float u1, float cmax) {
    float c = 1.0f - u0 + u0 * cmax;
    float s = sqrtf(1.0f - c * c);
    float phi = u1 * 2.0f * PI_F;
    return V3(cosf(phi) * s, sinf(phi) * s, c);
}
inline float uniform_sphere_pdf() { return 1.0f / (4.0f * PI_F); }
inline float uniform_cone_pdf(float cmax) { return 1.0f / (2.0f * PI_F * (1.0f - cmax)); }

// ---- bxdf.jl:52-95
inline bool refract(V3 wi, V3 n, float eta, V3& wt) {
    float ci = dot(n, wi);
    float s2i = jl_max(0.0f, 1.0f - ci * ci);
    float s2t = (eta * eta) * s2i;
    if (s2t >= 1.0f) { wt = V3(0.0f); return false; }
    float ct = sqrtf(1.0f - s2t);
    wt = (-eta) * wi + (eta * ci - ct) * n;
    return true;
}
inline float fresnel_dielectric(float ci, float ei, float et) {
    ci = jl_clamp(ci, -1.0f, 1.0f);
    if (ci <= 0.0f) { float t = ei; ei = et; et = t; ci = fabsf(ci); }
    float si = sqrtf(jl_max(0.0f, 1.0f - ci * ci));
    float st = si * ei / et;
    if (st >= 1.0f) return 1.0f;
    float ct = sqrtf(jl_max(0.0f, 1.0f - st * st));
    float rpar = (et * ci - ei * ct) / (et * ci + ei * ct);
    float rper = (ei * ci - et * ct) / (ei * ci + et * ct);
    return 0.5f * (rpar * rpar + rper * rper);
}
inline float lobe_fresnel(const Lobe& l, float c) { return l.fresnel ? fresnel_dielectric(c, l.fi, l.ft) : 1.0f; }

// ---- microfacet.jl:53-201
inline float tr_lambda(const Lobe& l, V3 w) {
    float th = fabsf(tan_t(w));
    if (std::isinf(th)) return 0.0f;
    float a = sqrtf(pow2(cos_p(w)) * pow2(l.ax) + pow2(sin_p(w)) * pow2(l.ay));
    float a2t2 = pow2(a * th);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
inline float tr_G1(const Lobe& l, V3 w) { return 1.0f / (1.0f + tr_lambda(l, w)); }
inline float tr_G(const Lobe& l, V3 wo, V3 wi) { return 1.0f / (1.0f + tr_lambda(l, wo) + tr_lambda(l, wi)); }
inline float tr_D(const Lobe& l, V3 w) {
    float t2 = pow2(tan_t(w));
    if (std::isinf(t2)) return 0.0f;
    float c4 = pow4(cos_t(w));
    float e = (pow2(cos_p(w)) / pow2(l.ax) + pow2(sin_p(w)) / pow2(l.ay)) * t2;
    return 1.0f / (PI_F * l.ax * l.ay * c4 * pow2(1.0f + e));
}
inline float tr_pdf(const Lobe& l, V3 wo, V3 wh) {               // :107-110 (sample_visible_area = true)
    return tr_D(l, wh) * tr_G1(l, wo) * fabsf(dot(wo, wh)) / fabsf(cos_t(wo));
}
inline void tr_sample11(float cth, float u1, float u2, float& sx, float& sy) {   // :112-153
    if (cth > 0.9999f) {
        float r = sqrtf(u1 / (1.0f - u1));
        double phi = 6.28318530718 * (double)u2;
        sx = (float)((double)r * cos(phi)); sy = (float)((double)r * sin(phi));
        return;
    }
    float sth = sqrtf(jl_max(0.0f, 1.0f - cth * cth));
    float tth = sth / cth;
    float a = 1.0f / tth;
    float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
    a = 2.0f * u1 / g1 - 1.0f;
    float tmp = 1.0f / (a * a - 1.0f);
    if (tmp > 1e10f) tmp = 1e10f;
    float b = tth, b2 = b * b;
    float d = sqrtf(jl_max(0.0f, b2 * (tmp * tmp) - (a * a - b2) * tmp));
    float x1 = b * tmp - d, x2 = b * tmp + d;
    sx = (a < 0.0f || x2 > 1.0f / tth) ? x1 : x2;
    float s;
    if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
    else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
              (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    sy = s * z * sqrtf(1.0f + sx * sx);
}
inline V3 tr_sample(V3 wi, float ax, float ay, float u1, float u2) {             // :155-171
    V3 ws = normalize(V3(wi.x * ax, wi.y * ay, wi.z));
    float sx, sy;
    tr_sample11(cos_t(ws), u1, u2, sx, sy);
    float c = cos_p(ws), s = sin_p(ws);
    float tmp = c * sx - s * sy;
    sy = s * sx + c * sy;
    sx = tmp;
    sx *= ax; sy *= ay;
    return normalize(V3(-sx, -sy, 1.0f));
}
inline V3 tr_sample_wh(const Lobe& l, V3 wo, float u0, float u1) {               // :173-182
    bool flip = wo.z < 0.0f;
    V3 wh = tr_sample(flip ? -wo : wo, l.ax, l.ay, u0, u1);
    return flip ? -wh : wh;
}
inline float roughness_to_alpha(float r) {                                       // :79-84
    r = jl_max(1e-3f, r);
    float x = logf(r);
    float x3 = x * x * x;
    float x4 = pow4(x);
    return 1.62142f + 0.819955f * x + 0.1734f * (x * x) + 0.0171201f * x3 + 0.000640711f * x4;
}

// ---- BxDF evaluation f(wo, wi) in the local frame
inline RGB lobe_f(const Lobe& l, V3 wo, V3 wi) {
    switch (l.kind) {
    case L_LAMBERT: return l.r * (1.0f / PI_F);                                  // lambertian.jl:22-24
    case L_MICRO_REFL: {                                                         // microfacet.jl:221-234
        float co = fabsf(cos_t(wo)), ci = fabsf(cos_t(wi));
        V3 wh = wi + wo;
        if (ci == 0.0f || co == 0.0f) return RGB(0.0f);
        if (is_zero(wh)) return RGB(0.0f);
        wh = normalize(wh);
        float f = lobe_fresnel(l, dot(wi, face_forward(wh, V3(0, 0, 1))));
        return l.r * tr_D(l, wh) * tr_G(l, wo, wi) * f / (4.0f * ci * co);
    }
    case L_MICRO_TRANS: {                                                        // microfacet.jl:280-305
        if (same_hemisphere(wo, wi)) return RGB(0.0f);
        float co = cos_t(wo), ci = cos_t(wi);
        if (co == 0.0f || ci == 0.0f) return RGB(0.0f);
        float eta = cos_t(wo) > 0.0f ? (l.eta_b / l.eta_a) : (l.eta_a / l.eta_b);
        V3 wh = normalize(wo + wi * eta);
        if (wh.z < 0.0f) wh = -wh;
        float d_o = dot(wo, wh), d_i = dot(wi, wh);
        if (d_o * d_i > 0.0f) return RGB(0.0f);
        float f = fresnel_dielectric(d_o, l.eta_a, l.eta_b);
        float denom = d_o + eta * d_i;
        float factor = 1.0f;                                                     // Q6: `T isa Radiance` is never true
        float dd = tr_D(l, wh), dg = tr_G(l, wo, wi);
        float v = fabsf(dd * dg * d_o * d_i * (eta * eta) * (factor * factor) / (ci * co * (denom * denom)));
        return ((RGB(1.0f) - f) * l.t) * v;
    }
    case L_OREN_NAYAR: {                                                         // microfacet.jl:21-42
        float si = sin_t(wi), so = sin_t(wo);
        float max_cos = 0.0f;
        if (si > 1e-4f && so > 1e-4f) {
            float sin_pi = sin_p(wi), cos_pi = cos_p(wi), sin_po = sin_p(wo), cos_po = cos_p(wo);
            max_cos = jl_max(0.0f, cos_pi * cos_po + sin_pi * sin_po);
        }
        float sa, tb;
        bool cond = cos_t(wi) > fabsf(cos_t(wo));      // abs(Bool) == Bool (Q18)
        if (cond) { sa = so; tb = si / fabsf(cos_t(wi)); }
        else { sa = si; tb = so / fabsf(cos_t(wo)); }
        return l.r * (1.0f / PI_F) * (l.on_a + l.on_b * max_cos * sa * tb);
    }
    default: return RGB(0.0f);                                                   // specular lobes: specular.jl:23-25,69-73,133-137
    }
}

// ---- compute_pdf(bxdf, wo, wi)
inline float lobe_pdf(const Lobe& l, V3 wo, V3 wi) {
    switch (l.kind) {
    case L_FRESNEL_SPEC: return 0.0f;                                            // specular.jl:139
    case L_MICRO_REFL: {                                                         // microfacet.jl:252-258
        if (!same_hemisphere(wo, wi)) return 0.0f;
        V3 wh = normalize(wo + wi);
        return tr_pdf(l, wo, wh) / dot(4.0f * wo, wh);
    }
    case L_MICRO_TRANS: {                                                        // microfacet.jl:322-337
        if (same_hemisphere(wo, wi)) return 0.0f;
        float eta = cos_t(wo) > 0.0f ? (l.eta_b / l.eta_a) : (l.eta_a / l.eta_b);
        V3 wh = normalize(wo + wi * eta);
        float d_o = dot(wo, wh), d_i = dot(wi, wh);
        if (d_o * d_i > 0.0f) return 0.0f;
        float denom = d_o + eta * d_i;
        float dwh = fabsf(d_i * (eta * eta) / (denom * denom));
        return tr_pdf(l, wo, wh) * dwh;
    }
    default:                                                                     // bxdf.jl:23-25
        return same_hemisphere(wo, wi) ? fabsf(cos_t(wi)) * (1.0f / PI_F) : 0.0f;
    }
}

struct LobeSample { V3 wi; float pdf; RGB f; int sampled_type; /* -1 == nothing */ };

// ---- sample_f(bxdf, wo, u)
inline LobeSample lobe_sample(const Lobe& l, V3 wo, float u0, float u1) {
    LobeSample s; s.wi = V3(0.0f); s.pdf = 0.0f; s.f = RGB(0.0f); s.sampled_type = -1;
    switch (l.kind) {
    case L_SPEC_REFL: {                                                          // specular.jl:32-39
        s.wi = V3(-wo.x, -wo.y, wo.z);
        s.pdf = 1.0f;
        s.f = (lobe_fresnel(l, cos_t(s.wi)) * l.r) / fabsf(cos_t(s.wi));
        return s;
    }
    case L_SPEC_TRANS: {                                                         // specular.jl:80-104
        bool entering = cos_t(wo) > 0.0f;
        float ei = entering ? l.eta_a : l.eta_b, et = entering ? l.eta_b : l.eta_a;
        V3 wi;
        if (!refract(wo, face_forward(V3(0, 0, 1), wo), ei / et, wi)) return s;
        s.wi = wi; s.pdf = 1.0f;
        float cw = cos_t(wi);
        RGB ft = l.t * (RGB(1.0f) - fresnel_dielectric(cw, l.eta_a, l.eta_b));
        s.f = ft / fabsf(cw);                                                    // Q6: no eta^2 factor
        return s;
    }
    case L_FRESNEL_SPEC: {                                                       // specular.jl:145-173
        float fd = fresnel_dielectric(cos_t(wo), l.eta_a, l.eta_b);
        if (u0 < fd) {
            s.wi = V3(-wo.x, -wo.y, wo.z);
            s.sampled_type = BSDF_SPECULAR | BSDF_REFLECTION;
            s.pdf = fd;
            s.f = (fd * l.r) / fabsf(cos_t(s.wi));
            return s;
        }
        float ei, et;
        if (cos_t(wo) > 0.0f) { ei = l.eta_a; et = l.eta_b; } else { ei = l.eta_b; et = l.eta_a; }
        V3 wi;
        bool ok = refract(wo, face_forward(V3(0, 0, 1), wo), ei / et, wi);
        if (!ok) { s.wi = wi; s.pdf = fd; s.f = RGB(0.0f); return s; }
        float pdf = 1.0f - fd;
        RGB ft = l.t * pdf;
        s.wi = wi; s.pdf = pdf; s.f = ft / fabsf(cos_t(wi));
        s.sampled_type = BSDF_SPECULAR | BSDF_TRANSMISSION;
        return s;
    }
    case L_MICRO_REFL: {                                                         // microfacet.jl:236-250
        if (wo.z == 0.0f) return s;
        V3 wh = tr_sample_wh(l, wo, u0, u1);
        if (dot(wo, wh) < 0.0f) return s;
        V3 wi = -wo + (2.0f * dot(wo, wh)) * wh;                                 // reflect, Trace.jl:126
        if (!same_hemisphere(wo, wi)) return s;
        s.wi = wi;
        s.pdf = lobe_pdf(l, wo, wh);                                             // Q27: BxDF-level pdf with wh in the wi slot
        s.f = lobe_f(l, wo, wi);
        return s;
    }
    case L_MICRO_TRANS: {                                                        // microfacet.jl:307-320
        if (wo.z == 0.0f) return s;
        V3 wh = tr_sample_wh(l, wo, u0, u1);
        if (dot(wo, wh) < 0.0f) return s;
        float eta = cos_t(wo) > 0.0f ? (l.eta_b / l.eta_a) : (l.eta_a / l.eta_b);
        V3 wi;
        if (!refract(wo, wh, eta, wi)) return s;
        s.wi = wi; s.pdf = lobe_pdf(l, wo, wi); s.f = lobe_f(l, wo, wi);
        return s;
    }
    default: {                                                                   // bxdf.jl:33-42 (Lambertian, Oren-Nayar)
        V3 wi = cosine_sample_hemisphere(u0, u1);
        if (wo.z < 0.0f) wi = V3(wi.x, wi.y, -wi.z);
        s.wi = wi; s.pdf = lobe_pdf(l, wo, wi); s.f = lobe_f(l, wo, wi);
        return s;
    }
    }
}

// ---- BSDF container, materials/bsdf.jl
inline void bsdf_init(BSDF& b, const SurfaceInteraction& si, float eta) {        // :37-51
    b.eta = eta; b.ng = si.ng; b.ns = si.ns;
    b.ss = normalize(si.sh_dpdu);
    b.ts = cross(b.ns, b.ss);
    b.n = 0;
}
inline V3 world_to_local(const BSDF& b, V3 v) { return V3(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }
inline V3 local_to_world(const BSDF& b, V3 v) {                                  // :72-74
    return V3((b.ss.x * v.x + b.ts.x * v.y) + b.ns.x * v.z, (b.ss.y * v.x + b.ts.y * v.y) + b.ns.y * v.z,
              (b.ss.z * v.x + b.ts.z * v.y) + b.ns.z * v.z);
}
inline int num_components(const BSDF& b, uint8_t flags) {
    int n = 0;
    for (int i = 0; i < b.n; ++i) if (matches(b.lobes[i], flags)) n++;
    return n;
}
inline RGB bsdf_f(const BSDF& b, V3 wo_w, V3 wi_w, uint8_t flags = BSDF_ALL) {   // :79-100
    V3 wo = world_to_local(b, wo_w);
    if (wo.z == 0.0f) return RGB(0.0f);
    V3 wi = world_to_local(b, wi_w);
    bool reflect = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
    RGB out(0.0f);
    for (int i = 0; i < b.n; ++i) {
        const Lobe& l = b.lobes[i];
        if (matches(l, flags) && ((reflect && (l.type & BSDF_REFLECTION)) || (!reflect && (l.type & BSDF_TRANSMISSION))))
            out = out + lobe_f(l, wo, wi);
    }
    return out;
}
struct BSDFSample { V3 wi; RGB f; float pdf; uint8_t type; };
inline BSDFSample bsdf_sample_f(const BSDF& b, V3 wo_w, float u0, float u1, uint8_t type) {   // :107-175
    BSDFSample none; none.wi = V3(0.0f); none.f = RGB(0.0f); none.pdf = 0.0f; none.type = BSDF_NONE;
    int mc = num_components(b, type);
    if (mc == 0) return none;
    int64_t comp = (int64_t)ceilf(u0 * (float)mc);
    if (comp < 1) comp = 1;
    if (comp > mc) comp = mc;
    int count = (int)comp;
    comp -= 1;
    int chosen = -1;
    for (int i = 0; i < b.n; ++i) {
        if (matches(b.lobes[i], type)) {
            if (count == 1) { chosen = i; break; }
            count -= 1;
        }
    }
    const Lobe& bx = b.lobes[chosen];
    float ur0 = jl_min(u0 * (float)mc - (float)comp, 1.0f);
    V3 wo = world_to_local(b, wo_w);
    if (wo.z == 0.0f) return none;
    uint8_t sampled = bx.type;
    LobeSample ls = lobe_sample(bx, wo, ur0, u1);
    if (ls.sampled_type >= 0) sampled = (uint8_t)ls.sampled_type;
    if (ls.pdf == 0.0f) return none;
    float pdf = ls.pdf;
    RGB f = ls.f;
    V3 wi = ls.wi;
    V3 wi_w = local_to_world(b, wi);
    bool spec = (bx.type & BSDF_SPECULAR) != 0;
    if (!spec && mc > 1) {
        for (int i = 0; i < b.n; ++i)
            if (i != chosen && matches(b.lobes[i], type)) pdf += lobe_pdf(b.lobes[i], wo, wi);
    }
    if (mc > 1) pdf /= (float)mc;
    if (!spec) {
        bool reflect = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
        f = RGB(0.0f);
        for (int i = 0; i < b.n; ++i) {
            const Lobe& l = b.lobes[i];
            if (matches(l, type) && ((reflect && (l.type & BSDF_REFLECTION)) || (!reflect && (l.type & BSDF_TRANSMISSION))))
                f = f + lobe_f(l, wo, wi);
        }
    }
    BSDFSample r; r.wi = wi_w; r.f = f; r.pdf = pdf; r.type = sampled;
    return r;
}

// ---- materials/material.jl: lobes per material. multi = allow_multiple_lobes.
inline Lobe mk_lobe(int kind, uint8_t type) {
    Lobe l;
    l.kind = kind; l.type = type; l.r = RGB(0.0f); l.t = RGB(0.0f);
    l.eta_a = l.eta_b = 1.0f; l.fresnel = 0; l.fi = l.ft = 1.0f; l.ax = l.ay = 0.0f; l.on_a = l.on_b = 0.0f;
    return l;
}
inline void tr_alphas(Lobe& l, float ax, float ay) { l.ax = jl_max(1e-3f, ax); l.ay = jl_max(1e-3f, ay); }   // microfacet.jl:58-62
inline void material_lobes(const trace_material& m, BSDF& b, bool multi) {
    RGB A = clamp0(RGB(m.a[0], m.a[1], m.a[2])), Bc = clamp0(RGB(m.b[0], m.b[1], m.b[2]));
    switch (m.kind) {
    case TRACE_MAT_MATTE: {                                                      // :16-31
        if (is_black(A)) return;
        float sigma = jl_clamp(m.rough_u, 0.0f, 90.0f);
        if (sigma == 0.0f) { Lobe l = mk_lobe(L_LAMBERT, BSDF_DIFFUSE | BSDF_REFLECTION); l.r = A; b.lobes[b.n++] = l; }
        else {
            Lobe l = mk_lobe(L_OREN_NAYAR, BSDF_DIFFUSE | BSDF_REFLECTION); l.r = A;
            float s = sigma * (PI_F / 180.0f);                                   // deg2rad
            float s2 = s * s;
            l.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
            l.on_b = 0.45f * s2 / (s2 + 0.09f);
            b.lobes[b.n++] = l;
        }
        return;
    }
    case TRACE_MAT_MIRROR: {                                                     // :39-46
        if (is_black(A)) return;
        Lobe l = mk_lobe(L_SPEC_REFL, BSDF_SPECULAR | BSDF_REFLECTION); l.r = A; l.fresnel = 0; b.lobes[b.n++] = l;
        return;
    }
    case TRACE_MAT_GLASS: {                                                      // :75-116
        float eta = m.eta, ur = m.rough_u, vr = m.rough_v;
        b.eta = eta;
        if (is_black(A) && is_black(Bc)) return;
        bool is_spec = ur == 0.0f && vr == 0.0f;
        if (is_spec && multi) {
            Lobe l = mk_lobe(L_FRESNEL_SPEC, BSDF_SPECULAR | BSDF_TRANSMISSION | BSDF_REFLECTION);
            l.r = A; l.t = Bc; l.eta_a = 1.0f; l.eta_b = eta; b.lobes[b.n++] = l;
            return;
        }
        if (m.remap) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
        if (!is_black(A)) {
            Lobe l = mk_lobe(is_spec ? L_SPEC_REFL : L_MICRO_REFL,
                             is_spec ? (BSDF_SPECULAR | BSDF_REFLECTION) : (BSDF_REFLECTION | BSDF_GLOSSY));
            l.r = A; l.fresnel = 1; l.fi = 1.0f; l.ft = eta;
            if (!is_spec) tr_alphas(l, ur, vr);
            b.lobes[b.n++] = l;
        }
        if (!is_black(Bc)) {
            Lobe l = mk_lobe(is_spec ? L_SPEC_TRANS : L_MICRO_TRANS,
                             is_spec ? (BSDF_SPECULAR | BSDF_TRANSMISSION) : (BSDF_TRANSMISSION | BSDF_GLOSSY));
            l.t = Bc; l.eta_a = 1.0f; l.eta_b = eta;
            if (!is_spec) tr_alphas(l, ur, vr);
            b.lobes[b.n++] = l;
        }
        return;
    }
    case TRACE_MAT_PLASTIC: {                                                    // :135-151
        if (!is_black(A)) { Lobe l = mk_lobe(L_LAMBERT, BSDF_DIFFUSE | BSDF_REFLECTION); l.r = A; b.lobes[b.n++] = l; }
        if (is_black(Bc)) return;
        float rough = m.rough_u;
        if (m.remap) rough = roughness_to_alpha(rough);
        Lobe l = mk_lobe(L_MICRO_REFL, BSDF_REFLECTION | BSDF_GLOSSY);
        l.r = Bc; l.fresnel = 1; l.fi = 1.5f; l.ft = 1.0f;
        tr_alphas(l, rough, rough);
        b.lobes[b.n++] = l;
        return;
    }
    }
}
inline void compute_scattering(const Scene& s, const SurfaceInteraction& si, bool multi, BSDF& b) {   // primitive.jl:29-35
    bsdf_init(b, si, 1.0f);
    material_lobes(s.materials[si.material], b, multi);
}

// ---- lights
inline float spot_falloff(const trace_light& l, V3 w) {                          // spot.jl:32-40
    V3 wl = normalize(xf_vector(l.inv_m, w));
    float c = wl.z;
    if (c < l.cos_total_width) return 0.0f;
    if (c >= l.cos_falloff_start) return 1.0f;
    float d = (c - l.cos_total_width) / (l.cos_falloff_start - l.cos_total_width);
    return pow4(d);
}
inline void sample_li(const trace_light& l, V3 p, RGB& radiance, V3& wi, float& pdf, V3& light_pos) {   // point.jl:50-58, spot.jl:22-30
    if (l.kind == TRACE_LIGHT_DIRECTIONAL) {       // directional.jl:39-47: position = direction, cos_total_width = world_radius
        wi = V3(l.position[0], l.position[1], l.position[2]);
        pdf = 1.0f;
        light_pos = p + wi * (2.0f * l.cos_total_width);
        radiance = RGB(l.I[0], l.I[1], l.I[2]);
        return;
    }
    V3 pos(l.position[0], l.position[1], l.position[2]);
    wi = normalize(pos - p);
    pdf = 1.0f;
    light_pos = pos;
    V3 dd = pos - p;
    float d2 = dot(dd, dd);
    RGB I(l.I[0], l.I[1], l.I[2]);
    if (l.kind == TRACE_LIGHT_POINT) radiance = I / d2;
    else radiance = (I * spot_falloff(l, -wi)) / d2;
}
inline RGB light_power(const trace_light& l) {                                   // point.jl:74-76, spot.jl:42-44
    RGB I(l.I[0], l.I[1], l.I[2]);
    if (l.kind == TRACE_LIGHT_POINT) return (4.0f * PI_F) * I;
    if (l.kind == TRACE_LIGHT_DIRECTIONAL) return (I * PI_F) * (l.cos_total_width * l.cos_total_width);   // directional.jl:54-56
    return ((I * 2.0f) * PI_F) * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width));
}
struct LeSample { RGB le; Ray ray; V3 n; float pdf_pos, pdf_dir; };
inline LeSample sample_le(const trace_light& l, float u0, float u1) {            // point.jl:60-69, spot.jl:46-55
    LeSample s;
    V3 pos(l.position[0], l.position[1], l.position[2]);
    RGB I(l.I[0], l.I[1], l.I[2]);
    s.ray.o = pos; s.ray.t_max = INF_F; s.ray.time = 0.0f;
    s.pdf_pos = 1.0f;
    if (l.kind == TRACE_LIGHT_POINT) {
        s.ray.d = uniform_sample_sphere(u0, u1);
        s.n = s.ray.d;
        s.pdf_dir = uniform_sphere_pdf();
        s.le = I;
    } else {
        s.ray.d = xf_vector(l.m, uniform_sample_cone(u0, u1, l.cos_total_width));
        s.n = s.ray.d;
        s.pdf_dir = uniform_cone_pdf(l.cos_total_width);
        s.le = I * spot_falloff(l, s.ray.d);
    }
    return s;
}
// unoccluded(VisibilityTester(p0, p1)) : spawn_ray(p0, p1) with un-normalised d, t_max = Inf (Q5), Trace.jl:196-202
inline Ray shadow_ray(V3 p0, V3 p1) {
    Ray r;
    V3 d = p1 - p0;
    r.o = p0 + 1e-6f * d;
    r.d = d; r.t_max = INF_F; r.time = 0.0f;
    return r;
}
inline Ray spawn_ray_dir(V3 p, V3 dir) {                                         // Trace.jl:206-211
    Ray r; r.o = p + 1e-6f * dir; r.d = dir; r.t_max = INF_F; r.time = 0.0f;
    return r;
}

}  // namespace ref
