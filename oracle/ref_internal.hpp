// ref_internal.hpp — shared declarations of the CPU ORACLE (test infrastructure only).
#pragma once
#include <vector>

#include "ref_math.hpp"
#include "trace_ref.h"

namespace ref {

struct Ray {                 // src/ray.jl:1-6
    V3 o, d;
    float t_max;
    float time;
};

struct Tri {
    V3 p[3];
    V3 n[3];
    bool has_normals;
    bool flip;
};

struct Scene {
    std::vector<trace_bvh_node> nodes;
    std::vector<trace_prim> prims;
    std::vector<Tri> tris;
    std::vector<trace_sphere> spheres;
    std::vector<trace_material> materials;
    std::vector<trace_light> lights;
};

struct Hit {                 // what the winning candidate's SurfaceInteraction is rebuilt from
    bool hit = false;
    int32_t prim = -1;       // index into Scene::prims (BVH order)
    float t = 0.0f;
    float b[3] = {0, 0, 0};  // triangle barycentrics
};

struct Counters { uint64_t nodes = 0, prims = 0, max_stack = 0; };

struct SurfaceInteraction {  // the fields of src/surface_interaction.jl:1-49 the hot path reads
    V3 p, wo, ng, ns;        // core.p, core.wo, core.n, shading.n
    V3 sh_dpdu;              // shading.∂p∂u (BSDF tangent source, materials/bsdf.jl:41-45)
    float u, v;
    int32_t prim;
    uint32_t material;
};

// closest hit / any hit (src/accel/bvh.jl:212-299)
bool intersect_closest(const Scene& s, Ray& ray, Hit& hit, int slab, Counters* c);
bool intersect_any(const Scene& s, Ray& ray, int slab, Counters* c);
void build_interaction(const Scene& s, const Ray& ray_at_hit, const Hit& hit, SurfaceInteraction& si);
bool slab_test(const float* bmin, const float* bmax, const Ray& r, V3 inv_dir, const int neg[3], int slab);

}  // namespace ref

struct ref_scene { ref::Scene s; };
