// ref_bvh.cpp — CPU ORACLE (test infrastructure only): literal restatement of the BVHAccel constructor.
// Follows src/accel/bvh.jl:55-206 (BVHAccel, _init, _unroll), src/Trace.jl:128-137 (partition!),
// src/bounds.jl:82-85,112-120,134-143. Quirks reproduced: SAH buckets start as the
// point (0,0,0); the cost uses bucket *counts of buckets*, not primitive counts; partition! never tests
// the first element and the split is [from,mid] | [mid+1,to]; zero-primitive leaves can be produced.
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "ref_internal.hpp"

namespace ref {

struct PrimInfo {           // BVHPrimitiveInfo, bvh.jl:3-14
    uint32_t number;
    B3 bounds;
    V3 centroid;
};

struct BuildNode {          // BVHNode, bvh.jl:16-35
    B3 bounds;
    int child[2];
    uint8_t axis;
    uint32_t offset, n;
};

struct Builder {
    std::vector<PrimInfo> info;
    std::vector<BuildNode> nodes;
    std::vector<uint32_t> ordered;
    int max_node_primitives;
    int max_depth = 0;

    int create_leaf(int64_t from, int64_t to, const B3& bounds) {       // bvh.jl:98-107
        BuildNode nd;
        nd.bounds = bounds; nd.child[0] = nd.child[1] = -1; nd.axis = 0;
        nd.offset = (uint32_t)ordered.size();
        nd.n = (uint32_t)(to - from + 1);
        for (int64_t i = from; i <= to; ++i) ordered.push_back(info[i].number);
        nodes.push_back(nd);
        return (int)nodes.size() - 1;
    }

    static int bucket_of(const B3& cb, V3 c, int dim) {                  // bvh.jl:133-136 (0-based result)
        float b = floorf(12.0f * offset(cb, c)[dim]);
        int bi = (int)b;
        if (bi == 12) bi = 11;
        return bi;
    }

    int init(int64_t from, int64_t to, int depth) {                      // bvh.jl:87-185, inclusive range, 0-based
        if (depth > 20000) throw std::runtime_error("BVH recursion too deep");
        if (depth > max_depth) max_depth = depth;
        int64_t n = to - from + 1;
        B3 bounds;
        for (int64_t i = from; i <= to; ++i) bounds = unite(bounds, info[i].bounds);
        if (n == 1) return create_leaf(from, to, bounds);
        B3 cb;
        for (int64_t i = from; i <= to; ++i) cb = unite(cb, B3(info[i].centroid));
        int dim = maximum_extent(cb);
        if (!is_valid(cb) || cb.lo[dim] == cb.hi[dim]) return create_leaf(from, to, bounds);
        int64_t mid;
        if (n <= 2) {
            // partialsort!(view, 1, by = centroid[dim]) on two elements: insertion sort, strict '<'
            mid = (from + to) / 2;
            if (info[to].centroid[dim] < info[from].centroid[dim]) std::swap(info[from], info[to]);
        } else {
            const int NB = 12;
            uint32_t count[NB];
            B3 bb[NB];
            for (int b = 0; b < NB; ++b) { count[b] = 0; bb[b] = B3(V3(0.0f)); }   // Bounds3(Point3f(0f0))
            for (int64_t i = from; i <= to; ++i) {
                int b = bucket_of(cb, info[i].centroid, dim);
                count[b] += 1;
                bb[b] = unite(bb[b], info[i].bounds);
            }
            float costs[NB - 1];
            float sa_all = surface_area(bounds);
            for (int i = 1; i <= NB - 1; ++i) {          // 1-based i as in the reference
                // it1 = 1:i, it2 = (i+1):(NB-1)
                B3 u1 = bb[0];
                for (int b = 2; b <= i; ++b) u1 = unite(u1, bb[b - 1]);
                float s1 = (float)i * surface_area(u1);
                float s2 = 0.0f;
                int len2 = (NB - 1) - (i + 1) + 1;
                if (len2 > 0) {
                    B3 u2 = bb[i];
                    for (int b = i + 2; b <= NB - 1; ++b) u2 = unite(u2, bb[b - 1]);
                    s2 = (float)len2 * surface_area(u2);
                }
                costs[i - 1] = 1.0f + (s1 + s2) / sa_all;
            }
            // argmin(costs): first minimum; a NaN wins over everything (Julia findmin semantics)
            int min_id = 0;
            for (int i = 0; i < NB - 1; ++i) {
                if (std::isnan(costs[i])) { min_id = i; break; }
                if (costs[i] < costs[min_id]) min_id = i;
            }
            if (!(n > max_node_primitives || costs[min_id] < (float)n)) return create_leaf(from, to, bounds);
            // partition!(primitives_info, from:to, pred)  — Trace.jl:128-137
            int64_t left = from;
            for (int64_t i = from; i <= to; ++i) {
                if (left != i && bucket_of(cb, info[i].centroid, dim) <= min_id) {
                    std::swap(info[i], info[left]);
                    left += 1;
                }
            }
            mid = left;
        }
        int slot = (int)nodes.size();
        nodes.push_back(BuildNode());
        int l = init(from, mid, depth + 1);
        int r = init(mid + 1, to, depth + 1);
        BuildNode nd;
        nd.bounds = unite(nodes[l].bounds, nodes[r].bounds);
        nd.child[0] = l; nd.child[1] = r; nd.axis = (uint8_t)dim; nd.offset = 0; nd.n = 0;
        nodes[slot] = nd;
        return slot;
    }
};

}  // namespace ref

using namespace ref;

struct ref_bvh {
    std::vector<trace_bvh_node> nodes;
    std::vector<uint32_t> order;
    int max_depth;
};

static void unroll(const std::vector<BuildNode>& bn, int idx, std::vector<trace_bvh_node>& out) {   // bvh.jl:187-206
    // explicit stack preorder flatten (first child = parent + 1)
    struct Item { int node; int parent_slot; };
    std::vector<Item> st;
    st.push_back({idx, -1});
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        const BuildNode& b = bn[it.node];
        int slot = (int)out.size();
        if (it.parent_slot >= 0) out[it.parent_slot].offset = (uint32_t)slot;   // we are a second child
        trace_bvh_node ln;
        ln.bmin[0] = b.bounds.lo.x; ln.bmin[1] = b.bounds.lo.y; ln.bmin[2] = b.bounds.lo.z;
        ln.bmax[0] = b.bounds.hi.x; ln.bmax[1] = b.bounds.hi.y; ln.bmax[2] = b.bounds.hi.z;
        if (b.child[0] < 0) {
            ln.offset = b.offset; ln.meta = TRACE_NODE_LEAF | b.n;
            out.push_back(ln);
        } else {
            ln.offset = 0; ln.meta = (uint32_t)b.axis << 30;
            out.push_back(ln);
            st.push_back({b.child[1], slot});
            st.push_back({b.child[0], -1});
        }
    }
}

extern "C" int ref_bvh_build(const float* pb, int64_t n, int max_node_primitives, ref_bvh** out) {
    if (!out) return 1;
    *out = nullptr;
    try {
        Builder B;
        B.max_node_primitives = std::min(255, max_node_primitives);      // bvh.jl:58
        B.info.resize(n);
        for (int64_t i = 0; i < n; ++i) {
            PrimInfo& p = B.info[i];
            p.number = (uint32_t)i;
            p.bounds = B3(V3(pb[6 * i], pb[6 * i + 1], pb[6 * i + 2]), V3(pb[6 * i + 3], pb[6 * i + 4], pb[6 * i + 5]));
            p.centroid = 0.5f * p.bounds.lo + 0.5f * p.bounds.hi;         // bvh.jl:11
        }
        ref_bvh* r = new ref_bvh();
        r->max_depth = 0;
        if (n > 0) {
            int root = B.init(0, n - 1, 1);
            r->nodes.reserve(B.nodes.size());
            unroll(B.nodes, root, r->nodes);
            r->order.swap(B.ordered);
            r->max_depth = B.max_depth;
        }
        *out = r;
        return 0;
    } catch (...) {
        return 2;
    }
}
extern "C" int64_t ref_bvh_num_nodes(const ref_bvh* b) { return (int64_t)b->nodes.size(); }
extern "C" int ref_bvh_max_depth(const ref_bvh* b) { return b->max_depth; }
extern "C" int ref_bvh_copy(const ref_bvh* b, trace_bvh_node* nodes_out, uint32_t* order_out) {
    if (nodes_out && !b->nodes.empty()) memcpy(nodes_out, b->nodes.data(), b->nodes.size() * sizeof(trace_bvh_node));
    if (order_out && !b->order.empty()) memcpy(order_out, b->order.data(), b->order.size() * sizeof(uint32_t));
    return 0;
}
extern "C" void ref_bvh_free(ref_bvh* b) { delete b; }
