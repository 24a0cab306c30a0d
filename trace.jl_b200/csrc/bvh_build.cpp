// bvh_build.cpp — host-side BVH build of libtrace_cuda.so (trace_bvh_build, include/trace_cuda.h).
//
// north_star: "The BVH is built on the host with the reference's SAH split logic, then flattened into a
// 32-byte-aligned LinearBVH node array".  The split logic is the reference's, quirks included
// (src/accel/bvh.jl:87-185, src/Trace.jl:128-137): 12 buckets that start as the point (0,0,0),
// a cost that weighs each side by its NUMBER OF BUCKETS, a partition that never tests the first element and splits
// [from..mid] | [mid+1..to], two-primitive nodes split at the smaller centroid, zero-primitive leaves kept.
// Unlike the reference's recursive, allocation-per-node build this one is iterative: SoA bounds/centroids, one index
// permutation that is partitioned in place, and nodes emitted directly in the flattened preorder (first child =
// parent + 1), so building the 10 M-triangle scene needs no recursion and no per-node heap traffic.
// Compiled with -ffp-contract=off: the cost arithmetic must round like the reference's.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/trace_cuda.h"

namespace {

const float kInf = std::numeric_limits<float>::infinity();

// Julia's min/max on floats: NaN propagates, -0.0 orders below +0.0.
inline float fmin_jl(float a, float b) {
    if (a != a || b != b) return a + b;
    if (b < a) return b;
    if (a == b && std::signbit(b)) return b;
    return a;
}
inline float fmax_jl(float a, float b) {
    if (a != a || b != b) return a + b;
    if (b > a) return b;
    if (a == b && !std::signbit(b)) return b;
    return a;
}

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = kInf; hi[k] = -kInf; } }
    void point(float v) { for (int k = 0; k < 3; ++k) { lo[k] = v; hi[k] = v; } }
    void grow(const float* l, const float* h) {
        for (int k = 0; k < 3; ++k) { lo[k] = fmin_jl(lo[k], l[k]); hi[k] = fmax_jl(hi[k], h[k]); }
    }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * ((dx * dy + dx * dz) + dy * dz);
    }
    bool valid() const {
        for (int k = 0; k < 3; ++k) if (lo[k] == kInf || hi[k] == -kInf) return false;
        return true;
    }
    int widest() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx > dy && dx > dz) return 0;
        if (dy > dz) return 1;
        return 2;
    }
};

struct Task { int64_t from, to; int64_t patch; };   // inclusive range into the permutation; node whose `offset` we are

}  // namespace

struct trace_bvh {
    std::vector<trace_bvh_node> nodes;
    std::vector<uint32_t> order;
};

extern "C" int trace_bvh_build(const float* pb, int64_t n, int max_node_primitives, trace_bvh** out) {
    if (!out || n < 0 || (n > 0 && !pb)) return 1;
    *out = nullptr;
    trace_bvh* bvh = new (std::nothrow) trace_bvh();
    if (!bvh) return 2;
    if (n == 0) { *out = bvh; return 0; }
    const int max_prims = max_node_primitives < 255 ? max_node_primitives : 255;
    const int NB = 12;
    std::vector<float> cen((size_t)n * 3);
    std::vector<uint32_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        perm[i] = (uint32_t)i;
        for (int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * pb[6 * i + k] + 0.5f * pb[6 * i + 3 + k];
    }
    try {
        bvh->nodes.reserve((size_t)(2 * n + 16));
        bvh->order.reserve((size_t)n);
    } catch (...) { delete bvh; return 2; }

    std::vector<Task> todo;
    todo.push_back({0, n - 1, -1});
    const int64_t node_limit = 8 * n + 1024;     // a guard, never reached by terminating inputs
    while (!todo.empty()) {
        const Task t = todo.back();
        todo.pop_back();
        const int64_t slot = (int64_t)bvh->nodes.size();
        if (slot > node_limit) { delete bvh; return 3; }
        if (t.patch >= 0) bvh->nodes[t.patch].offset = (uint32_t)slot;
        const int64_t count = t.to - t.from + 1;

        Box all; all.reset();
        for (int64_t i = t.from; i <= t.to; ++i) { const float* b = pb + 6 * (size_t)perm[i]; all.grow(b, b + 3); }
        trace_bvh_node node;
        for (int k = 0; k < 3; ++k) { node.bmin[k] = all.lo[k]; node.bmax[k] = all.hi[k]; }

        bool leaf = (count == 1);
        int axis = 0;
        Box cb; cb.reset();
        if (!leaf) {
            for (int64_t i = t.from; i <= t.to; ++i) { const float* c = &cen[3 * (size_t)perm[i]]; cb.grow(c, c); }
            axis = cb.widest();
            if (!cb.valid() || cb.lo[axis] == cb.hi[axis]) leaf = true;
        }
        int64_t mid = t.from;
        if (!leaf) {
            // relative position of a centroid along `axis`, bounds.jl:134-143 (offset)
            const bool any_extent = cb.hi[0] > cb.lo[0] || cb.hi[1] > cb.lo[1] || cb.hi[2] > cb.lo[2];
            const float extent = cb.hi[axis] > cb.lo[axis] ? cb.hi[axis] - cb.lo[axis] : 1.0f;
            auto bucket = [&](uint32_t prim) -> int {
                float o = cen[3 * (size_t)prim + axis] - cb.lo[axis];
                if (any_extent) o = o / extent;
                int b = (int)std::floor(12.0f * o);
                return b == NB ? NB - 1 : b;
            };
            if (count <= 2) {
                // partialsort!(view, 1, by = centroid[axis]) on two entries; mid = (from + to) ÷ 2 = from
                if (cen[3 * (size_t)perm[t.to] + axis] < cen[3 * (size_t)perm[t.from] + axis]) std::swap(perm[t.from], perm[t.to]);
                mid = (t.from + t.to) / 2;
            } else {
                Box bk[NB];
                for (int b = 0; b < NB; ++b) bk[b].point(0.0f);
                for (int64_t i = t.from; i <= t.to; ++i) {
                    const uint32_t p = perm[i];
                    const float* b = pb + 6 * (size_t)p;
                    bk[bucket(p)].grow(b, b + 3);
                }
                // prefix unions 0..i and suffix unions i..10 (bucket 11 never enters the right-hand side)
                Box pre[NB], suf[NB];
                pre[0] = bk[0];
                for (int b = 1; b < NB; ++b) { pre[b] = pre[b - 1]; pre[b].grow(bk[b]); }
                // the reference folds the right side left-to-right starting at bucket i+1; min/max are exact, so a
                // suffix scan yields the same box
                suf[NB - 2] = bk[NB - 2];
                for (int b = NB - 3; b >= 0; --b) { suf[b] = bk[b]; suf[b].grow(suf[b + 1]); }
                const float total_area = all.area();
                int best = 0;
                float best_cost = 0.0f;
                bool best_nan = false;
                for (int i = 0; i < NB - 1; ++i) {            // split after bucket i (0-based)
                    float left = (float)(i + 1) * pre[i].area();
                    float right = 0.0f;
                    const int n_right = (NB - 1) - (i + 1);
                    if (n_right > 0) right = (float)n_right * suf[i + 1].area();
                    const float cost = 1.0f + (left + right) / total_area;
                    if (i == 0) { best = 0; best_cost = cost; best_nan = cost != cost; }
                    else if (!best_nan && (cost != cost || cost < best_cost)) { best = i; best_cost = cost; best_nan = cost != cost; }
                }
                if (!(count > max_prims || (double)best_cost < (double)count)) leaf = true;
                else {
                    // partition!, Trace.jl:128-137
                    int64_t left = t.from;
                    for (int64_t i = t.from; i <= t.to; ++i) {
                        if (left != i && bucket(perm[i]) <= best) { std::swap(perm[i], perm[left]); ++left; }
                    }
                    mid = left;
                }
            }
        }
        if (leaf) {
            node.offset = (uint32_t)bvh->order.size();
            node.meta = TRACE_NODE_LEAF | (uint32_t)(count < 0 ? 0 : count);
            for (int64_t i = t.from; i <= t.to; ++i) bvh->order.push_back(perm[i]);
            bvh->nodes.push_back(node);
        } else {
            node.offset = 0;
            node.meta = (uint32_t)axis << 30;
            bvh->nodes.push_back(node);
            todo.push_back({mid + 1, t.to, slot});     // second child: patched into `offset` when it is emitted
            todo.push_back({t.from, mid, -1});          // first child: emitted next, at slot + 1
        }
    }
    *out = bvh;
    return 0;
}

// Opt-in alternative: a conventional binned SAH over the same inputs, emitted in the same node format,
// so every kernel runs on it unchanged.  Differences from the literal build above: buckets start EMPTY, each side is
// weighted by its primitive COUNT (cost = 1/8 + (nL*aL + nR*aR)/A), a node becomes a leaf when that is cheaper and it
// holds <= max_node_primitives, the partition tests every element, and a degenerate split falls back to the median.
// The closest hit of a ray does not depend on the tree (ties between equal t aside), so images agree; traversal work
// drops because the literal cost function is degenerate (Q16: depth 42 and 2 663 empty leaves on the caustic mesh).
extern "C" int trace_bvh_build_sah(const float* pb, int64_t n, int max_node_primitives, trace_bvh** out) {
    if (!out || n < 0 || (n > 0 && !pb)) return 1;
    *out = nullptr;
    trace_bvh* bvh = new (std::nothrow) trace_bvh();
    if (!bvh) return 2;
    if (n == 0) { *out = bvh; return 0; }
    const int max_prims = max_node_primitives < 1 ? 1 : (max_node_primitives < 255 ? max_node_primitives : 255);
    const int NB = 16;
    std::vector<float> cen((size_t)n * 3);
    std::vector<uint32_t> perm((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        perm[i] = (uint32_t)i;
        for (int k = 0; k < 3; ++k) cen[3 * i + k] = 0.5f * pb[6 * i + k] + 0.5f * pb[6 * i + 3 + k];
    }
    try {
        bvh->nodes.reserve((size_t)(2 * n + 16));
        bvh->order.reserve((size_t)n);
    } catch (...) { delete bvh; return 2; }
    std::vector<Task> todo;
    todo.push_back({0, n - 1, -1});
    while (!todo.empty()) {
        const Task t = todo.back();
        todo.pop_back();
        const int64_t slot = (int64_t)bvh->nodes.size();
        if (t.patch >= 0) bvh->nodes[t.patch].offset = (uint32_t)slot;
        const int64_t count = t.to - t.from + 1;
        Box all; all.reset();
        Box cb; cb.reset();
        for (int64_t i = t.from; i <= t.to; ++i) {
            const float* b = pb + 6 * (size_t)perm[i];
            all.grow(b, b + 3);
            const float* c = &cen[3 * (size_t)perm[i]];
            cb.grow(c, c);
        }
        trace_bvh_node node;
        for (int k = 0; k < 3; ++k) { node.bmin[k] = all.lo[k]; node.bmax[k] = all.hi[k]; }
        const int axis = cb.widest();
        bool leaf = count == 1 || !cb.valid() || !(cb.hi[axis] > cb.lo[axis]);
        if (leaf && count > max_prims && count > 1) leaf = false;        // identical centroids: split at the median
        int64_t mid = (t.from + t.to) / 2;                                 // last index of the left side
        if (!leaf) {
            bool split_done = false;
            if (cb.valid() && cb.hi[axis] > cb.lo[axis] && count >= 2) {
                const float scale = (float)NB / (cb.hi[axis] - cb.lo[axis]);
                auto bucket = [&](uint32_t prim) -> int {
                    int b = (int)((cen[3 * (size_t)prim + axis] - cb.lo[axis]) * scale);
                    return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                };
                Box bk[NB];
                int64_t cnt[NB];
                for (int b = 0; b < NB; ++b) { bk[b].reset(); cnt[b] = 0; }
                for (int64_t i = t.from; i <= t.to; ++i) {
                    const uint32_t p = perm[i];
                    const float* b = pb + 6 * (size_t)p;
                    const int k = bucket(p);
                    bk[k].grow(b, b + 3); cnt[k]++;
                }
                float right_area[NB];
                int64_t right_cnt[NB];
                Box acc; acc.reset();
                int64_t c = 0;
                for (int b = NB - 1; b >= 1; --b) {
                    if (cnt[b]) acc.grow(bk[b]);
                    c += cnt[b];
                    right_area[b] = c ? acc.area() : 0.0f; right_cnt[b] = c;
                }
                acc.reset(); c = 0;
                const float total_area = all.area();
                float best_cost = kInf;
                int best = -1;
                for (int b = 0; b < NB - 1; ++b) {                          // split after bucket b
                    if (cnt[b]) acc.grow(bk[b]);
                    c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    const float cost = 0.125f + ((float)c * acc.area() + (float)right_cnt[b + 1] * right_area[b + 1]) /
                                                    (total_area > 0.0f ? total_area : 1.0f);
                    if (cost < best_cost) { best_cost = cost; best = b; }
                }
                if (best >= 0 && (count > max_prims || best_cost < (float)count)) {
                    int64_t left = t.from;
                    for (int64_t i = t.from; i <= t.to; ++i)
                        if (bucket(perm[i]) <= best) { std::swap(perm[i], perm[left]); ++left; }
                    mid = left - 1;
                    split_done = mid >= t.from && mid < t.to;
                } else if (best >= 0 || count <= max_prims) {
                    leaf = count <= max_prims;
                }
            }
            if (!leaf && !split_done) {
                // median split along the axis (also: two primitives, or all centroids in one bucket)
                mid = (t.from + t.to) / 2;
                std::nth_element(perm.begin() + t.from, perm.begin() + mid, perm.begin() + t.to + 1, [&](uint32_t a, uint32_t b) {
                    return cen[3 * (size_t)a + axis] < cen[3 * (size_t)b + axis];
                });
            }
        }
        if (leaf) {
            node.offset = (uint32_t)bvh->order.size();
            node.meta = TRACE_NODE_LEAF | (uint32_t)count;
            for (int64_t i = t.from; i <= t.to; ++i) bvh->order.push_back(perm[i]);
            bvh->nodes.push_back(node);
        } else {
            node.offset = 0;
            node.meta = (uint32_t)axis << 30;
            bvh->nodes.push_back(node);
            todo.push_back({mid + 1, t.to, slot});
            todo.push_back({t.from, mid, -1});
        }
    }
    *out = bvh;
    return 0;
}

extern "C" int64_t trace_bvh_num_nodes(const trace_bvh* b) { return b ? (int64_t)b->nodes.size() : 0; }
extern "C" int64_t trace_bvh_num_prims(const trace_bvh* b) { return b ? (int64_t)b->order.size() : 0; }
extern "C" int trace_bvh_copy(const trace_bvh* b, trace_bvh_node* nodes_out, uint32_t* order_out) {
    if (!b) return 1;
    if (nodes_out && !b->nodes.empty()) memcpy(nodes_out, b->nodes.data(), b->nodes.size() * sizeof(trace_bvh_node));
    if (order_out && !b->order.empty()) memcpy(order_out, b->order.data(), b->order.size() * sizeof(uint32_t));
    return 0;
}
extern "C" void trace_bvh_free(trace_bvh* b) { delete b; }
