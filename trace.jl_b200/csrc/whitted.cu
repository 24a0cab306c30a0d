// whitted.cu — wavefront Whitted integrator: generate -> extend (closest hit) -> shade -> shadow (any hit) ->
// accumulate, with SoA float4 ray / hit / shadow queues in HBM and warp-aggregated atomic compaction between stages.
//
// Replaces (i::SamplerIntegrator)(scene), li, specular_reflect, specular_transmit (src/integrators/sampler.jl:12-199)
// and add_sample! / merge_film_tile! (src/film.jl:134-193).
//
// The reference's `li` is a depth-first recursion that returns L bottom-up; here every ray carries the product of
// the f·|cos|/pdf factors above it (its weight), every shadow ray carries weight·f·Li·|cos|/pdf, and an unoccluded
// shadow ray adds that to the per-sample accumulator.  A sample is splatted into the film only when its whole ray
// tree has drained, which keeps the reference's per-sample `isnan(l) -> 0` rule (sampler.jl:46).
//
// Slot layout of one batch: slot = (local tile, pixel in tile, sample); 32 consecutive slots are 2 pixels x 16 spp or
// (spp = 1) two 16-pixel rows of one 16x16 tile, so primary rays of a warp are coherent.
#include "context.hpp"
#include "shading.cuh"
#include "wavefront.cuh"

// what may change between two renders that replay the same CUDA graph: read through a pointer, not baked into a node
struct WhittedFrame {
    DeviceCamera cam;
    uint64_t seed;
};

struct WhittedLaunch {
    DeviceScene sc;
    const WhittedFrame* frame;
    DeviceFilm film;
    int spp, max_depth;
    long long slot_begin;      // first slot of this batch (global over the rank's tile list)
    int n_slots;
    const int* tiles;          // tile indices owned by this rank
    float4 *ro[2], *rd[2], *rw[2];   // ray queues: {o, tmax} {d, slot} {weight, -}
    float4* hits;              // {b2, prim+1 (bits), b0, b1}
    float4 *so, *sd, *sc_contrib;     // shadow queue: {o,-} {d, slot} {contribution,-}
    float4* accum;             // per slot: radiance
    float2* filmpos;           // per slot: p_film (x < -1e29: inactive slot)
    int* counters;             // [level] rays in queue at that level (1-based), [32] shadow rays of the whole batch
    int cap_rays, cap_shadow;
    float4* film_rgbw;         // per film pixel: sum(L*w) rgb, sum(w)
};

__device__ __forceinline__ void slot_to_pixel(const WhittedLaunch& L, long long slot, int& px, int& py, int& s, int& tile) {
    const int per_tile = 256 * L.spp;
    const long long lt = slot / per_tile;
    const int within = (int)(slot - lt * per_tile);
    const int pix = within / L.spp;
    s = within - pix * L.spp;
    tile = L.tiles[lt];
    const int tx = tile % L.film.tiles_x, ty = tile / L.film.tiles_x;
    px = L.film.sb_x0 + tx * 16 + (pix & 15);
    py = L.film.sb_y0 + ty * 16 + (pix >> 4);
}

__global__ void __launch_bounds__(256) k_wh_generate(WhittedLaunch L) {
    const DeviceCamera cam = L.frame->cam;
    const uint64_t seed = L.frame->seed;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n_slots; i += gridDim.x * blockDim.x) {
        int px, py, s, tile;
        slot_to_pixel(L, L.slot_begin + i, px, py, s, tile);
        L.accum[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (px > L.film.sb_x1 || py > L.film.sb_y1) { L.filmpos[i] = make_float2(-1e30f, -1e30f); continue; }
        const uint32_t pix = (uint32_t)((py - L.film.sb_y0) * (L.film.sb_x1 - L.film.sb_x0 + 1) + (px - L.film.sb_x0));
        const float u0 = rng_uniform(seed, pix, (uint32_t)s, 0), u1 = rng_uniform(seed, pix, (uint32_t)s, 1);
        const float fx = (float)px + u0, fy = (float)py + u1;
        float l0 = 0.0f, l1 = 0.0f;
        if (cam.lens_radius > 0.0f) { l0 = rng_uniform(seed, pix, (uint32_t)s, 2); l1 = rng_uniform(seed, pix, (uint32_t)s, 3); }
        float3 o, d;
        generate_camera_ray(cam, fx, fy, l0, l1, o, d);
        L.filmpos[i] = make_float2(fx, fy);
        const int q = queue_claim(&L.counters[1]);
        L.ro[0][q] = f4(o, TR_INF);
        L.rd[0][q] = f4(d, __int_as_float(i));
        L.rw[0][q] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(128) k_wh_shade(WhittedLaunch L, int level) {
    const int cur = (level - 1) & 1, nxt = level & 1;
    const int n = min(L.counters[level], L.cap_rays);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 h = L.hits[i];
        const uint32_t prim1 = __float_as_uint(h.y);
        if (prim1 == 0u) continue;                          // miss: le(light, ray) == 0 (lights/light.jl:41)
        const float4 o4 = L.ro[cur][i], d4 = L.rd[cur][i];
        const float3 w = xyz(L.rw[cur][i]);
        float3 d = xyz(d4);
        if (d.x == 0.0f) d.x = 0.0f;                        // the ray as intersect! left it (check_direction!)
        if (d.y == 0.0f) d.y = 0.0f;
        if (d.z == 0.0f) d.z = 0.0f;
        const uint32_t prim = prim1 - 1u;
        const float b2 = third_barycentric(L.sc, prim, xyz(o4), d);
        const Interaction it = build_interaction(L.sc, prim, xyz(o4), d, h.z, h.w, b2);
        const Frame fr = make_frame(it);
        LobeSet lobes;
        material_lobes(L.sc.materials[it.material], false, lobes);
        // direct lighting: one shadow ray per light (sampler.jl:83-92)
        for (int li = 0; li < L.sc.n_lights; ++li) {
            float3 wi, lpos;
            const float3 Li = sample_li_any(L.sc.lights[li], it.p, wi, lpos);
            if (is_black3(Li)) continue;
            const float3 f = bsdf_f(lobes, fr, it.wo, wi, LB_ALL);
            if (is_black3(f)) continue;
            const float3 contrib = w * ((f * Li) * fabsf(dot3(wi, it.ns)) / 1.0f);
            const float3 sdir = lpos - it.p;                 // spawn_ray(p0, p1): un-normalised, t_max = Inf (Q5)
            const int q = queue_claim(&L.counters[32]);
            if (q < L.cap_shadow) {
                L.so[q] = f4(it.p + 1e-6f * sdir, TR_INF);
                L.sd[q] = f4(sdir, d4.w);
                L.sc_contrib[q] = f4(contrib, 0.0f);
            } else L.counters[IC_OVERFLOW] = 1;
        }
        if (level + 1 <= L.max_depth) {
            // specular_reflect / specular_transmit (sampler.jl:103-199); u is irrelevant for the delta lobes
            #pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t type = (pass == 0 ? LB_REFLECTION : LB_TRANSMISSION) | LB_SPECULAR;
                const BSDFSample bs = bsdf_sample(lobes, fr, it.wo, 0.5f, 0.5f, type);
                const float adot = fabsf(dot3(bs.wi, it.ns));
                if (!(bs.pdf > 0.0f && !is_black3(bs.f) && adot != 0.0f)) continue;
                const float3 wn = w * (bs.f * adot / bs.pdf);
                const int q = queue_claim(&L.counters[level + 1]);
                if (q < L.cap_rays) {
                    L.ro[nxt][q] = f4(it.p + 1e-6f * bs.wi, TR_INF);     // spawn_ray(si, wi), Trace.jl:206-211
                    L.rd[nxt][q] = f4(bs.wi, d4.w);
                    L.rw[nxt][q] = f4(wn, 0.0f);
                } else L.counters[IC_OVERFLOW] = 1;
            }
        }
    }
}

// add_sample! with the reference's footprint / table indexing quirks (film.jl:134-164, Q4) and the clipping to the
// FilmTile bounds of the sample's 16x16 tile (film.jl:120-125).
__global__ void __launch_bounds__(256) k_wh_splat(WhittedLaunch L) {
    if (L.counters[IC_OVERFLOW]) return;                      // the host re-runs the batch in smaller pieces
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n_slots; i += gridDim.x * blockDim.x) {
        const float2 fp = L.filmpos[i];
        if (fp.x < -1e29f) continue;
        float4 a = L.accum[i];
        if (isnan(a.x) || isnan(a.y) || isnan(a.z)) { a.x = 0.0f; a.y = 0.0f; a.z = 0.0f; }
        int px, py, s, tile;
        slot_to_pixel(L, L.slot_begin + i, px, py, s, tile);
        const DeviceFilm& F = L.film;
        const int tx = tile % F.tiles_x, ty = tile / F.tiles_x;
        const int bx0 = F.sb_x0 + tx * 16, by0 = F.sb_y0 + ty * 16;
        const int bx1 = min(bx0 + 15, F.sb_x1), by1 = min(by0 + 15, F.sb_y1);
        const float tbx0 = fmaxf(ceilf((float)bx0 - 0.5f - F.rx), (float)F.crop_x0);
        const float tby0 = fmaxf(ceilf((float)by0 - 0.5f - F.ry), (float)F.crop_y0);
        const float tbx1 = fminf(floorf((float)bx1 - 0.5f + F.rx) + 1.0f, (float)F.crop_x1);
        const float tby1 = fminf(floorf((float)by1 - 0.5f + F.ry) + 1.0f, (float)F.crop_y1);
        const float dx = fp.x - 0.5f, dy = fp.y - 0.5f;
        const float p0x = fmaxf(ceilf(dx - F.rx), fmaxf(tbx0, 1.0f)), p0y = fmaxf(ceilf(dy - F.ry), fmaxf(tby0, 1.0f));
        const float p1x = fminf(floorf(dx + F.rx) + 1.0f, tbx1), p1y = fminf(floorf(dy + F.ry) + 1.0f, tby1);
        for (float y = p0y; y <= p1y; y += 1.0f) {
            const int oy = (int)clampf(floorf(fabsf((y - dy) * F.inv_ry * 16.0f)), 1.0f, 16.0f);
            for (float x = p0x; x <= p1x; x += 1.0f) {
                const int ox = (int)clampf(ceilf(fabsf((x - dx) * F.inv_rx * 16.0f)), 1.0f, 16.0f);
                const float wgt = __ldg(&F.table[(oy - 1) * 16 + (ox - 1)]);
                const int ix = (int)x - F.crop_x0, iy = (int)y - F.crop_y0;
                atomicAdd(&L.film_rgbw[(size_t)iy * F.width + ix], make_float4(a.x * wgt, a.y * wgt, a.z * wgt, wgt));
            }
        }
    }
}

// merge into the caller's film: xyz += to_XYZ(contrib), weight += w (film.jl:182-193)
// `skip` (device flag, may be null): set when a batch of this render overflowed a queue or a traversal failed - the host
// then repairs the render and finalizes again, unconditionally.
__global__ void k_film_finalize(const float4* __restrict__ rgbw, float4* __restrict__ film, int n, const int* __restrict__ skip) {
    if (skip && *skip) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 c = rgbw[i];
        float4 f = film[i];
        f.x += (0.412453f * c.x + 0.357580f * c.y) + 0.180423f * c.z;
        f.y += (0.212671f * c.x + 0.715160f * c.y) + 0.072169f * c.z;
        f.z += (0.019334f * c.x + 0.119193f * c.y) + 0.950227f * c.z;
        f.w += c.w;
        film[i] = f;
    }
}

// any[0] = 1 when one of the render's batches overflowed or a traversal ran out of stack
__global__ void k_wh_any_flag(const int* __restrict__ batch_flags, int n, const int* __restrict__ error_flag, int* __restrict__ any) {
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += 32) bad |= batch_flags[i];
    bad = __any_sync(0xffffffffu, bad != 0);
    if (threadIdx.x == 0) *any = (bad || *error_flag) ? 1 : 0;
}

__global__ void k_wh_batch_stats(int* counters, unsigned long long* stats, int max_depth, int cap_rays, int cap_shadow,
                                 int* batch_flag) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (batch_flag) *batch_flag = counters[IC_OVERFLOW];
        unsigned long long e = 0, s = 0;
        for (int l = 1; l <= max_depth; ++l) e += min(counters[l], cap_rays);
        s = min(counters[32], cap_shadow);
        if (!counters[IC_OVERFLOW]) { atomicAdd(&stats[ST_RAYS_EXTEND], e); atomicAdd(&stats[ST_RAYS_SHADOW], s); }
    }
}

// Enqueues one batch.  batch_flag == nullptr: synchronous (waits, checks the overflow flag, re-runs in halves).
// batch_flag != nullptr: asynchronous - the batch's overflow flag is left in *batch_flag (device memory) and the caller
// inspects all flags once after the last batch, so the host never waits on the GPU between batches (the per-batch
// sync cost ~0.9 ms of GPU idle time per batch on B200: 17 launches to enqueue behind an empty stream).
static int run_batch(trace_ctx* c, WhittedLaunch& L, long long begin, long long count, int depth_guard, int* batch_flag = nullptr) {
    if (count <= 0) return 0;
    L.slot_begin = begin;
    L.n_slots = (int)count;
    int* ic = ctx_icounters(c);
    unsigned long long* st = ctx_stats64(c);
    TR_CUDA(c, cudaMemsetAsync(ic, 0, 60 * sizeof(int), c->cur_stream));
    TR_CUDA(c, cudaMemsetAsync(ic + 64, 0, 64 * sizeof(int), c->cur_stream)); c->work_slot = 0;
    TR_CUDA(c, cudaMemsetAsync(ic + IC_OVERFLOW, 0, sizeof(int), c->cur_stream));
    const int g_stream = persistent_grid(c, 8), g_trav = persistent_grid(c, 16);
    k_wh_generate<<<g_stream, 256, 0, c->cur_stream>>>(L);
    c->stats.kernel_launches++;
    for (int level = 1; level <= L.max_depth; ++level) {
        const int cur = (level - 1) & 1;
        launch_extend(c, g_trav, L.sc, (const float4*)L.ro[cur], (const float4*)L.rd[cur], (const int*)(ic + level), L.cap_rays,
                      L.hits, st + ST_NODES, ctx_icounters_lane(c, 0) + IC_ERROR);
        k_wh_shade<<<occupancy_grid(c, k_wh_shade, 128), 128, 0, c->cur_stream>>>(L, level);
        c->stats.kernel_launches++;
    }
    // one any-hit launch over the shadow rays of ALL bounce levels of the batch (they only feed the accumulators): a
    // single wide launch instead of max_depth launches that each wait for their slowest ray
    launch_shadow(c, g_trav, L.sc, (const float4*)L.so, (const float4*)L.sd, (const float4*)L.sc_contrib,
                  (const int*)(ic + 32), L.cap_shadow, L.accum, st + ST_NODES, ctx_icounters_lane(c, 0) + IC_ERROR);
    k_wh_splat<<<g_stream, 256, 0, c->cur_stream>>>(L);
    k_wh_batch_stats<<<1, 32, 0, c->cur_stream>>>(ic, st, L.max_depth, L.cap_rays, L.cap_shadow, batch_flag);
    c->stats.kernel_launches += 2;
    TR_CUDA(c, cudaGetLastError());
    if (batch_flag) return 0;
    TR_CUDA(c, cudaMemcpyAsync(c->h_flags, ic + IC_OVERFLOW, sizeof(int), cudaMemcpyDeviceToHost, c->cur_stream));
    TR_CUDA(c, cudaMemcpyAsync(c->h_flags + 1, ctx_icounters_lane(c, 0) + IC_ERROR, sizeof(int), cudaMemcpyDeviceToHost, c->cur_stream));
    TR_CUDA(c, cudaStreamSynchronize(c->cur_stream));
    c->kev_collect();
    if (c->h_flags[1]) {
        cudaMemsetAsync(ctx_icounters_lane(c, 0) + IC_ERROR, 0, sizeof(int), c->cur_stream);
        return c->fail("traversal stack overflow (more than 64 pending nodes; the reference would throw a BoundsError, bvh.jl:222)");
    }
    if (c->h_flags[0]) {                                  // a queue overflowed: nothing was splatted, redo in halves
        c->stats.queue_overflows++;
        if (count < 2048 || depth_guard > 24) return c->fail("ray queue overflow that halving the batch cannot resolve");
        const long long half = count / 2;
        if (run_batch(c, L, begin, half, depth_guard + 1)) return 1;
        return run_batch(c, L, begin + half, count - half, depth_guard + 1);
    }
    return 0;
}

int whitted_render_device(trace_ctx* c, const trace_camera* cam, const trace_film_desc* film, int spp, int max_depth,
                          uint64_t seed, float* film_dev) {
    if (c->rank < 0 || c->rank >= c->world) return c->fail("rank %d outside world %d", c->rank, c->world);
    if (max_depth > 28) return c->fail("max_depth too large");
    WhittedLaunch L;
    memset(&L, 0, sizeof(L));                                   // padding too: the struct's bytes key the graph cache
    L.sc = c->scene;
    WhittedFrame frame;
    memset(&frame, 0, sizeof(frame));
    ctx_device_camera(cam, &frame.cam);
    frame.seed = seed;
    TR_CUDA(c, c->b_misc[3].ensure(sizeof(WhittedFrame)));
    TR_CUDA(c, cudaMemcpyAsync(c->b_misc[3].p, &frame, sizeof(frame), cudaMemcpyHostToDevice, c->stream));
    L.frame = c->b_misc[3].as<WhittedFrame>();
    if (ctx_device_film(c, film, &L.film, &c->b_misc[0])) return 1;
    L.spp = spp; L.max_depth = max_depth;
    // this rank's tiles: k = rank, rank + world, ...   (16x16 sample tiles, sampler.jl:24-31)
    const int total_tiles = L.film.tiles_x * L.film.tiles_y;
    std::vector<int> tiles;
    for (int k = c->rank; k < total_tiles; k += c->world) tiles.push_back(k);
    const long long total_slots = (long long)tiles.size() * 256 * spp;
    // batches of (at most) c->batch samples, evened out.  Deep bounce levels hold few but expensive rays (a launch
    // cannot end before its slowest ray: ~0.2 ms for a 1000-node walk), so batches are large and up to `lanes` of them
    // run concurrently on side streams, each lane with its own queues and counters.
    // c->batch bounds the samples in flight over ALL lanes (queue memory ~ 350 B per sample in flight)
    int K = 1;
    if (total_slots >= (long long)c->lanes * 131072) K = c->lanes;
    const long long per_lane = std::max<long long>(65536, c->batch / K);
    long long nb = std::max<long long>(K, (total_slots + per_lane - 1) / per_lane);
    nb = (nb + K - 1) / K * K;                                  // whole rounds of K lanes
    long long batch = (total_slots + nb - 1) / nb;
    batch = std::max<long long>(256, (batch + 255) / 256 * 256);
    nb = (total_slots + batch - 1) / batch;
    {
        // Batches are contiguous slot ranges, i.e. runs of the tile list.  In image order a run is a band of the image,
        // and bands differ wildly in cost (sky vs. geometry): lanes with cheap bands drain at once and the expensive
        // ones finish alone.  Deal groups of the rank's tiles (default: 2 tile rows - whole rows keep the BVH working set
        // of a batch compact) to the batches round-robin instead, so every batch samples the whole image.  Measured on
        // tess-1M: 1/8 of the frame 4.73 -> 4.53 ms, the full frame unchanged (profiles/r1_experiments.md).  The image
        // does not depend on the order: the RNG is keyed by the pixel.
        std::vector<int> dealt;
        dealt.reserve(tiles.size());
        const size_t G = c->deal > 0 ? (size_t)c->deal : std::max<size_t>(1, (size_t)(-c->deal) * L.film.tiles_x / c->world);
        const size_t groups = (tiles.size() + G - 1) / G;
        if (c->deal == 0) dealt = tiles;
        else
            for (long long j = 0; j < nb; ++j)
                for (size_t g = (size_t)j; g < groups; g += (size_t)nb)
                    for (size_t k = g * G; k < std::min(tiles.size(), (g + 1) * G); ++k) dealt.push_back(tiles[k]);
        tiles.swap(dealt);
    }
    TR_CUDA(c, c->b_misc[1].ensure((tiles.size() + 1) * sizeof(int)));
    if (c->wh_tiles_dev != c->b_misc[1].p || c->wh_tiles != tiles) {        // (re-rendering the same film: already there)
        if (!tiles.empty()) TR_CUDA(c, cudaMemcpyAsync(c->b_misc[1].p, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        TR_CUDA(c, cudaStreamSynchronize(c->stream));
        c->wh_tiles = tiles; c->wh_tiles_dev = c->b_misc[1].p;
    }
    L.tiles = c->b_misc[1].as<int>();
    std::vector<long long> b_begin, b_count;
    for (long long bi = 0; bi < nb; ++bi) { b_begin.push_back(bi * batch); b_count.push_back(std::min(batch, total_slots - bi * batch)); }
    const size_t cap_rays = (size_t)batch * (size_t)c->cap_percent / 100;
    const int shadow_mult = std::max(1, std::min(L.sc.n_lights, 4));
    const size_t cap_shadow = cap_rays * shadow_mult * 2;     // all bounce levels of a batch share one shadow queue
    L.cap_rays = (int)cap_rays; L.cap_shadow = (int)cap_shadow;
    for (int k = 0; k < 7; ++k) TR_CUDA(c, c->b_queue[k].ensure(K * cap_rays * sizeof(float4)));
    for (int k = 7; k < 10; ++k) TR_CUDA(c, c->b_queue[k].ensure(K * cap_shadow * sizeof(float4)));
    TR_CUDA(c, c->b_queue[10].ensure((size_t)K * batch * sizeof(float4)));
    TR_CUDA(c, c->b_queue[11].ensure((size_t)K * batch * sizeof(float2)));
    const size_t npix = (size_t)L.film.width * L.film.height;
    TR_CUDA(c, c->b_queue[12].ensure(npix * sizeof(float4)));
    L.film_rgbw = c->b_queue[12].as<float4>();
    std::vector<WhittedLaunch> lane((size_t)K, L);
    for (int l = 0; l < K; ++l) {
        WhittedLaunch& W = lane[l];
        W.ro[0] = c->b_queue[0].as<float4>() + l * cap_rays; W.ro[1] = c->b_queue[1].as<float4>() + l * cap_rays;
        W.rd[0] = c->b_queue[2].as<float4>() + l * cap_rays; W.rd[1] = c->b_queue[3].as<float4>() + l * cap_rays;
        W.rw[0] = c->b_queue[4].as<float4>() + l * cap_rays; W.rw[1] = c->b_queue[5].as<float4>() + l * cap_rays;
        W.hits = c->b_queue[6].as<float4>() + l * cap_rays;
        W.so = c->b_queue[7].as<float4>() + l * cap_shadow; W.sd = c->b_queue[8].as<float4>() + l * cap_shadow;
        W.sc_contrib = c->b_queue[9].as<float4>() + l * cap_shadow;
        W.accum = c->b_queue[10].as<float4>() + (size_t)l * batch; W.filmpos = c->b_queue[11].as<float2>() + (size_t)l * batch;
        W.counters = ctx_icounters_lane(c, l);
    }
    TR_CUDA(c, c->b_misc[2].ensure((size_t)(nb + 1) * sizeof(int)));
    int* d_flags = c->b_misc[2].as<int>();
    TR_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    // everything up to the lanes' join: film clear, then batch bi on lane bi % K
    auto enqueue = [&]() -> int {
        TR_CUDA(c, cudaMemsetAsync(L.film_rgbw, 0, npix * sizeof(float4), c->stream));
        if (K > 1) {
            TR_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
            for (int l = 0; l < K; ++l) TR_CUDA(c, cudaStreamWaitEvent(c->side[l], c->ev_fork, 0));
        }
        int rc = 0;
        for (long long bi = 0; bi < nb && !rc; ++bi) {
            const int l = (int)(bi % K);
            c->cur_lane = l;
            c->cur_stream = K > 1 ? c->side[l] : c->stream;
            rc = run_batch(c, lane[l], b_begin[bi], b_count[bi], 0, d_flags + bi);
        }
        if (K > 1) {
            for (int l = 0; l < K; ++l) {
                cudaEventRecord(c->ev_join[l], c->side[l]);
                cudaStreamWaitEvent(c->stream, c->ev_join[l], 0);
            }
        }
        c->cur_lane = 0;
        c->cur_stream = c->stream;
        return rc;
    };
    int rc = 0;
    if (c->graph && !c->time_kernels && (size_t)c->stream > 2) {       // (legacy / per-thread default streams cannot be captured)
        std::string key((const char*)lane.data(), lane.size() * sizeof(WhittedLaunch));
        key.append((const char*)b_count.data(), b_count.size() * sizeof(long long));
        const long long extra[] = {nb, batch, K, total_slots, (long long)(size_t)d_flags, c->slab, c->persist, c->count_nodes,
                                   (long long)(size_t)c->stream};
        key.append((const char*)extra, sizeof(extra));
        if (!c->wh_graph || key != c->wh_graph_key) {
            if (c->wh_graph) { cudaGraphExecDestroy(c->wh_graph); c->wh_graph = nullptr; }
            const unsigned long long before = c->stats.kernel_launches;
            TR_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            rc = enqueue();
            cudaGraph_t g = nullptr;
            const cudaError_t e_end = cudaStreamEndCapture(c->stream, &g);
            c->wh_graph_launches[0] = c->stats.kernel_launches - before;
            c->stats.kernel_launches = before;
            if (rc || e_end != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                return rc ? 1 : c->fail("CUDA graph capture of the render failed: %s", cudaGetErrorString(e_end));
            }
            const cudaError_t e_inst = cudaGraphInstantiate(&c->wh_graph, g, 0);
            cudaGraphDestroy(g);
            if (e_inst != cudaSuccess) { c->wh_graph = nullptr; return c->fail("cudaGraphInstantiate: %s", cudaGetErrorString(e_inst)); }
            c->wh_graph_key = key;
        }
        TR_CUDA(c, cudaGraphLaunch(c->wh_graph, c->stream));
        c->stats.kernel_launches += c->wh_graph_launches[0];
    } else {
        rc = enqueue();
    }
    if (rc) return 1;
    // optimistic tail: fold the flags, merge the film unless something went wrong, read the flags back - ONE host wait
    int* d_err = ctx_icounters_lane(c, 0) + IC_ERROR;
    k_wh_any_flag<<<1, 32, 0, c->stream>>>(d_flags, (int)nb, d_err, d_flags + nb);
    if (c->film_upload_pending) { TR_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->film_upload_pending = false; }
    k_film_finalize<<<persistent_grid(c, 4), 256, 0, c->stream>>>(L.film_rgbw, (float4*)film_dev, (int)npix, d_flags + nb);
    c->stats.kernel_launches += 2;
    TR_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    std::vector<int> h_flags((size_t)nb + 2, 0);
    TR_CUDA(c, cudaMemcpyAsync(h_flags.data(), d_flags, (size_t)(nb + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TR_CUDA(c, cudaMemcpyAsync(&h_flags[(size_t)nb + 1], d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->kev_collect();
    if (h_flags[(size_t)nb + 1]) {
        cudaMemsetAsync(d_err, 0, sizeof(int), c->stream);
        return c->fail("traversal stack overflow (more than 64 pending nodes; the reference would throw a BoundsError, bvh.jl:222)");
    }
    if (h_flags[(size_t)nb]) {
        for (long long bi = 0; bi < nb; ++bi) {
            if (!h_flags[bi]) continue;                     // overflowed batches were not splatted: redo them in halves
            c->stats.queue_overflows++;
            const long long b = b_begin[bi], cnt = b_count[bi], half = cnt / 2;
            if (cnt < 2048) return c->fail("ray queue overflow that halving the batch cannot resolve");
            if (run_batch(c, lane[0], b, half, 1) || run_batch(c, lane[0], b + half, cnt - half, 1)) return 1;
        }
        k_film_finalize<<<persistent_grid(c, 4), 256, 0, c->stream>>>(L.film_rgbw, (float4*)film_dev, (int)npix, nullptr);
        c->stats.kernel_launches++;
        TR_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        TR_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.ms_total = ms;
    return 0;
}
