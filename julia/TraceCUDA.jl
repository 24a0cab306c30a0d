# TraceCUDA.jl — the reference-side binding a Trace.jl maintainer would add to use libtrace_cuda.so as the GPU backend.
#
# NOT RUNNABLE IN THIS IMAGE (no julia binary); it documents, in the reference's own language, exactly which calls the
# C ABI (include/trace_cuda.h) replaces.  The Python package trace.jl_b200/ is the runnable mirror of this file.
#
#   using Trace, TraceCUDA
#   scene |> TraceCUDA.gpu(Trace.SPPMIntegrator(camera, 0.025f0, 5, 100))      # instead of  scene |> integrator
#
module TraceCUDA

using Trace
using GeometryBasics

const LIB = get(ENV, "TRACE_CUDA_LIB", "libtrace_cuda.so")

# ---- POD mirrors of include/trace_cuda.h --------------------------------------------------------------------------
struct BVHNode            # trace_bvh_node, 32 B
    bmin::NTuple{3,Float32}
    bmax::NTuple{3,Float32}
    offset::UInt32
    meta::UInt32
end
struct Prim               # trace_prim
    kind::UInt32
    index::UInt32
    material::UInt32
    original::UInt32
end
struct SphereP            # trace_sphere
    m::NTuple{16,Float32}
    inv_m::NTuple{16,Float32}
    radius::Float32; z_min::Float32; z_max::Float32; θ_min::Float32; θ_max::Float32; ϕ_max::Float32
    flip::UInt32; pad::UInt32
end
struct MaterialP          # trace_material
    kind::UInt32
    a::NTuple{3,Float32}; b::NTuple{3,Float32}
    η::Float32; rough_u::Float32; rough_v::Float32
    remap::UInt32
end
struct LightP             # trace_light
    kind::UInt32
    m::NTuple{16,Float32}; inv_m::NTuple{16,Float32}
    i::NTuple{3,Float32}; position::NTuple{3,Float32}
    cos_total_width::Float32; cos_falloff_start::Float32
end
struct SceneDesc          # trace_scene_desc
    n_nodes::Int64; nodes::Ptr{BVHNode}
    n_prims::Int64; prims::Ptr{Prim}
    n_tris::Int64; tri_vertices::Ptr{Float32}; tri_normals::Ptr{Float32}; tri_flags::Ptr{UInt8}
    n_spheres::Int64; spheres::Ptr{SphereP}
    n_materials::Int64; materials::Ptr{MaterialP}
    n_lights::Int64; lights::Ptr{LightP}
end
struct CameraP            # trace_camera
    raster_to_camera::NTuple{16,Float32}
    camera_to_world::NTuple{16,Float32}
    lens_radius::Float32; focal_distance::Float32; shutter_open::Float32; shutter_close::Float32
end
struct FilmP              # trace_film_desc
    crop::NTuple{4,Int32}
    filter_radius::NTuple{2,Float32}
    filter_table::NTuple{256,Float32}
    scale::Float32
end

rowmajor(m::Mat4f) = ntuple(k -> m[(k - 1) ÷ 4 + 1, (k - 1) % 4 + 1], 16)

# ---- context ------------------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    lock::ReentrantLock
    function Context(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:trace_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cvoid}), out, device, C_NULL)
        rc == 0 || error("trace_create failed ($rc): no CUDA device (there is no CPU fallback)")
        ctx = new(out[], ReentrantLock())
        finalizer(c -> ccall((:trace_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h), ctx)
    end
end
check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:trace_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))

# ---- flattening Trace.Scene -> SceneDesc (what trace.jl_b200/scene.py:FlatScene does) -------------------------------
# BVHAccel.nodes are already the reference's LinearBVH array (1-based); convert to 0-based trace_bvh_node.
function flatten_nodes(bvh::Trace.BVHAccel)
    map(bvh.nodes) do n
        if n isa Trace.LinearBVHLeaf
            BVHNode(Tuple(n.bounds.p_min), Tuple(n.bounds.p_max), n.primitives_offset - 1, 0xC0000000 | n.n_primitives)
        else
            BVHNode(Tuple(n.bounds.p_min), Tuple(n.bounds.p_max), n.second_child_offset - 1, UInt32(n.split_axis - 1) << 30)
        end
    end
end
# (primitive / material / light flattening is mechanical: see scene.py; nested BVHAccel primitives are spliced in place.)

# ---- the two functors ------------------------------------------------------------------------------------------------
struct GPU{I<:Trace.Integrator}
    integrator::I
    ctx::Context
end
gpu(i::Trace.Integrator, ctx::Context = Context()) = GPU(i, ctx)

camera_pod(c::Trace.PerspectiveCamera) = CameraP(
    rowmajor(c.core.raster_to_camera.m), rowmajor(c.core.core.camera_to_world.m),
    c.core.lens_radius, c.core.focal_distance, c.core.core.shutter_open, c.core.core.shutter_close)

function film_pod(f::Trace.Film)
    FilmP(Int32.((f.crop_bounds.p_min..., f.crop_bounds.p_max...)), Tuple(f.filter.radius),
          ntuple(k -> f.filter_table[(k - 1) ÷ 16 + 1, (k - 1) % 16 + 1], 256), f.scale)
end

# (i::SamplerIntegrator)(scene)  — src/integrators/sampler.jl:12-56
function (g::GPU{Trace.WhittedIntegrator})(scene::Trace.Scene)
    i, ctx = g.integrator, g.ctx
    film = Trace.get_film(i.camera)
    H, W = size(film.pixels)
    buf = zeros(Float32, 4, W, H)                       # row-major [y][x][4] on the C side
    for y in 1:H, x in 1:W
        p = film.pixels[y, x]
        buf[1:3, x, y] .= p.xyz; buf[4, x, y] = p.filter_weight_sum
    end
    lock(ctx.lock) do
        upload!(ctx, scene)
        GC.@preserve buf check(ctx, ccall((:trace_render_whitted, LIB), Cint,
            (Ptr{Cvoid}, Ref{CameraP}, Ref{FilmP}, Cint, Cint, UInt64, Ptr{Float32}),
            ctx.h, camera_pod(i.camera), film_pod(film), i.sampler.samples_per_pixel, i.max_depth, rand(UInt64), buf))
    end
    for y in 1:H, x in 1:W
        film.pixels[y, x].xyz = Point3f(buf[1:3, x, y]); film.pixels[y, x].filter_weight_sum = buf[4, x, y]
    end
    Trace.save(film)
end

# (i::SPPMIntegrator)(scene)  — src/integrators/sppm.jl:132-173
function (g::GPU{Trace.SPPMIntegrator})(scene::Trace.Scene)
    i, ctx = g.integrator, g.ctx
    film = Trace.get_film(i.camera)
    H, W = size(film.pixels)
    rgb = zeros(Float32, 3, W, H)
    function on_image(::Ptr{Cvoid}, iteration::Cint, p::Ptr{Float32})::Cvoid
        img = unsafe_wrap(Array, p, (3, W, H))
        Trace.set_image!(film, [Trace.RGBSpectrum(img[1, x, y], img[2, x, y], img[3, x, y]) for y in 1:H, x in 1:W])
        Trace.save(film)
        nothing
    end
    cb = @cfunction($on_image, Cvoid, (Ptr{Cvoid}, Cint, Ptr{Float32}))
    lock(ctx.lock) do
        upload!(ctx, scene)
        GC.@preserve rgb cb check(ctx, ccall((:trace_render_sppm, LIB), Cint,
            (Ptr{Cvoid}, Ref{CameraP}, Ref{FilmP}, Cfloat, Cint, Cint, Int64, Cint, UInt64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float32}),
            ctx.h, camera_pod(i.camera), film_pod(film), i.initial_search_radius, i.max_depth, i.n_iterations,
            i.photons_per_iteration, i.write_frequency, rand(UInt64), cb, C_NULL, rgb))
    end
end

# batch intersect!(scene, rays) / intersect_p — src/accel/bvh.jl:212-299
function intersect!(ctx::Context, o::Matrix{Float32}, d::Matrix{Float32}, t_max::Vector{Float32})
    n = size(o, 2)
    prim = zeros(UInt32, n); bary = zeros(Float32, 2, n)
    GC.@preserve o d t_max prim bary check(ctx, ccall((:trace_intersect, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Int64, Ptr{UInt32}, Ptr{Float32}),
        ctx.h, o, d, t_max, n, prim, bary))
    prim, bary            # t_max now holds the hit distances, like ray.t_max after intersect!
end

function upload! end      # SceneDesc assembly + ccall(:trace_scene_upload, ...): see trace.jl_b200/scene.py

end # module
