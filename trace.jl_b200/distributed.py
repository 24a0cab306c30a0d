"""Multi-GPU drivers: one process per GPU, scene replicated.  The two exchange steps the path has run
INSIDE libtrace_cuda.so on its own NCCL communicator (include/trace_cuda.h: trace_comm_*); this module only hands the
communicator id from rank 0 to the other ranks (through any torch.distributed backend, or any other channel) and
keeps the pure-Python partition helpers that the CPU tests exercise with gloo:

  * Whitted: 16x16 sample tiles are dealt round-robin to the ranks (tile k -> rank k % world, the reference's
    Threads.@threads loop over tiles, src/integrators/sampler.jl:24); every rank splats into a private full-resolution
    film; ONE reduce(sum) of the (X, Y, Z, w) film at the end.
  * SPPM: the camera pass is sharded by image rows (row y -> rank y % world; the RNG is keyed by the raster pixel, so
    the visible points do not depend on the rank count) and the visible-point records are all-gathered (5 float4
    arrays); every rank then builds the identical hash grid, traces its slice of the iteration's photons (Halton index
    range, src/integrators/sppm.jl:328-336) into a private (Phi, M) buffer, ONE all_reduce(sum) of that buffer per
    iteration, then the identical per-pixel update.  Ld stays sharded until the image is assembled (one all-gather).

The host-side partition helpers are pure Python so they can be tested with gloo on CPU.
"""
import ctypes as C

import numpy as np

from . import _lib


def tile_shard(n_tiles, rank, world):
    """Tiles owned by `rank`: k = rank, rank + world, ... (must match whitted.cu)."""
    return list(range(rank, n_tiles, world))


def photon_range(photons_per_iteration, rank, world):
    """Contiguous photon-index slice [begin, end) of one iteration for `rank`."""
    p = int(photons_per_iteration)
    return (p * rank) // world, (p * (rank + 1)) // world


def storage_layout(width, height, world):
    """SPPM per-pixel arrays in STORAGE order (must match sppm.cu: storage_to_raster / raster_to_storage): image row y
    belongs to rank y % world and is that rank's local row y // world; every rank's rows are contiguous and padded to
    chunk_rows = ceil(height / world) rows, so rank r owns the slice [r, r + 1) * chunk_rows * width of every array.
    Returns (chunk_rows, n_slots)."""
    chunk_rows = (int(height) + int(world) - 1) // int(world)
    return chunk_rows, int(world) * chunk_rows * int(width)


def raster_to_storage(x, y, width, height, world):
    chunk_rows, _ = storage_layout(width, height, world)
    return ((y % world) * chunk_rows + y // world) * width + x


def storage_to_raster(slot, width, height, world):
    """(x, y) of a storage slot, or None for a padding slot."""
    chunk_rows, _ = storage_layout(width, height, world)
    row, x = divmod(int(slot), int(width))
    owner, local = divmod(row, chunk_rows)
    y = local * world + owner
    return (x, y) if y < height else None


def n_sample_tiles(film):
    sb = film.get_sample_bounds()
    ext = sb.p_max - sb.p_min
    return int(np.floor((ext[0] + 16) / 16)) * int(np.floor((ext[1] + 16) / 16))


class _DevicePtr:
    """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def broadcast_id(make_id, rank, group=None):
    """Rank 0 creates the communicator id (bytes), every rank returns it.  Any torch.distributed backend will do (it
    is 128 bytes on the host); tested with gloo on CPU."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(_lib.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        dev = buf.cuda()
        dist.broadcast(dev, src=0, group=group)
        buf = dev.cpu()
    else:
        dist.broadcast(buf, src=0, group=group)
    return bytes(buf.numpy().tobytes())


def allreduce_sum(tensor, group=None):
    """all_reduce(sum) when a process group exists; identity otherwise (host-side bookkeeping of the drivers, e.g. ray
    counters; backend-agnostic: nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def init_comm(ctx, rank, world, group=None):
    """Give `ctx` the library's NCCL communicator over the ranks of the torch.distributed group."""
    from .render import comm_unique_id
    if world <= 1:
        return
    ctx.comm_init(broadcast_id(comm_unique_id, rank, group), rank, world)


def render_whitted_sharded(ctx, scene, camera, spp, max_depth, seed, film_tensor, rank=0, world=1, reduce=False):
    """One Whitted render into `film_tensor` (torch float32 [H, W, 4] on the context's GPU).  With a communicator
    (init_comm) this is the whole multi-GPU render: the library renders the rank's tiles and sums the films (option
    "film_mode": everything on rank 0, or one band per rank).  Without one, (rank, world) only select the tiles - used to
    play one rank of N on a single GPU."""
    flat = ctx.upload(scene)
    if flat.has_unshaded:
        raise ValueError("every primitive needs a material")
    if not ctx.comm_info()["nccl_version"]:
        ctx.set_option("world", world)
        ctx.set_option("rank", rank)
    cam, fd = camera.pod(), camera.film.desc()
    ctx.check(ctx.lib.trace_render_whitted_device(ctx.h, C.byref(cam), C.byref(fd), int(spp), int(max_depth),
                                                  C.c_uint64(seed), C.c_void_p(film_tensor.data_ptr())))
    return film_tensor


class SPPMSession:
    """SPPM over the ranks of the context's communicator (or on one GPU): trace_sppm_begin, then step() = one iteration
    (sppm.jl:153-165) enqueued by trace_sppm_iterate - camera paths sharded by image rows, photons by index range, the
    all-gather of the visible points and the all-reduce of (Phi, M) inside the library.  `emulate=(rank, world)` plays
    one rank without a communicator through the stepwise API (the caller moves the slices; tests only)."""

    def __init__(self, ctx, scene, camera, initial_search_radius, max_depth, photons_per_iteration=-1, seed=0x5EED0001,
                 emulate=None):
        self.ctx, self.camera = ctx, camera
        flat = ctx.upload(scene)
        if flat.has_unshaded:
            raise ValueError("every primitive needs a material")
        if photons_per_iteration <= 0:
            photons_per_iteration = int(camera.film.crop_bounds.area())
        self.photons = int(photons_per_iteration)
        cam, fd = camera.pod(), camera.film.desc()
        info = ctx.comm_info()
        self.rank, self.world = info["rank"], info["world"]
        if emulate is not None:
            self.rank, self.world = emulate
            ctx.set_option("world", self.world)
            ctx.set_option("rank", self.rank)
        elif not info["nccl_version"]:
            ctx.set_option("world", 1)
            ctx.set_option("rank", 0)
            self.rank, self.world = 0, 1
        ctx.check(ctx.lib.trace_sppm_begin(ctx.h, C.byref(cam), C.byref(fd), float(initial_search_radius), int(max_depth),
                                           self.photons, C.c_uint64(seed)))
        self.iteration = 0

    def step(self, n=1):
        self.ctx.check(self.ctx.lib.trace_sppm_iterate(self.ctx.h, self.iteration + 1, int(n)))
        self.iteration += int(n)

    def buffers(self):
        """The seven per-pixel device buffers (storage order) as torch views: 0 flux, 1 Ld, 2..6 visible points."""
        import torch
        out = []
        for which in range(7):
            n = C.c_int64()
            ptr = self.ctx.lib.trace_sppm_buffer_device(self.ctx.h, which, C.byref(n))
            out.append(torch.as_tensor(_DevicePtr(ptr, n.value), device=f"cuda:{torch.cuda.current_device()}"))
        return out

    def image(self):
        h, w = self.camera.film.pixels.shape[:2]
        rgb = np.zeros((h, w, 3), dtype=np.float32)
        self.ctx.check(self.ctx.lib.trace_sppm_image(self.ctx.h, max(1, self.iteration), _lib.ptr(rgb)))
        return rgb

    def close(self):
        self.ctx.check(self.ctx.lib.trace_sppm_end(self.ctx.h))
