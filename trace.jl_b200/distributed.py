"""Multi-GPU drivers: one process per GPU, scene replicated, torch.distributed (NCCL over NVLink) for the two
exchange steps the path has:

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


def shares_torch_stream(ctx):
    """True when the context enqueues on torch's current stream, so torch / NCCL work orders against it by itself."""
    import torch
    return bool(ctx.stream_handle) and ctx.stream_handle == torch.cuda.current_stream().cuda_stream


class _Fence:
    """Orders a block of torch work (collectives) against the context's stream.  Nothing to do when both use the same
    stream; otherwise the context's stream is drained before and torch's current stream after."""

    def __init__(self, ctx):
        self.ctx, self.shared = ctx, shares_torch_stream(ctx)

    def __enter__(self):
        if not self.shared:
            self.ctx.synchronize()
        return self

    def __exit__(self, *exc):
        if not self.shared:
            import torch
            torch.cuda.current_stream().synchronize()
        return False


def allreduce_sum(tensor, group=None):
    """all_reduce(sum) when a process group exists; identity otherwise.  Backend-agnostic (nccl on GPUs, gloo in tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def render_whitted_sharded(ctx, scene, camera, spp, max_depth, seed, film_tensor, rank=0, world=1, group=None, reduce=True):
    """Render this rank's tiles into `film_tensor` (torch float32 [H, W, 4] on the context's GPU), then reduce to rank 0."""
    import torch.distributed as dist
    flat = ctx.upload(scene)
    if flat.has_unshaded:
        raise ValueError("every primitive needs a material")
    ctx.set_option("world", world)
    ctx.set_option("rank", rank)
    cam, fd = camera.pod(), camera.film.desc()
    ctx.check(ctx.lib.trace_render_whitted_device(ctx.h, C.byref(cam), C.byref(fd), int(spp), int(max_depth),
                                                  C.c_uint64(seed), C.c_void_p(film_tensor.data_ptr())))
    if reduce and world > 1:
        with _Fence(ctx):
            dist.reduce(film_tensor, dst=0, op=dist.ReduceOp.SUM, group=group)
    return film_tensor


class ShardedWhittedRenderer:
    """Whitted render over `world` ranks with a HOST film (the reference's accumulate-into-film semantics, film.jl:182-193):
    every rank renders its tiles into a zeroed device film, one reduce(sum) to rank 0, and rank 0 adds the caller's
    film - uploaded on a copy stream WHILE the render runs - and copies the sum back.  `host_film` (rank 0; pinned for
    the copies to be asynchronous) is float32 [H, W, 4]; other ranks pass None."""

    def __init__(self, ctx, scene, camera, rank=0, world=1, group=None):
        import torch
        self.ctx, self.scene, self.camera, self.rank, self.world, self.group = ctx, scene, camera, rank, world, group
        h, w = camera.film.pixels.shape[:2]
        dev = f"cuda:{torch.cuda.current_device()}"
        self.film_dev = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        self.staging = torch.zeros((h, w, 4), dtype=torch.float32, device=dev) if rank == 0 else None
        self.copy_stream = torch.cuda.Stream() if rank == 0 else None
        self.uploaded = torch.cuda.Event() if rank == 0 else None

    def render(self, host_film, spp, max_depth, seed):
        import torch
        cur = torch.cuda.current_stream()
        if self.rank == 0:
            self.copy_stream.wait_stream(cur)                    # (the previous render's add_ has read `staging`)
            with torch.cuda.stream(self.copy_stream):
                self.staging.copy_(host_film, non_blocking=True)
                self.uploaded.record()
        self.film_dev.zero_()
        render_whitted_sharded(self.ctx, self.scene, self.camera, spp, max_depth, seed, self.film_dev, self.rank, self.world,
                               self.group, reduce=True)
        if self.rank == 0:
            cur.wait_event(self.uploaded)
            self.film_dev.add_(self.staging)
            host_film.copy_(self.film_dev, non_blocking=True)
        cur.synchronize()
        return host_film


class SPPMSession:
    """Stepwise SPPM (trace_sppm_begin / camera_pass / photon_pass / update / image) with photon sharding."""

    def __init__(self, ctx, scene, camera, initial_search_radius, max_depth, photons_per_iteration=-1, seed=0x5EED0001,
                 rank=0, world=1, group=None):
        import torch
        self.ctx, self.camera, self.rank, self.world, self.group = ctx, camera, rank, world, group
        flat = ctx.upload(scene)
        if flat.has_unshaded:
            raise ValueError("every primitive needs a material")
        if photons_per_iteration <= 0:
            photons_per_iteration = int(camera.film.crop_bounds.area())
        self.photons = int(photons_per_iteration)
        cam, fd = camera.pod(), camera.film.desc()
        ctx.set_option("world", world)
        ctx.set_option("rank", rank)
        ctx.check(ctx.lib.trace_sppm_begin(ctx.h, C.byref(cam), C.byref(fd), float(initial_search_radius), int(max_depth),
                                           self.photons, C.c_uint64(seed)))
        self.buffers = None
        self._vp_recv = None
        if world > 1:
            dev = f"cuda:{torch.cuda.current_device()}"
            self.buffers = []
            for which in range(7):          # 0 flux, 1 Ld, 2..6 visible-point records (storage order, rank-major slices)
                n = C.c_int64()
                ptr = ctx.lib.trace_sppm_buffer_device(ctx.h, which, C.byref(n))
                self.buffers.append(torch.as_tensor(_DevicePtr(ptr, n.value), device=dev))
        self.iteration = 0

    def _all_gather(self, which):
        import torch.distributed as dist
        t = self.buffers[which]
        n = t.numel() // self.world
        dist.all_gather_into_tensor(t, t[self.rank * n:(self.rank + 1) * n].clone(), group=self.group)

    def _gather_visible_points(self):
        """ONE all-gather for the five visible-point arrays: this rank's five slices are packed into one send buffer,
        gathered as [world][5][slice], and scattered back into the arrays' rank slices (two small copy kernels instead
        of four more collectives, whose launch latency dominates at these sizes)."""
        import torch
        import torch.distributed as dist
        arrays = self.buffers[2:7]
        n = arrays[0].numel() // self.world
        send = torch.stack([a[self.rank * n:(self.rank + 1) * n] for a in arrays])          # [5][n]
        if self._vp_recv is None:
            self._vp_recv = torch.empty((self.world, len(arrays), n), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(self._vp_recv, send, group=self.group)
        for k, a in enumerate(arrays):
            a.view(self.world, n).copy_(self._vp_recv[:, k, :])

    def step(self):
        """One SPPM iteration (sppm.jl:153-165) over all ranks."""
        self.iteration += 1
        ctx = self.ctx
        b, e = photon_range(self.photons, self.rank, self.world)
        ctx.check(ctx.lib.trace_sppm_trace_photons(ctx.h, self.iteration, b, e))     # asynchronous: overlaps what follows
        ctx.check(ctx.lib.trace_sppm_camera_pass(ctx.h, self.iteration))
        if self.world > 1:
            with _Fence(ctx):
                self._gather_visible_points()
            ctx.check(ctx.lib.trace_sppm_build_grid(ctx.h))
        ctx.check(ctx.lib.trace_sppm_photon_pass(ctx.h, self.iteration, b, e))
        if self.world > 1:
            with _Fence(ctx):
                allreduce_sum(self.buffers[0], self.group)
        ctx.check(ctx.lib.trace_sppm_update(ctx.h))

    def image(self):
        if self.world > 1:
            with _Fence(self.ctx):
                self._all_gather(1)         # Ld is accumulated only by the rank that owns the row
        h, w = self.camera.film.pixels.shape[:2]
        rgb = np.zeros((h, w, 3), dtype=np.float32)
        self.ctx.check(self.ctx.lib.trace_sppm_image(self.ctx.h, max(1, self.iteration), _lib.ptr(rgb)))
        return rgb

    def close(self):
        self.ctx.check(self.ctx.lib.trace_sppm_end(self.ctx.h))
