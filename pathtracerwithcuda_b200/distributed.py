"""Sample-sharded multi-GPU rendering: the scene is replicated on every rank,
rank r of G renders passes r+1, r+1+G, ... (each pass keeps the seed it has in a single-GPU run,
Kernel/path_tracer_kernel.cu:712), accumulates locally, and ONE sum-reduce of the float
accumulation buffer per image lands on rank 0, which then runs the mean / gamma / 8-bit step.
No data-path collective exists besides that reduce (and the one-off broadcast of the parsed scene).

Two drivers:
* `DistRenderer` — the product path, a THIN caller of the C ABI (include/ptb200.h, csrc/multi.inc): the library owns the NCCL
  communicator, broadcasts the parsed scene from rank 0, shards the passes and reduces on its own render stream; the host only
  hands the 128-byte NCCL id around (torch.distributed / gloo / MPI / a file — `torch_exchange` below uses torch.distributed).
  In ONE process with all GPUs use `pathtracerwithcuda_b200.MultiRenderer` (ptb_multi_*) instead.
* `ShardedRenderer` — the same sharding logic over an abstract backend and torch.distributed collectives; it is what the CPU
  tests drive with gloo (the oracle as stand-in compute) and what tests emulate two ranks with on one GPU.
"""
import numpy as np


def shard_passes(total_passes, rank, world_size):
    """(first_pass, stride, count) of the 1-based pass indices rank `rank` renders."""
    if total_passes <= rank:
        return rank + 1, world_size, 0
    count = (total_passes - rank - 1) // world_size + 1
    return rank + 1, world_size, count


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias a raw device pointer."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class CudaBackend:
    """Adapter over pathtracerwithcuda_b200.Renderer for ShardedRenderer."""

    def __init__(self, renderer):
        self.r = renderer
        self._tensor = None

    def clear(self):
        self.r.clear()

    def render_strided(self, first, stride, count):
        if count > 0:
            self.r.render_strided(first, stride, count)

    def accumulation_tensor(self):
        import torch
        if self._tensor is None:
            ptr = self.r.image_device_ptr()
            if not ptr:
                raise RuntimeError("renderer has no device accumulation buffer (host-only handle)")
            self._holder = _CudaArray(ptr, self.r.width * self.r.height * 3)
            self._tensor = torch.as_tensor(self._holder, device="cuda:%d" % self.r.device)
        return self._tensor

    def synchronize(self):
        self.r.synchronize()

    def after_collective(self):
        """torch.distributed's NCCL ops run on torch's own stream and only order themselves against torch's CURRENT stream;
        the renderer works on its own stream, so the host waits here until the collective has written the buffer."""
        import torch
        torch.cuda.current_stream(torch.device("cuda", self.r.device)).synchronize()

    def finalize(self, total_passes):
        self.r.finalize(total_passes)

    def image_u8(self):
        return self.r.image_u8()

    def image_f32(self):
        return self.r.image_f32()


def torch_exchange(dist):
    """unique-id exchange for DistRenderer over an initialised torch.distributed group (plumbing only)."""
    def exchange(payload):
        box = [payload]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return exchange


class DistRenderer:
    """One process per GPU through the C ABI.  exchange(bytes-or-None) -> bytes must return rank 0's argument on every rank."""

    def __init__(self, renderer, rank=0, world_size=1, exchange=None):
        from . import api
        self.r = renderer
        self.rank, self.world_size = rank, world_size
        self.local_passes = 0
        self.timing = {}
        if world_size > 1:
            if exchange is None:
                raise ValueError("DistRenderer: world_size > 1 needs an exchange callable for the NCCL id")
            uid = exchange(api.dist_unique_id() if rank == 0 else None)
            renderer.dist_init(rank, world_size, uid)

    def load_scene(self, scene_json_path, asset_root="", root=0):
        """rank `root` reads and parses the scene files; every other rank receives the parsed scene over NCCL."""
        if self.world_size == 1:
            self.r.load_scene(scene_json_path, asset_root)
            return
        import time
        t0 = time.perf_counter()
        try:
            self.r.dist_load_scene(scene_json_path if self.rank == root else "", asset_root if self.rank == root else "", root)
        finally:
            self.timing = {"load_s": time.perf_counter() - t0, "broadcast": self.r.dist_broadcast_timing()}

    def begin(self):
        self.local_passes = 0
        if self.world_size > 1:
            self.r.dist_clear()
        else:
            self.r.clear()

    def render(self, total_passes):
        """the next total_passes passes of the image, sharded over the ranks (same argument on every rank)"""
        if self.world_size > 1:
            self.r.dist_render(total_passes)
        else:
            self.r.render(total_passes)

    def render_local(self, n_local_passes):
        self.render(n_local_passes * self.world_size)
        self.local_passes += n_local_passes

    def reduce(self, root=0):
        if self.world_size > 1:
            self.r.dist_reduce(root)

    def image_f32(self):
        """(sum image, passes) of the whole job — valid on the root after reduce()"""
        if self.world_size > 1:
            return self.r.merged_image_f32()
        return self.r.image_f32(), self.r.pass_counter()

    def image_u8(self, out=None):
        if self.world_size > 1:
            return self.r.merged_image_u8(out)
        return self.r.image_u8(out)

    def close(self):
        if self.world_size > 1:
            self.r.dist_shutdown()


class ShardedRenderer:
    """Drives one backend per process; `dist` is torch.distributed (already initialised) or None."""

    def __init__(self, backend, rank=0, world_size=1, dist=None):
        self.backend = backend
        self.rank = rank
        self.world_size = world_size
        self.dist = dist
        self.local_passes = 0
        self._reduced = False

    def begin(self):
        self.backend.clear()
        self.local_passes = 0
        self._reduced = False

    def render_local(self, n_local_passes):
        """Renders this rank's next n_local_passes passes (global indices rank+1 + k*world)."""
        if self._reduced:
            raise RuntimeError("ShardedRenderer: reduce() summed IN PLACE into rank 0's accumulation buffer; call begin() before rendering more "
                               "(DistRenderer / MultiRenderer reduce into a separate merged image and may continue)")
        first = self.rank + 1 + self.local_passes * self.world_size
        self.backend.render_strided(first, self.world_size, n_local_passes)
        self.local_passes += n_local_passes

    def render_total(self, total_passes):
        first, stride, count = shard_passes(total_passes, self.rank, self.world_size)
        remaining = count - self.local_passes
        if remaining > 0:
            self.render_local(remaining)

    def reduce(self, total_passes=None):
        """Sum the per-rank accumulation buffers onto rank 0 and finish the image there."""
        if total_passes is None:
            total_passes = self.local_passes * self.world_size
        if self._reduced:
            raise RuntimeError("ShardedRenderer: reduce() is one-shot per begin() — a second in-place reduce would count the other ranks' passes twice")
        self._reduced = True
        if self.dist is not None and self.world_size > 1:
            t = self.backend.accumulation_tensor()
            self.backend.synchronize()
            self.dist.reduce(t, dst=0, op=self.dist.ReduceOp.SUM)
            after = getattr(self.backend, "after_collective", None)
            if after is not None:
                after()            # the reduced sum must be in place before finalize / clear touch the buffer on another stream
        if self.rank == 0:
            self.backend.finalize(total_passes)
        return total_passes


def merge_reference(images, counts=None):
    """numpy reference of the reduce: float32 sum in rank order (what NCCL's sum produces up to re-association)."""
    acc = np.zeros_like(images[0], dtype=np.float32)
    for im in images:
        acc = acc + im.astype(np.float32)
    return acc
