"""Sample-sharded multi-GPU rendering: the scene is replicated on every rank,
rank r of G renders passes r+1, r+1+G, ... (each pass keeps the seed it has in a single-GPU run,
Kernel/path_tracer_kernel.cu:712), accumulates locally, and ONE sum-reduce of the float
accumulation buffer per image lands on rank 0, which then runs the mean / gamma / 8-bit step.
No data-path collective exists besides that reduce; torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests) is plumbing only.
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


class ShardedRenderer:
    """Drives one backend per process; `dist` is torch.distributed (already initialised) or None."""

    def __init__(self, backend, rank=0, world_size=1, dist=None):
        self.backend = backend
        self.rank = rank
        self.world_size = world_size
        self.dist = dist
        self.local_passes = 0

    def begin(self):
        self.backend.clear()
        self.local_passes = 0

    def render_local(self, n_local_passes):
        """Renders this rank's next n_local_passes passes (global indices rank+1 + k*world)."""
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
