"""Multi-GPU plumbing: the path shards by samples-per-pixel.

Rank g of G renders the full frame for samples [g*N/G, (g+1)*N/G) -- disjoint PCG32 streams, because the
stream id contains the sample index -- and the float4 accumulators (sum of finite samples, count) are
combined by ONE sum all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The mean and the
luminance clamp (main.cpp:168-173) come after the reduction, so the result does not depend on G beyond
float summation order.
"""
import torch
import torch.distributed as dist


def shard_range(n_samples, rank, world):
    """Half-open sample range of `rank`; ranges tile [0, n) exactly, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_samples * rank // world, n_samples * (rank + 1) // world


def allreduce_accumulator(acc):
    """In-place sum of the accumulator across all ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def finalize_torch(acc, max_luminance=1000.0):
    """Reference finalisation on a torch tensor (used by the CPU tests; the GPU path uses mrt_gpu_finalize_device)."""
    cnt = acc[..., 3:4]
    color = torch.where(cnt > 0, acc[..., :3] / cnt, torch.zeros_like(acc[..., :3]))
    c = torch.tensor([0.212655, 0.715158, 0.072187], dtype=acc.dtype, device=acc.device)
    prod = color * c
    lum = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    over = lum > max_luminance
    scale = torch.where(over, max_luminance / lum, torch.ones_like(lum))
    return torch.where(over[..., None], color * scale[..., None], color)


def progressive_schedule(n_samples, rank, world, passes):
    """Sample ranges of `rank` for each of `passes` refinement passes (the reference's
    draw2 / work_queue_dynamic renders sample-major passes and shows a running mean, main.cpp:193-243).

    The rank's own slice [b, e) = shard_range(...) is cut into `passes` consecutive pieces; pass p of every rank
    together covers n_samples/passes samples spread over the whole sample grid, so the running mean after any pass
    is an unbiased preview.  Empty pieces are returned as (x, x)."""
    b, e = shard_range(n_samples, rank, world)
    passes = max(1, int(passes))
    return [(b + (e - b) * p // passes, b + (e - b) * (p + 1) // passes) for p in range(passes)]


class ProgressiveReducer:
    """Per-pass sum all-reduce overlapped with the next pass.

    `render_pass(begin, end, out)` must ADD the samples [begin, end) of this rank to the float4 accumulator `out`
    (sum of finite radiance, finite-sample count) -- on the GPU that is Renderer.render_async(..., accumulate=True)
    into a bound accumulator; in the CPU tests a stand-in.  After every pass the cumulative local accumulator is
    snapshotted into one of two staging buffers and all-reduced asynchronously while the next pass renders; `preview`
    receives (pass_index, reduced_accumulator) once that reduction has finished (one pass late, by design).
    On CUDA the snapshot and the collective run on a side stream ordered after the RENDER stream by an event.  The render
    stream is `stream` (a torch.cuda.Stream; default: the current stream at construction): `render_pass` must launch on it --
    i.e. the renderer's mrt_gpu_set_stream must have been given this stream's handle; pass `renderer=` to have that done here.
    A renderer left on the legacy NULL stream while torch runs on a non-default stream would not be ordered against the
    snapshot copy, so the contract is explicit.
    """

    def __init__(self, acc, render_pass, preview=None, stream=None, renderer=None):
        self.acc = acc
        self.render_pass = render_pass
        self.preview = preview
        self.render_stream = None
        if acc.is_cuda:
            self.render_stream = stream if stream is not None else torch.cuda.current_stream(acc.device)
            if renderer is not None:
                renderer.set_stream(self.render_stream.cuda_stream)
        self.staging = [torch.empty_like(acc), torch.empty_like(acc)]
        self.cuda = acc.is_cuda
        self.side = torch.cuda.Stream(device=acc.device) if self.cuda else None
        self._pending = None   # (pass index, staging buffer, work handle | None, done event | None)
        self._snapshot_taken = None   # CUDA event: the staging copy has read `acc`, the next pass may write it

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _launch_reduce(self, p):
        buf = self.staging[p & 1]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(self.render_stream)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                buf.copy_(self.acc, non_blocking=True)
                self._snapshot_taken = torch.cuda.Event()
                self._snapshot_taken.record(self.side)
                work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True) if self._distributed() else None
                if work is not None:
                    work.wait()          # orders the side stream after the collective, does not block the host
                done = torch.cuda.Event()
                done.record(self.side)
            self._pending = (p, buf, None, done)
        else:
            buf.copy_(self.acc)
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True) if self._distributed() else None
            self._pending = (p, buf, work, None)

    def _finish_pending(self):
        if self._pending is None:
            return None
        p, buf, work, done = self._pending
        if work is not None:
            work.wait()
        if done is not None:
            done.synchronize()
        self._pending = None
        if self.preview is not None:
            self.preview(p, buf)
        return buf

    def run(self, schedule):
        """Render all passes of `schedule` (list of (begin, end)); returns the fully reduced accumulator (for an empty
        schedule: the all-reduced accumulator as it stands)."""
        if not schedule:
            self._launch_reduce(0)
            return self._finish_pending()
        last = None
        for p, (b, e) in enumerate(schedule):
            if e > b:
                if self._snapshot_taken is not None:   # the previous snapshot must have read acc before it changes
                    self.render_stream.wait_event(self._snapshot_taken)
                self.render_pass(b, e, self.acc)      # pass p renders while pass p-1 is being reduced
            last = self._finish_pending() if self._pending is not None else last
            self._launch_reduce(p)
        return self._finish_pending()
