"""Multi-GPU host logic: one process per GPU, scene replicated, samples sharded, accumulators summed on rank 0.

The path shards by sample index: with `world` ranks and `spp` samples per rank and frame, rank r renders
sample indices [r*spp, (r+1)*spp) of a (world*spp)-sample frame - the core seeds those samples exactly as a single
GPU rendering all world*spp samples would (RenderParams.sampleBase). The only exchange is one reduce of the float4
accumulator (w*h*16 bytes per rank per frame) to rank 0, which then finalizes (divides by the total sample count).
No collective is needed for the scene: every rank issues the same Set* calls.

`core` below is anything with the small surface used here (RenderCore on a GPU; tests drive the same code over gloo
with a CPU stand-in backed by the oracle)."""
import torch
import torch.distributed as dist


def sample_shard(rank, world, spp):
    """(firstSample, totalSpp) for lh2b_set_sample_shard."""
    return rank * spp, world * spp


class _CudaArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


def accumulator_tensor(core, device):
    """Zero-copy torch view of the core's device accumulator (float32 [h, w, 4])."""
    ptr, _ = core.AccumulatorDevicePtr()
    return torch.as_tensor(_CudaArray(ptr, (core.height, core.width, 4)), device=device)


class ShardedRenderer:
    """Drives one core per rank; rank 0 ends every frame with the finalized full-sample image."""

    def __init__(self, core, spp, rank=None, world=None, accumulator=None):
        self.core = core
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.spp = spp
        first, total = sample_shard(self.rank, self.world, spp)
        core.SetSampleShard(first, total)
        self.acc = accumulator            # tensor aliasing the core's accumulator
        self._sum = None

    def render(self, view, converge=1):
        """One frame on every rank + the reduce. Returns the samples accumulated so far (all ranks)."""
        self.core.Render(view, converge)
        if self.world > 1:
            # reduce into a scratch copy: the cores keep accumulating their own shard when converging
            if self._sum is None:
                self._sum = torch.empty_like(self.acc)
            self._sum.copy_(self.acc)
            dist.reduce(self._sum, dst=0, op=dist.ReduceOp.SUM)
            if self._sum.is_cuda:
                torch.cuda.current_stream().synchronize()
        else:
            self._sum = self.acc
        return self.core.SamplesTaken()

    def finalize(self):
        """Rank 0: pixels = summed accumulator / total samples (lh2b_finalize_external). The core counts samples of
        the whole sharded frame (all ranks), and follows the reference's restart rules (rendercore.cpp:827-833)."""
        total = self.core.SamplesTaken()
        if self.rank == 0:
            self.core.FinalizeExternal(self._sum, total)
        return total


class PipelinedShardedRenderer(ShardedRenderer):
    """The same sharding with frames in flight (GPU only): the core renders frame k+1 (Setting "pipeline") while frame k's
    accumulator snapshot is reduced over NCCL, finalized on rank 0 and copied to pinned host memory on a second stream.

        frame(view, converge, host_out)   enqueue one frame everywhere; rank 0's image lands in host_out asynchronously
        finish()                          wait for everything in flight
    """

    def __init__(self, core, spp, rank=None, world=None, device=None):
        super().__init__(core, spp, rank, world, None)
        core.Setting("pipeline", 1)
        self.device = device
        self._core_stream = torch.cuda.ExternalStream(core.Stream(), device=device)
        self._comm = torch.cuda.Stream(device=device)
        shape = (core.height, core.width, 4)
        self._sums = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(2)]
        self._img = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(2)] if self.rank == 0 else None
        self._done = [None, None]
        self._k = 0

    def frame(self, view, converge=1, host_out=None):
        slot = self._k & 1
        if self._done[slot] is not None:
            self._done[slot].synchronize()          # the buffers of frame k-2 are free again
        self.core.Render(view, converge, True)       # frame k enqueued, frame k-1 harvested
        total = self.core.SamplesTaken()
        buf = self._sums[slot]
        self.core.SnapshotAccumulator(buf)           # behind frame k on the core's stream, in front of frame k+1
        ready = self._core_stream.record_event()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            if self.world > 1:
                dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
            if self.rank == 0:
                self.core.FinalizeExternalOn(buf, total, self._img[slot], self._comm.cuda_stream)
                if host_out is not None:
                    host_out.copy_(self._img[slot], non_blocking=True)
            self._done[slot] = self._comm.record_event()
        self._k += 1
        return total

    def finish(self):
        self.core.WaitForRender()
        self._comm.synchronize()

    def last_event(self):
        """Event after the newest frame's reduce / finalize / read-back (for device-side timing)."""
        return self._done[(self._k - 1) & 1]


import os as _os
_SKIP_GATHER = _os.environ.get("LH2B_DEBUG_SKIP_GATHER") == "1"   # timing experiments only


class PeerGatherRenderer:
    """Sample-sharded frames with the core's own collective (csrc/gather.cu): peers push their accumulator snapshot into
    rank 0's memory with the copy engines over NVLink, rank 0 sums and finalizes in one kernel; hand-shakes are stream memory
    operations, so neither SMs nor the host wait. torch.distributed only carries the CUDA IPC handles at set-up.

        frame(view, converge, host_out)   enqueue one frame on this rank (rank 0: the image goes to pinned host_out, or None)
        finish()                          wait for everything this rank has in flight
    """

    def __init__(self, core, spp, rank=None, world=None):
        self.core = core
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        first, total = sample_shard(self.rank, self.world, spp)
        core.SetSampleShard(first, total)
        core.Setting("pipeline", 1)
        self.g = core.GatherCreate(self.rank, self.world)
        handles = [None] * self.world
        dist.all_gather_object(handles, core.GatherExport(self.g))
        core.GatherImport(self.g, b"".join(handles))
        dist.barrier()

    def frame(self, view, converge=1, host_out=None):
        self.core.Render(view, converge, True)
        total = self.core.SamplesTaken()
        if not _SKIP_GATHER:
            self.core.GatherFrame(self.g, total, host_out if self.rank == 0 else None)
        return total

    def finish(self):
        self.core.WaitForRender()
        self.core.GatherWait(self.g)

    def join(self, stream_handle):
        self.core.GatherJoin(self.g, stream_handle)

    def close(self):
        self.finish()
        dist.barrier()
        self.core.GatherDestroy(self.g)


class TileShardedRenderer:
    """Strong scaling of ONE frame (real-time mode, BASELINE.json configs[4]): rank r path-traces a band of rows, the bands are
    gathered on rank 0 over NVLink peer memory and rank 0 runs the frame's tail (SVGF / TAA chain or finalize) on the complete
    buffers (csrc/tile_gather.cu). Create after SetTarget and Setting("filter"). At 1 spp the frame equals the single-GPU frame
    bit for bit. filter_shard=1 (filter mode, up to 8 ranks): every rank also filters a band of the frame - history lookups go to the
    owning rank over NVLink from inside the filter kernels - and rank 0 only collects the presented bands; interleave picks the
    rendered rows of a rank (1: interleaved 4-row tile rows, 0: its filter band).

        frame(view, converge, host_out)   enqueue one frame on this rank (rank 0: the image goes to pinned host_out, or None)
        finish()                          wait for everything this rank has in flight
    """

    def __init__(self, core, rank=None, world=None, filter_shard=None, interleave=None):
        self.core = core
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        core.Setting("pipeline", 1)
        if filter_shard is not None:
            core.Setting("tileFilterShard", int(filter_shard))
        if interleave is not None:
            core.Setting("tileInterleave", int(interleave))
        self.g = core.TileCreate(self.rank, self.world)
        handles = [None] * self.world
        dist.all_gather_object(handles, core.TileExport(self.g))
        core.TileImport(self.g, b"".join(handles))
        self.rows = core.TileRows(self.g)
        dist.barrier()

    def frame(self, view, converge=1, host_out=None):
        self.core.Render(view, converge, True)
        self.core.TileFrame(self.g)
        if self.rank == 0 and host_out is not None:
            self.core.ReadPixelsAsync(host_out.numpy() if hasattr(host_out, "numpy") else host_out)

    def finish(self):
        self.core.WaitForRender()
        self.core.TileWait(self.g)
        if self.rank == 0:
            self.core.WaitReadPixels()

    def close(self):
        self.finish()
        dist.barrier()
        self.core.TileDestroy(self.g)
