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
