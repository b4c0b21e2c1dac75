"""One-process-per-GPU plumbing for the sample-sharded renderer (torch.distributed: NCCL on GPUs, gloo in CPU tests).

The path shards by sample index: rank r of W renders global samples [N*r//W, N*(r+1)//W) into its
own UNSCALED film; the films are summed to rank 0 with ONE reduce and rank 0 rescales by W*H/N. This replaces
contexts.combine_each(film->Accumulate) + Rescale of the reference scheduler (scheduler.cpp:280-288). Because the
random numbers are keyed by the global sample index, the result does not depend on W (up to fp32 summation order)."""


def shard_range(num_samples, rank, world, begin=0):
    """Contiguous slice of [begin, begin+num_samples) owned by `rank`; slices tile the range exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return begin + num_samples * rank // world, begin + num_samples * (rank + 1) // world


def reduce_film(film, dist=None, dst=0):
    """Sum the per-rank films into rank `dst` (in place). film: torch tensor (H,W,4) on the rank's device."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def film_scale(width, height, num_samples):
    """scheduler.cpp:288: Rescale((W*H)/processedSamples)."""
    return float(width * height) / float(num_samples)
