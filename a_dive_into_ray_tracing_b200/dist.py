"""Multi-GPU plumbing: one process per GPU (torch.distributed), sample-space split.

The path shards by SAMPLES: rank g of G renders the global sample
indices [g*spp/G, (g+1)*spp/G) of every pixel into its own float4 accumulation frame;
Philox counters are (pixel, global sample, event), so the reduced image does not depend
on G up to fp32 summation order. The one real exchange step is a sum-reduce of the
frames to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def sample_range(spp, rank, world):
    """Contiguous, exhaustive, non-overlapping split of [0, spp) over `world` ranks."""
    begin = (spp * rank) // world
    end = (spp * (rank + 1)) // world
    return begin, end - begin


def row_range(H, rank, world, align=4):
    """Contiguous bands of rows (multiples of the 8x4 tile height) for the image-space split."""
    bands = (H + align - 1) // align
    b0, b1 = (bands * rank) // world, (bands * (rank + 1)) // world
    return min(H, b0 * align), min(H, b1 * align)


def reduce_frames(accum, dst=0):
    """Sum the per-rank accumulation frames onto `dst` (in place on dst) with torch.distributed: the
    host-logic path of the gloo CPU tests. On GPUs the frames go through the library's own rt_reduce
    (init_comm + render_frame below)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


def init_comm(ctx, rank=None, world=None):
    """Give the context its NCCL communicator (C ABI rt_comm_init): rank 0 creates the unique id,
    torch.distributed ships the 128 bytes - plumbing only, the reduce itself is rt_reduce."""
    from . import capi
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    box = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
    ctx._has_comm = True


def render_frame(ctx, W, H, spp, accum, rank=0, world=1, stream=None, split="samples"):
    """Render this rank's share of the frame into the CUDA tensor `accum` ([H, W, 4] float32,
    zeroed by the caller) on the current torch stream, then reduce to rank 0. No host
    synchronisation. split = "samples": every pixel, a share of the sample indices (the default:
    perfectly balanced); "rows": all samples of a band of rows (latency of low-spp frames; the
    bands' cost depends on the image content)."""
    assert accum.is_cuda and accum.dtype == torch.float32 and accum.is_contiguous()
    assert tuple(accum.shape) == (H, W, 4)
    s = stream if stream is not None else torch.cuda.current_stream()
    if split == "rows":
        y0, y1 = row_range(H, rank, world)
        ctx.render_rows_device(W, H, y0, y1, spp, 0, accum.data_ptr(), s.cuda_stream)
    else:
        begin, count = sample_range(spp, rank, world)
        ctx.render_device(W, H, count, begin, accum.data_ptr(), s.cuda_stream)
    if world > 1:
        if getattr(ctx, "_has_comm", False):  # the library's own reduce: R,G,B (+ one count word) for the sample split
            ctx.reduce(W, H, accum.data_ptr(), root=0, uniform_count=(split != "rows"), stream_ptr=s.cuda_stream)
        else:
            if stream is not None:  # torch's collective is ordered against the CURRENT stream only
                torch.cuda.current_stream().wait_stream(s)
            reduce_frames(accum)
    return accum
