"""Frame sharding across the GPUs of one box + the single collective of the inference path.

The reference is a single-GPU program (``nn.DataParallel`` only, RCNet/rcnet_model.py:259-265;
SURVEY 2.2).  Frames are independent units at every stage of the hot path, so the
B200 design is one process per GPU (``torchrun``), a contiguous slice of the frames per rank and
NO data-path collective.  The only exchange is the final metric reduction: the reference
averages per-frame metrics with ``np.mean`` over all frames (val_zju.py:247-254); here every rank
contributes ``[sum_0 .. sum_{m-1}, n_frames]`` in float64 to ONE all-reduce (NCCL over
NVLink/NVSwitch on GPUs, gloo on CPU for the tests) and divides.

Nothing in this file touches the CUDA extension, so it is testable on CPU with world_size 2.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT.
    Returns (rank, world_size, local_rank).  A 1-process run never creates a group."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep NCCL's version banner off stdout (bench.py prints one JSON line)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_frames(n_frames, rank, world):
    """Contiguous slice [begin, end) of ``n_frames`` owned by ``rank``: the first
    ``n_frames % world`` ranks take one extra frame (batch 64 -> 8 frames/GPU)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank %d / world %d" % (rank, world))
    base, extra = divmod(int(n_frames), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_points(frame_offsets, begin, end):
    """Ragged companion of ``shard_frames``: the point range of frames [begin, end) and the
    re-based offsets array for the local call."""
    lo, hi = int(frame_offsets[begin]), int(frame_offsets[end])
    return lo, hi, [int(o) - lo for o in frame_offsets[begin:end + 1]]


def barrier():
    if dist.is_initialized():
        if torch.cuda.is_available() and dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()


def _device_for_collectives():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_metric_sums(sums, n_frames):
    """The final metric reduction.  ``sums``: per-rank sums of per-frame metrics (sequence or
    1-D tensor); returns (global means float64 tensor on the CPU, global frame count).
    Equals ``np.mean`` over all frames of all ranks (val_zju.py:247-254)."""
    vec = torch.cat([torch.as_tensor(sums, dtype=torch.float64).reshape(-1).cpu(),
                     torch.tensor([float(n_frames)], dtype=torch.float64)])
    if dist.is_initialized():
        vec = vec.to(_device_for_collectives())
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        vec = vec.cpu()
    total = vec[-1].item()
    means = vec[:-1] / total if total > 0 else torch.full_like(vec[:-1], float("nan"))
    return means, int(round(total))


def all_reduce_sums_device(vec):
    """In-place SUM all-reduce of a small device (or CPU) tensor on the current stream, no host synchronisation: the per-step
    form of ``reduce_metric_sums`` (the caller appends the frame count as the last element and divides when it finally
    reads the vector).  A 1-process run returns ``vec`` untouched."""
    if dist.is_initialized():
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def max_over_ranks(value):
    """Max of a python float over ranks (device timings are reported as the slowest rank's)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_gradients(parameters, bucket_bytes=32 << 20):
    """Config 4 (RC-Net training, rcnet_main.py:342-358): average gradients across ranks in
    flat buckets.  On NVSwitch the cost is launch latency, not link count, so buckets are large
    (the whole 23.6 MB fp32 gradient fits one).  BatchNorm statistics stay per replica like the
    reference (no SyncBN)."""
    if not dist.is_initialized():
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in parameters if p.grad is not None]
    n_calls, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, n_calls
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        o = 0
        for g in bucket:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
        bucket, size = [], 0
        n_calls += 1

    for g in grads:
        if bucket and (g.dtype != bucket[0].dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
    return n_calls
