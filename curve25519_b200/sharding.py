"""Multi-GPU plumbing: contiguous index sharding + the one collective of the path.

Every operation of the hot path is independent, so a batch of n operations is split into contiguous
slices [r*ceil(n/G), (r+1)*ceil(n/G)) over the G ranks of one node (one process per GPU), each rank runs the
same kernels on its slice, and the fixed-size result records are combined with ONE all-gather
(NCCL over NVLink on GPUs; gloo in the CPU tests).  SURVEY.md section 8e / north_star.
No reduce, no all-to-all; inputs are either pre-scattered (each rank only holds its slice) or replicated.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`; slices are ceil(n/world) long except the tail."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def padded_rows(n, world):
    return (n + world - 1) // world


def all_gather_records(local, n, group=None):
    """local: [rows_of_this_rank, rec] -> [n, rec] on every rank.  One all_gather_into_tensor; the last
    rank's slice is padded to ceil(n/world) rows so that all contributions have equal size."""
    world = dist.get_world_size(group)
    per = padded_rows(n, world)
    if local.dim() == 1:
        local = local.unsqueeze(1)
        squeeze = True
    else:
        squeeze = False
    if local.shape[0] != per:
        pad = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    out = torch.empty((per * world, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    out = out[:n]
    return out.squeeze(1) if squeeze else out


def sharded_apply(fn, n, inputs, rec_out, out_dtype=torch.uint8, group=None, presharded=False):
    """Run `fn(*slices) -> tensor [rows, rec_out]` on this rank's slice of every input and all-gather.

    inputs: tensors with n rows each (replicated) or, with presharded=True, already this rank's rows."""
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    lo, hi = shard_bounds(n, world, rank)
    mine = inputs if presharded else [x[lo:hi] for x in inputs]
    if hi > lo:
        local = fn(*mine)
    else:
        dev = inputs[0].device
        local = torch.zeros((0, rec_out) if rec_out else (0,), dtype=out_dtype, device=dev)
    return all_gather_records(local, n, group)
