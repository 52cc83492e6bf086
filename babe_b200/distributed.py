"""One process per GPU over ``torch.distributed``: chains and segments are
independent units, so ranks share nothing on the data path.

* independent mode (default): rank r owns rows [r*B/G, (r+1)*B/G) and runs its
  own sampler call; NCCL only gathers the final outputs and the estimated
  filter parameters.
* joint mode: the ranks reproduce ONE reference call on the whole batch, in
  which the filter (utils/blind_bwe_utils.py:295) and the guidance scale
  (testing/blind_bwe_sampler.py:125) couple the rows: the fit statistics
  (3F doubles, once per fit) and the squared gradient norm (1 double per
  guidance call) are all-reduced.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*.
    Returns (rank, local_rank, world_size); a plain single-process run gives
    (0, 0, 1) without creating a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_rows(n_rows, rank, world):
    """Contiguous block partition; the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device):
    """Scalar max over ranks (timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks_(t, group=None):
    """In-place SUM all-reduce (joint-mode statistics); no-op single process."""
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_rows(x, n_rows_total=None):
    """All-gather row blocks of possibly different sizes along dim 0."""
    if not dist.is_initialized():
        return x
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64, device=x.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device))
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = x if x.shape[0] == m else torch.cat((x, x.new_zeros((m - x.shape[0],) + tuple(x.shape[1:]))), 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def gather_params(p):
    """Stack every rank's (2,K) filter estimate -> (world, 2, K)."""
    if not dist.is_initialized():
        return p.unsqueeze(0)
    out = [torch.empty_like(p) for _ in range(dist.get_world_size())]
    dist.all_gather(out, p.contiguous())
    return torch.stack(out, 0)
