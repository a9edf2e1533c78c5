"""Data-parallel plumbing: one process per GPU, one replay shard per process (nothing in the memory path crosses
GPUs), ONE flat-buffer gradient all-reduce per learn step over NCCL / NVLink.

Replaces the reference's parameter-server scheme (tensorflow_components/architecture.py:485-521: workers add their
gradients to shared accumulators behind a spin barrier, the chief applies them, workers pull the weights): after the
all-reduce every rank holds the same summed gradient and applies the identical optimizer step, so weights stay in
lock-step without a weight broadcast.  ``scale_down`` mirrors
``scale_down_gradients_by_number_of_workers_for_sync_training`` (mean instead of sum; DDPG/TD3 set it False,
ddpg_agent.py:51,69).
"""
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); no-op for single-process runs."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or (dist.is_available() and dist.is_initialized()):
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)
    return world()


def is_distributed() -> bool:
    return world()[1] > 1


def allreduce_gradients(flat_grad: torch.Tensor, scale_down: bool = True) -> float:
    """Sums the flat gradient buffer over all ranks in place and returns the scaler the optimizer step must apply
    (1/world when averaging, else 1)."""
    _, ws = world()
    if ws == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / ws if scale_down else 1.0


def shard_seed(base_seed: int) -> int:
    """Per-rank seed for the rollout / replay shard (seeds 100 + rank)."""
    rank, _ = world()
    return int(base_seed) + rank


def max_over_ranks(value: float, device=None) -> float:
    """max-reduce of a host scalar (timing: a multi-GPU step takes as long as its slowest rank)."""
    _, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_running_stats(delta_sum: torch.Tensor, delta_sumsq: torch.Tensor, rows: int):
    """Shared observation statistics across rollout shards (the reference's SharedRunningStats keeps ONE set of
    (count, sum, sum of squares) that every worker pushes into, utilities/shared_running_stats.py:115-164): the ranks'
    increments of one push are merged with ONE all-reduce of the packed vector [sum | sumsq | rows].  In place on the two
    tensors; returns the total number of rows pushed by all ranks."""
    _, ws = world()
    if ws == 1:
        return int(rows)
    d = delta_sum.numel()
    packed = torch.empty(2 * d + 1, dtype=torch.float64, device=delta_sum.device)
    packed[:d] = delta_sum.reshape(-1)
    packed[d:2 * d] = delta_sumsq.reshape(-1)
    packed[2 * d] = float(rows)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    delta_sum.reshape(-1).copy_(packed[:d])
    delta_sumsq.reshape(-1).copy_(packed[d:2 * d])
    return int(round(float(packed[2 * d].item())))


def global_standardize_(x: torch.Tensor, n_valid: int):
    """(x - mean) / population std over the first n_valid entries of EVERY rank's vector (ClippedPPO advantage
    standardisation, clipped_ppo_agent.py:201, taken over the whole distributed rollout): two small all-reduces
    (count and sum, then the centred sum of squares -- two-pass, no cancellation).  In place; entries >= n_valid become
    NaN like the single-GPU kernel leaves them.  Returns (mean, std) as Python floats."""
    _, ws = world()
    v = x[:n_valid].to(torch.float64)
    a = torch.stack([torch.tensor(float(n_valid), dtype=torch.float64, device=x.device), v.sum()])
    if ws > 1:
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    n_all, mean = float(a[0].item()), float((a[1] / a[0]).item())
    b = ((v - mean) ** 2).sum().reshape(1)
    if ws > 1:
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    std = float(torch.sqrt(b[0] / n_all).item())
    x[:n_valid] = (v - mean) / std
    x[n_valid:] = float("nan")
    return mean, std
