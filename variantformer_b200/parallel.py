"""Multi-GPU plumbing: one process per GPU, (sample, gene) items sharded with no data-path collective; the only
collective is the final gather of expression / embedding tensors (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The reference is single-GPU (`Trainer(devices=1)`, processors/vcfprocessor.py:252-258): this layer is new."""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*) -> (rank, world, local_rank).  No-op for world 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def item_cost(n_cre_tokens, n_gene_tokens, C, G, T):
    """Relative cost model of one (sample, gene) item: seq2reg tokens + CRE stream + T gene streams."""
    return (5.24e6 * 6 * (n_cre_tokens + n_gene_tokens) + 24 * (2 * C * 18.9e6 + 4 * C * C * 1536)
            + 25 * T * (2 * (G + 1) * 18.9e6 + 4 * (G + 1) * (G + 1 + C) * 1536) + 25 * 2 * C * 4.7e6)


def shard_items(costs, world):
    """Longest-processing-time-first partition -> list (per rank) of item indices, deterministic on every rank."""
    costs = np.asarray(costs, np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world); parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i)); load[r] += costs[i]
    return [sorted(p) for p in parts]


def gather_rows(local: torch.Tensor, counts, world, rank):
    """All ranks contribute `local` [n_r, ...] (n_r = counts[rank]); every rank gets the concatenation in rank
    order.  Padded all_gather (fixed-size slabs) so a single collective moves everything."""
    if world == 1:
        return local
    m = int(max(counts))
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad) if local.is_cuda else dist.all_gather(list(out.chunk(world)), pad)
    return torch.cat([out[r * m: r * m + int(counts[r])] for r in range(world)])


def scatter_to_query_order(gathered: torch.Tensor, parts, rows_per_item):
    """Undo the sharding: gathered rows are ordered (rank, item-in-rank); return them in original item order."""
    rows_per_item = np.asarray(rows_per_item)
    order = np.concatenate([np.asarray(p, np.int64) for p in parts]) if parts else np.zeros(0, np.int64)
    starts = np.concatenate([[0], np.cumsum(rows_per_item[order])])
    dest_start = np.concatenate([[0], np.cumsum(rows_per_item)])
    idx = np.empty(int(rows_per_item.sum()), np.int64)
    for k, item in enumerate(order):
        idx[dest_start[item]: dest_start[item + 1]] = np.arange(starts[k], starts[k + 1])
    return gathered[torch.from_numpy(idx).to(gathered.device)]
