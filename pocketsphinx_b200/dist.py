"""Multi-GPU plumbing for the hot path: one process per GPU (torch.distributed), utterances
sharded across ranks, the acoustic model broadcast once at init.  No per-frame collective exists
on this path, so this module is all there is to the N>1 story.

Works with the `nccl` backend on GPUs (bench.py) and with `gloo` on CPU tensors
(tests/test_dist_gloo.py exercises the same code with world_size 2).
"""
import numpy as np

MODEL_BUFFERS = ("mean", "var", "det", "mixw", "mixw_cb", "sen2cb", "logadd8", "logadd_ms", "topn_beam")


def shard_utterances(n_utt, rank, world):
    """Static assignment utterance -> rank = utt_id mod world.  Returns the ids."""
    return np.arange(rank, n_utt, world, dtype=np.int64)


def shard_by_length(lengths, rank, world):
    """Length-balanced static assignment: longest-first greedy bins; deterministic on every rank."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(len(lengths), np.int64)
    for u in order:
        r = int(np.argmin(load))
        owner[u] = r
        load[r] += int(lengths[u])
    return np.flatnonzero(owner == rank)


def broadcast_model(pm, src=0, device=None):
    """Rank `src` holds the PackedModel; every rank returns a dict name -> torch tensor (on `device`)
    with identical contents, after one torch.distributed.broadcast per packed buffer.  Ranks other
    than `src` only need the model's *shapes* (pm may hold zeros)."""
    import torch
    import torch.distributed as dist
    out = {}
    for k in MODEL_BUFFERS:
        a = getattr(pm, k)
        if a.size == 0:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a))
        if device is not None:
            t = t.to(device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if dist.get_rank() != src:
                t = torch.empty_like(t)
            dist.broadcast(t, src)
        out[k] = t
    return out


def reduce_max_ms(ms_values, device=None):
    """Max over ranks of per-rank timings (the contract's 'time as the max over ranks')."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(ms_values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_counts(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t]
