"""CPU: the N>1 host logic (model broadcast, utterance sharding, max-over-ranks timing) with two
gloo ranks.  The data path has no collective, so this is the whole multi-GPU surface."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from pocketsphinx_b200 import dist as pdist
    from pocketsphinx_b200.model import PackedModel, synth_feats, synth_ptm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pm = synth_ptm(seed=3, n_density=32, n_sen=300)
    if rank != 0:                       # non-source ranks start from zeros of the right shapes
        for k in pdist.MODEL_BUFFERS:
            getattr(pm, k)[...] = 0
    bufs = pdist.broadcast_model(pm, src=0)
    for k, t in bufs.items():
        getattr(pm, k)[...] = t.numpy().reshape(getattr(pm, k).shape)
    # every rank scores its shard of a common batch with the (CPU) oracle; rank 0 checks the union
    feats = synth_feats(synth_ptm(seed=3, n_density=32, n_sen=300), 7, 9, seed=5)
    mine = pdist.shard_utterances(7, rank, world)
    om = oracle.OracleModel(pm)
    digest = {int(u): int(om.score_utt(feats[u]).astype(np.int64).sum()) for u in mine}
    n_frames, = pdist.sum_counts([9 * len(mine)])
    tmax, = pdist.reduce_max_ms([10.0 * (rank + 1)])
    q.put((rank, digest, n_frames, tmax, [int(x) for x in pdist.shard_by_length([5, 9, 1, 7, 7, 3], rank, world)]))
    dist.destroy_process_group()


def test_two_rank_shard_and_broadcast():
    from oracle import oracle
    from pocketsphinx_b200.model import synth_feats, synth_ptm
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pm = synth_ptm(seed=3, n_density=32, n_sen=300)
    feats = synth_feats(pm, 7, 9, seed=5)
    om = oracle.OracleModel(pm)
    want = {u: int(om.score_utt(feats[u]).astype(np.int64).sum()) for u in range(7)}
    got = {}
    for rank, digest, n_frames, tmax, bal in res:
        got.update(digest)
        assert n_frames == 63 and tmax == 20.0
    assert got == want                       # rank 1 scored with the broadcast model, not zeros
    assert sorted(res[0][4] + res[1][4]) == list(range(6))
    assert sorted(res[0][1]) == [0, 2, 4, 6] and sorted(res[1][1]) == [1, 3, 5]
