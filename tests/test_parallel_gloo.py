"""Multi-process path on CPU: gloo, world_size 2 — sharding, padded gather, scatter back to query order."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from variantformer_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = parallel.init_from_env(backend="gloo")
    rows = [3, 1, 2, 5, 1, 4, 2]                           # tissues per item
    costs = [parallel.item_cost(97 * c, 40000, c, 200, t) for c, t in zip([900, 100, 300, 1024, 64, 700, 256], rows)]
    parts = parallel.shard_items(costs, w)
    mine = parts[r]
    # "result" rows of item i are filled with i so the reassembly can be verified
    local = torch.cat([torch.full((rows[i], 4), float(i)) for i in mine]) if mine else torch.zeros(0, 4)
    counts = [sum(rows[i] for i in p) for p in parts]
    gathered = parallel.gather_rows(local, counts, w, r)
    ordered = parallel.scatter_to_query_order(gathered, parts, rows)
    want = torch.cat([torch.full((rows[i], 4), float(i)) for i in range(len(rows))])
    q.put((r, bool(torch.equal(ordered, want)), parts))
    torch.distributed.destroy_process_group()


def test_shard_gather_scatter_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]                          # identical partition on every rank
    parts = res[0][2]
    assert sorted(i for p in parts for i in p) == list(range(7))


def test_lpt_balance():
    rng = np.random.default_rng(1)
    costs = rng.lognormal(0, 0.6, 200)
    for w in (2, 4, 8):
        parts = parallel.shard_items(costs, w)
        load = np.array([costs[p].sum() for p in parts])
        assert load.max() / load.mean() < 1.05
