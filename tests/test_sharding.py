"""Instance sharding: contiguous ranges, no collective on the data path; the only
torch.distributed use is the max-over-ranks of the device time.  world_size-2 gloo run on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def test_shard_ranges_partition():
    from decentralized_ekf_mhe_b200.sharding import shard_range
    for n in (1, 7, 65536, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from decentralized_ekf_mhe_b200 import synth
    from decentralized_ekf_mhe_b200.sharding import max_over_ranks, shard_range, sum_over_ranks
    from oracle import pyoracle as po
    n_total = 6
    lo, hi = shard_range(n_total, rank, ws)
    # every rank generates the same global stream and slices its own contiguous instance range:
    st = synth.to_numpy(synth.make_stream(n_total, 40, vo_jitter=True))
    mine = {k: np.ascontiguousarray(v[..., lo:hi]) for k, v in st.items()}
    res, _, _ = po.run_batch(mine, po.go1_params(), po.ekf_params(rate=200), nthreads=1, want=("x",))
    t = max_over_ranks(1.0 + rank)
    cnt = sum_over_ranks(hi - lo)
    dist.barrier()
    q.put((rank, lo, hi, res["x"][-1].copy(), t, cnt))
    dist.destroy_process_group()


def test_sharded_results_equal_unsharded_gloo(oracle):
    from decentralized_ekf_mhe_b200 import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    st = synth.to_numpy(synth.make_stream(6, 40, vo_jitter=True))
    full, _, _ = oracle.run_batch(st, oracle.go1_params(), oracle.ekf_params(rate=200), nthreads=2, want=("x",))
    for rank, lo, hi, x, t, cnt in got:
        assert t == 2.0 and cnt == 6.0
        np.testing.assert_array_equal(x, full["x"][-1][:, lo:hi])  # identical per-instance results regardless of rank count
