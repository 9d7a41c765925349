"""World-size-2 gloo test of the multi-GPU path's host logic: each rank steps only its own
shard of environment copies (here with the one-lane host build of the kernel source) and
the gathered results equal a single process stepping the whole batch -- the batch shards
with no data-path collective."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "hostsim"))
sys.path.insert(0, HERE)


def _worker(rank, world, port, blobs, env_instance, actions, rnd, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import build as hostsim
    from deep_reinforcement_learning_for_fjsp_b200 import sharding
    lo, hi = sharding.shard_range(len(env_instance), rank, world)
    vec = hostsim.HostSimVec(blobs, env_instance[lo:hi], "MO_DFJSP")
    vec.reset()
    st, rw, dn, rec = vec.step(actions[:, lo:hi], rnd[:, lo:hi], 1)
    info = vec.info()
    stats = sharding.gather_episode_stats(info["completion"], info["delay_sum"], info["energy"])
    ms, steps = sharding.reduce_timing(10.0 + rank, rec.shape[0] * rec.shape[1])
    # the agents' data-parallel update: every rank's gradients averaged in one flat bucket
    import torch
    net = torch.nn.Linear(3, 2)
    with torch.no_grad():
        for p_ in net.parameters():
            p_.fill_(0.5)
    net(torch.full((4, 3), float(rank + 1))).sum().backward()
    n_el = sharding.allreduce_gradients(net.parameters())
    grads = torch.cat([p_.grad.reshape(-1) for p_ in net.parameters()]).numpy()
    if rank == 0:
        q.put((stats, ms, steps, n_el, grads))
    q.put(("rec", lo, hi, rec))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    import build as hostsim
    import parity_common as pc
    from deep_reinforcement_learning_for_fjsp_b200 import sharding
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    insts, env_instance = pc.random_batch(61, "MO_DFJSP", 3, 3)
    env_instance = env_instance[:7]            # uneven split: 4 + 3
    blobs = [i.to_blob() for i in insts]
    B, T = len(env_instance), 24
    rng = np.random.default_rng(0)
    actions = np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32)
    rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
    whole = hostsim.HostSimVec(blobs, env_instance, "MO_DFJSP")
    whole.reset()
    _, _, _, rec_all = whole.step(actions, rnd, 1)
    info_all = whole.info()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blobs, env_instance, actions, rnd, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec = np.zeros_like(rec_all)
    for item in got:
        if isinstance(item[0], str):
            _, lo, hi, part = item
            rec[:, lo:hi] = part
        else:
            stats, ms, steps, n_el, grads = item
    assert np.array_equal(rec, rec_all)
    assert ms == 11.0 and steps == T * B
    # d(sum of outputs)/dW = 4 * x per output row, averaged over the ranks' x = 1 and x = 2; bias gradient 4
    assert n_el == 8 and np.allclose(grads, [6.0] * 6 + [4.0] * 2)
    assert np.array_equal(stats[:, 0], info_all["completion"]) and np.array_equal(stats[:, 2], info_all["energy"])
