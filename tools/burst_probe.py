#!/usr/bin/env python
"""Launch times right after reset(): every environment copy meets its order arrivals (two fluid LPs each on the bench
workload) within the first few dozen steps, all at once -- the worst case for the LP service.
usage (GPU box): python tools/burst_probe.py [--envs 4096] [--T 32]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--T", type=int, default=32)
    ap.add_argument("--launches", type=int, default=6)
    a = ap.parse_args()
    import torch
    import bench
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    cfg = dict(bench.CONFIGS["mo_4096"]); cfg["envs"] = a.envs
    blobs, env_inst = bench.config_blobs(cfg, 2026, 0)
    vec = FJSPVecEnv(None, env_inst, "MO_DFJSP", device=0, blobs=blobs)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    for rep in range(2):
        vec.reset()
        lp0 = int(vec.info()["lp_solves"].sum())
        ms = []
        for i in range(a.launches):
            x, r = bench.make_actions(rng, a.T, a.envs, "MO_DFJSP")
            xa, ra = torch.from_numpy(x).to(dev), torch.from_numpy(r.view(np.int32)).to(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            vec.rollout(xa, ra, reward_policy=1, state_dtype=torch.float32)
            e1.record()
            torch.cuda.synchronize()
            ms.append(round(e0.elapsed_time(e1), 3))
        inf = vec.info()
        print("after reset(): launch ms %s; fluid LPs solved in these launches %d; errors %d" % (ms, int(inf["lp_solves"].sum()) - lp0, int((inf["error"] != 0).sum())))


if __name__ == "__main__":
    main()
