#!/usr/bin/env python
"""T=1 launches (one reference step() per launch) on the bench workload: per-launch device times and
the LP work they contain.  Diagnostic; also the target of `ncu -k regex:fjsp_step_kernel` captures of the
LP servers (python tools/t1_probe.py --launches 40)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--launches", type=int, default=200)
    ap.add_argument("--burnin", type=int, default=2048)
    ap.add_argument("--T", type=int, default=1)
    a = ap.parse_args()
    import torch
    from bench import make_instances, make_actions
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    B = a.envs
    insts = make_instances(min(B, 4096), 2026, 10, 3)
    vec = FJSPVecEnv(insts, np.arange(B) % len(insts), "MO_DFJSP", device=0)
    vec.reset()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    x, r = make_actions(rng, 32, B, "MO_DFJSP")
    xa, ra = torch.from_numpy(x).to(dev), torch.from_numpy(r.view(np.int32)).to(dev)
    for i in range(a.burnin // 32):
        vec.rollout(xa, ra, reward_policy=1, want_state=False)
    T = a.T
    x, r = make_actions(rng, T * 8, B, "MO_DFJSP")
    xs, rs = torch.from_numpy(x).to(dev), torch.from_numpy(r.view(np.int32)).to(dev)
    out = {"state": torch.empty((T, B, vec.state_size), dtype=torch.float32, device=dev),
           "reward": torch.empty((T, B), dtype=torch.float64, device=dev),
           "done": torch.empty((T, B), dtype=torch.int32, device=dev)}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.launches)]
    lp0 = vec.info()["lp_iterations"].copy()
    its = []
    for i in range(a.launches):
        k = (i % 8) * T
        ev[i][0].record()
        vec.rollout(xs[k:k + T], rs[k:k + T], reward_policy=1, out=out, state_dtype=torch.float32)
        ev[i][1].record()
        if a.launches <= 64:
            lp1 = vec.info()["lp_iterations"]
            its.append(int((lp1 - lp0).max()))
            lp0 = lp1.copy()
    torch.cuda.synchronize()
    ms = np.array([x.elapsed_time(y) for x, y in ev])
    print("T=%d, %d envs: launch ms min %.3f median %.3f mean %.3f max %.3f -> %.2f M env-steps/s" % (
        T, B, ms.min(), np.median(ms), ms.mean(), ms.max(), B * T / ms.mean() / 1e3))
    if its:
        print("per launch (ms, largest LP iterations of one env):", [(round(float(m), 3), n) for m, n in zip(ms, its)])


if __name__ == "__main__":
    main()
