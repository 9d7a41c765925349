#!/usr/bin/env python
"""Diagnostic: runs a bench config's large batch, finds environments with a non-zero error flag and replays
them on the CPU oracle with the same inputs.  usage (GPU box): python tools/debug_errors.py [--config mo_4096 --envs 65536]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="mo_4096")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--launches", type=int, default=40)
    ap.add_argument("--compare", type=int, default=8, help="also compare this many envs' final state / clock with the oracle")
    a = ap.parse_args()
    import torch
    import oracle_py
    import bench
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    cfg = dict(bench.CONFIGS[a.config]); cfg["envs"] = min(cfg["envs"], 4096) if not cfg.get("brandimarte") else cfg["envs"]
    blobs, _ = bench.config_blobs(cfg, 2026, 0)
    variant, T = cfg["variant"], cfg["T"]
    B = a.envs
    env_inst = (np.arange(B) % len(blobs)).astype(np.int32)
    vec = FJSPVecEnv(None, env_inst, variant, device=0, blobs=blobs)
    vec.reset()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    hist = []
    watch = np.unique(np.linspace(0, B - 1, a.compare).astype(np.int64))
    first_err = {}
    for l in range(a.launches):
        x, r = bench.make_actions(rng, T, B, variant)
        hist.append((x, r))
        o = vec.rollout(torch.from_numpy(x).to(dev), torch.from_numpy(r.view(np.int32)).to(dev), reward_policy=1, state_dtype=torch.float32)
        inf = vec.info()
        bad = np.nonzero(inf["error"])[0]
        for e in bad:
            if int(e) not in first_err:
                first_err[int(e)] = (l, int(inf["error"][e]))
        if len(first_err) >= 4:
            break
    print("launches", len(hist), "envs with errors:", len(first_err), list(first_err.items())[:8])
    inf = vec.info()
    todo = list(first_err)[:3] + [int(w) for w in watch]
    for e in todo:
        env = oracle_py.OracleEnv(blobs[env_inst[e]], variant)
        env.reset()
        err = None
        for l, (x, r) in enumerate(hist):
            try:
                oracle_py.batch_rollout([env], x[:, e:e + 1], r[:, e:e + 1], 1, want_state=False, want_rec=False)
            except AssertionError as ex:
                err = (l, str(ex)); break
        oi = env.info()
        print("env", e, "inst", env_inst[e], "gpu: err", int(inf["error"][e]), "time", int(inf["step_time"][e]), "steps", int(inf["step_count"][e]),
              "lp", int(inf["lp_solves"][e]), int(inf["lp_iterations"][e]), "| oracle:", err, "time", oi["step_time"], "steps", oi["step_count"], "lp", oi["lp_solves"], oi["lp_iters"], "err", oi["error"])


if __name__ == "__main__":
    main()
