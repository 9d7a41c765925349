#!/usr/bin/env python
"""Where an end-to-end call's time goes (GPU box): the bench workload after its burn-in, alternating phases of
device-timed launches (CUDA events around vec.rollout, inputs resident) and blocking fjsp_vec_step_host calls
(wall clock, page-locked host buffers) by each output route; FJSP_HOST_DEBUG=1 adds the library's own time line
of every host call on stderr.  usage: python tools/e2e_probe.py [--config mo_4096 --calls 30]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="mo_4096")
    ap.add_argument("--calls", type=int, default=30)
    a = ap.parse_args()
    import torch
    import bench
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    from deep_reinforcement_learning_for_fjsp_b200 import _lib
    cfg = bench.CONFIGS[a.config]
    blobs, env_inst = bench.config_blobs(cfg, 2026, 0)
    variant, B, T = cfg["variant"], cfg["envs"], cfg["T"]
    vec = FJSPVecEnv(None, env_inst, variant, device=0, blobs=blobs)
    vec.reset()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(2026)
    NP_ = 8
    pool = [bench.make_actions(rng, T, B, variant) for _ in range(NP_)]
    acts = [torch.from_numpy(x).to(dev) for x, _ in pool]
    rnds = [torch.from_numpy(r.view(np.int32)).to(dev) for _, r in pool]
    ha = [torch.from_numpy(x).pin_memory() for x, _ in pool]
    hr = [torch.from_numpy(r.view(np.int32)).pin_memory() for _, r in pool]
    out = {"state": torch.empty((T, B, vec.state_size), dtype=torch.float32, device=dev),
           "reward": torch.empty((T, B), dtype=torch.float64, device=dev),
           "done": torch.empty((T, B), dtype=torch.int32, device=dev)}
    hs = torch.empty((T, B, vec.state_size), dtype=torch.float32).pin_memory()
    hrw = torch.empty((T, B), dtype=torch.float64).pin_memory()
    hdn = torch.empty((T, B), dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev)
    n = [0]

    def dev_phase(k):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for i in range(k):
            ev[i][0].record(stream)
            vec.rollout(acts[n[0] % NP_], rnds[n[0] % NP_], reward_policy=1, out=out, state_dtype=torch.float32)
            ev[i][1].record(stream)
            n[0] += 1
        torch.cuda.synchronize(dev)
        return [x.elapsed_time(y) for x, y in ev]

    def host_phase(k, zc):
        if zc is None:
            os.environ.pop("FJSP_ZEROCOPY", None)
        else:
            os.environ["FJSP_ZEROCOPY"] = zc
        ts = []
        for i in range(k):
            t0 = time.perf_counter()
            _lib.check(vec._L.fjsp_vec_step_host(vec._h, T, ha[n[0] % NP_].data_ptr(), hr[n[0] % NP_].data_ptr(), 1, 1.0, 1.0, 1.0, 1,
                                                 None, hs.data_ptr(), hrw.data_ptr(), hdn.data_ptr(), None))
            ts.append((time.perf_counter() - t0) * 1e3)
            n[0] += 1
        os.environ.pop("FJSP_ZEROCOPY", None)
        return ts

    dev_phase(cfg["burnin"] // T + 3)
    k = a.calls
    for name, f in (("device", lambda: dev_phase(k)), ("host progressive", lambda: host_phase(k, None)), ("device", lambda: dev_phase(k)),
                    ("host staged-after", lambda: host_phase(k, "0")), ("device", lambda: dev_phase(k)),
                    ("host kernel-stores", lambda: host_phase(k, "1")), ("device", lambda: dev_phase(k))):
        ts = f()
        print("%-20s launches %4d..: mean %.3f ms  min %.3f  max %.3f | %s" % (name, n[0] - k, np.mean(ts), min(ts), max(ts),
                                                                              " ".join("%.2f" % t for t in ts)), flush=True)
    assert (vec.info()["error"] == 0).all()


if __name__ == "__main__":
    main()
