#!/usr/bin/env python
"""Latency of one fluid LP on a dedicated SM: times reset() of the bench batch, which solves one order-0
LP per environment copy with fjsp_lp_kernel (one 256-thread CTA per LP, the same fast path the step
kernel's LP servers run).  usage (GPU box): python tools/lp_probe.py [--envs 4096]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--machines", type=int, default=10)
    ap.add_argument("--orders", type=int, default=3)
    ap.add_argument("--profile", default="DA3C")
    a = ap.parse_args()
    import torch
    from bench import make_instances
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    insts = make_instances(a.envs, 2026, a.machines, a.orders, a.profile)
    vec = FJSPVecEnv(insts, np.arange(a.envs), "MO_DFJSP", device=0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for i in range(4):
        torch.cuda.synchronize()
        ev[0].record()
        vec.reset()
        ev[1].record()
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    if os.environ.get('FJSP_B200_LIB', '').endswith('trace.so'):
        from deep_reinforcement_learning_for_fjsp_b200 import _lib
        q = vec.query()
        tr = np.zeros((q['grid'], 33, 8), dtype=np.int64)
        L = _lib.load()
        _lib.check(L.fjsp_vec_trace(vec._h, tr.ctypes.data, 1))
        vec.reset()
        torch.cuda.synchronize()
        _lib.check(L.fjsp_vec_trace(vec._h, tr.ctypes.data, 1))
        lpt = tr[:, 32, :].sum(0).astype(np.float64)
        names = ['setup', 'pricing+argmin', 'wait A', 'w+ratio', 'argmin-out', 'xB/pivot-row/y', 'barrier B']
        print('LP kernel phases (cycles per iteration; NOTE the trace build keeps its counters in local memory, every mark costs ~1.5 k cycles: relative only): ' + ', '.join('%s %.0f' % (n, lpt[k] / max(lpt[7], 1)) for k, n in enumerate(names)) + '; iterations %d' % lpt[7])
    inf = vec.info()
    iters = inf["lp_iterations"].astype(np.float64) / np.maximum(inf["lp_solves"], 1)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    best = min(ms[1:])
    print("reset() of %d envs: %s ms; mean simplex iterations per LP %.1f (max %.0f)" % (a.envs, ["%.3f" % x for x in ms], iters.mean(), iters.max()))
    print("=> %.1f us per LP per SM-resident CTA (one CTA per SM assumed), %.0f cycles per iteration @1.965 GHz" % (
        best * 1e3 / (a.envs / sms), best * 1e-3 * 1.965e9 / (a.envs / sms) / iters.mean()))


if __name__ == "__main__":
    main()
