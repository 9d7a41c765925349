#!/usr/bin/env python
"""Per-CTA / per-SM time breakdown of the step kernel on the bench workload (diagnostic).

Needs the trace build:  python -c "from deep_reinforcement_learning_for_fjsp_b200 import build; build.build_variant('trace', ['-DFJ_TRACE'])"
Run on the GPU box:     FJSP_B200_LIB=deep_reinforcement_learning_for_fjsp_b200/libfjsp_b200_trace.so python tools/cta_trace.py
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--rollout", type=int, default=32)
    ap.add_argument("--launches", type=int, default=8)
    ap.add_argument("--burnin", type=int, default=2048)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from bench import make_instances, make_actions
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    from deep_reinforcement_learning_for_fjsp_b200 import _lib
    B, T = a.envs, a.rollout
    insts = make_instances(B, 2026, 10, 3)
    vec = FJSPVecEnv(insts, np.arange(B), "MO_DFJSP", device=0)
    vec.reset()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    sets = []
    for _ in range(4):
        x, r = make_actions(rng, T, B, "MO_DFJSP")
        sets.append((torch.from_numpy(x).to(dev), torch.from_numpy(r.view(np.int32)).to(dev)))
    out = {"state": torch.empty((T, B, vec.state_size), dtype=torch.float32, device=dev),
           "reward": torch.empty((T, B), dtype=torch.float64, device=dev),
           "done": torch.empty((T, B), dtype=torch.int32, device=dev)}
    for i in range(a.burnin // T):
        vec.rollout(sets[i % 4][0], sets[(i + 1) % 4][1], reward_policy=1, out=out, state_dtype=torch.float32)
    q = vec.query()
    srv = q["lp_server_ctas"]
    grid = q["grid"] - srv          # env CTAs (the grid's first CTAs are the LP servers)
    inf = vec.info()
    ops = np.array([x.total_operations for x in insts]); rub = np.array([x.machine_count + 2 * sum(x.ntask) - x.kind_count for x in insts])
    likely = (inf["done"] != 0) | (inf["next_order"] < 3) | (ops - inf["step_count"] <= T - 8)
    print("predicted LP envs: %d of %d (done %d, young %d, ending %d); heavy %d medium %d light %d" % (
        likely.sum(), B, (inf["done"] != 0).sum(), (inf["next_order"] < 3).sum(), (ops - inf["step_count"] <= T - 8).sum(),
        (likely & (rub >= 85)).sum(), (likely & (rub >= 55) & (rub < 85)).sum(), (likely & (rub < 55)).sum()))
    tr_all = np.zeros((grid + srv, 33, 8), dtype=np.int64)
    tr = tr_all[srv:]
    L = _lib.load()
    _lib.check(L.fjsp_vec_trace(vec._h, tr_all.ctypes.data, 1))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = np.zeros_like(tr)
    acc_lp = np.zeros(8, dtype=np.int64)
    ms = 0.0
    for i in range(a.launches):
        ev0.record()
        vec.rollout(sets[i % 4][0], sets[(i + 1) % 4][1], reward_policy=1, out=out, state_dtype=torch.float32)
        ev1.record()
        torch.cuda.synchronize()
        ms_i = ev0.elapsed_time(ev1)
        ms += ms_i / a.launches
        _lib.check(L.fjsp_vec_trace(vec._h, tr_all.ctypes.data, 1))
        nwi = q["env_warps"]
        ct = tr[:, :nwi, 0].max(1)
        nl = tr[:, :nwi, 5].sum(1)
        ns = L.fjsp_vec_slots(vec._h, None, 0)
        slm = np.zeros(ns, dtype=np.int32)
        L.fjsp_vec_slots(vec._h, slm.ctypes.data, ns)
        act = (slm.reshape(-1, nwi)[:grid] >= 0).sum(1)
        top = np.argsort(-ct)[:4]
        print("launch %d: %.3f ms = %.2f M cycles @1.965 GHz | median CTA %.2f M, CTAs with LPs %d, LPs %d | slowest CTAs: %s" % (
            i, ms_i, ms_i * 1.965, np.median(ct) / 1e6, (nl > 0).sum(), nl.sum(),
            "; ".join("#%d %.2f M (%d envs, %d LPs, longest LP wait %.2f M, busy %.2f M)" % (b, ct[b] / 1e6, act[b], nl[b], tr[b, :nwi, 3].max() / 1e6,
                                                                  (tr[b, :nwi, 1] + tr[b, :nwi, 2] + tr[b, :nwi, 4]).max() / 1e6) for b in top)))
        ones = np.where(act == 1)[0]
        if len(ones):
            print("   single-env CTAs: %d, total mean %.2f M max %.2f M, LP mean %.2f M, LPs mean %.1f" % (
                len(ones), ct[ones].mean() / 1e6, ct[ones].max() / 1e6, tr[ones, :nwi, 3].max(1).mean() / 1e6, nl[ones].mean()))
        acc += tr
        acc_lp += tr_all[:, 32, :].sum(0)
        if i == 0:
            ns = L.fjsp_vec_slots(vec._h, None, 0)
            sl = np.zeros(ns, dtype=np.int32)
            L.fjsp_vec_slots(vec._h, sl.ctypes.data, ns)
            sl = sl.reshape(-1, nwi)
            occ = (sl >= 0).sum(1)
            print("   slot map: %d virtual CTAs x %d; envs placed %d (distinct %d); occupancy histogram %s; first CTAs %s" % (
                sl.shape[0], nwi, (sl >= 0).sum(), len(set(sl[sl >= 0].tolist())), np.bincount(occ, minlength=nwi + 1).tolist(), occ[:24].tolist()))
    tr = acc
    tr = tr.astype(np.float64) / a.launches
    nw = q["env_warps"]
    if True:
        lpt = acc_lp.astype(np.float64)
        solves = tr[:, :nw, 5].sum() * a.launches
        names = ["setup", "pricing+argmin", "wait A (helpers: previous rank-1)", "w+ratio", "argmin-out", "xB/pivot-row/y", "barrier B"]
        print("LP phases (cycles per iteration): " + ", ".join("%s %.0f" % (n, lpt[k] / max(lpt[7], 1)) for k, n in enumerate(names) if k)
              + "; setup per LP %.0f; iterations per LP %.1f; LPs %d" % (lpt[0] / max(solves, 1), lpt[7] / max(solves, 1), solves))
    tr = tr[:, :nw]
    tot, front, clock, lp, back, nlp = (tr[:, :, k] for k in range(6))
    sm = tr[:, 0, 7] * a.launches
    cta_tot = tot.max(1)
    busy = front + clock + lp + back
    print(f"launch {ms:.3f} ms = {ms * 1.9e6:.0f} cycles @1.9GHz; grid {grid} x {nw} warps")
    print(f"CTA total cycles: min {cta_tot.min():.0f} mean {cta_tot.mean():.0f} max {cta_tot.max():.0f}")
    print("per-warp mean cycles: total %.0f busy %.0f (front %.0f clock %.0f lp %.0f back %.0f) barrier-wait %.0f" % (
        tot.mean(), busy.mean(), front.mean(), clock.mean(), lp.mean(), back.mean(), (tot - busy).mean()))
    print("LPs per CTA per launch: mean %.2f max %.2f; cycles its slowest warp waited for LPs: mean %.0f max %.0f" % (
        nlp.sum(1).mean(), nlp.sum(1).max(), lp.max(1).mean(), lp.max(1).max()))
    # deciles of CTAs by index (CTA index = rank in decreasing static walk length)
    for lo in range(0, grid, max(1, grid // 10)):
        hi = min(grid, lo + max(1, grid // 10))
        s = slice(lo, hi)
        print("CTA %3d-%3d: total %.0f | max-warp busy %.0f mean-warp busy %.0f | front %.0f clock %.0f lp %.0f back %.0f" % (
            lo, hi - 1, cta_tot[s].mean(), busy[s].max(1).mean(), busy[s].mean(), front[s].mean(), clock[s].mean(),
            lp[s].mean(), back[s].mean()))
    # per SM: sum of its CTAs' totals vs the longest
    sms = {}
    for b in range(grid):
        sms.setdefault(int(sm[b]), []).append(cta_tot[b])
    per_sm = np.array([max(v) for v in sms.values()])
    print(f"SMs used {len(sms)}; per-SM busy-until: min {per_sm.min():.0f} mean {per_sm.mean():.0f} max {per_sm.max():.0f}")
    if a.out:
        np.save(a.out, tr)


if __name__ == "__main__":
    main()
