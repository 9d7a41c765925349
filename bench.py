#!/usr/bin/env python
"""FJSP env-steps/sec benchmark (BASELINE.json metric) for the B200 vector environment.

  python bench.py [--config NAME] [--gpus N --steps K --warmup W]   (N > 1: launched under torchrun)
  python bench.py --impl reference [--config NAME] ...              (the CPU port of the reference env)

Configs (BASELINE.json `configs`, one JSON line each; the default is the one the metric is quoted on):
  so_single        [0] SO_DFJSP, one generated instance (10 machines, 3 orders), one step() per launch
  mo_4096          [1] MO_DFJSP, 4096 generated instances per GPU, random composite-rule actions   (default)
  breakdown_65536  [2] MO_DFJSP_breakdown, 65 536 copies per GPU (4096 generated instances with machine
                       breakdown / repair intervals x 16), 32-step rollouts (MPPPO-style)
  brandimarte_1m   [3] SO_DFJSP on Brandimarte mk01-mk10, 131 072 copies per GPU (1 M over 8 GPUs)
  large_m20        [4] MO_DFJSP, 20 machines, 5 orders (HMPSAC generator profile), 4096 instances per GPU

One bench "step" = ONE launch of the step kernel = `T` env steps for every environment copy of the
rank (so_single: 64 launches of one step).  `value` is device-timed (CUDA events on the launching
stream, inputs resident in HBM, L2 flushed between timed launches); `e2e` goes through the host-buffer
C-ABI call with pinned host buffers, H2D of actions/draws and D2H of state/reward/done inside the timed
region.  After the timed region a sample of environments of that very batch is replayed on the CPU
oracle from reset() through every launch and compared (`parity_sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fjsp_env_steps_per_sec"
UNIT = "env_steps/s"

CONFIGS = {
    "so_single": dict(variant="SO_DFJSP", envs=1, T=1, launches_per_step=64, machines=10, orders=3, profile="DA3C",
                      workload="SO_DFJSP_single_generated_instance_M10_S3_one_step_per_launch", large=0, burnin=256),
    "mo_4096": dict(variant="MO_DFJSP", envs=4096, T=32, machines=10, orders=3, profile="DA3C",
                    workload="MO_DFJSP_4096_generated_instances_random_rules", large=65536, burnin=2048),
    "breakdown_65536": dict(variant="MO_DFJSP_breakdown", envs=65536, distinct=4096, T=32, machines=10, orders=3,
                            profile="DA3C", breakdowns=True, large=0, burnin=1024,
                            workload="MO_DFJSP_breakdown_65536_copies_of_4096_generated_instances_32_step_rollouts"),
    "brandimarte_1m": dict(variant="SO_DFJSP", envs=131072, brandimarte=True, T=32, large=0, burnin=512,
                           workload="SO_DFJSP_Brandimarte_mk01_mk10_131072_copies_per_gpu"),
    "large_m20": dict(variant="MO_DFJSP", envs=4096, T=32, machines=20, orders=5, profile="HMPSAC", large=0,
                      burnin=1024, workload="MO_DFJSP_4096_generated_instances_M20_S5_HMPSAC_profile"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mo_4096", choices=sorted(CONFIGS))
    ap.add_argument("--envs", type=int, default=None, help="environment copies per GPU (default: the config's)")
    ap.add_argument("--rollout", type=int, default=None, help="env steps per kernel launch (default: the config's)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--burnin", type=int, default=None,
                    help="untimed env steps per copy before timing, so episodes (order arrivals, resets) desynchronise")
    ap.add_argument("--large-envs", type=int, default=None,
                    help="also time this many copies (the same instances, replicated) for a few launches; 0 = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the T=1 / T=128 rollout-length lines")
    ap.add_argument("--no-policy", dest="policy", action="store_false", help="skip the policy-in-the-loop lines")
    ap.add_argument("--parity-envs", type=int, default=16, help="environments replayed on the CPU oracle after timing")
    ap.add_argument("--e2e-all", action="store_true", help="also time fjsp_vec_step_host with kernel stores into host memory (FJSP_ZEROCOPY=1)")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: do not evict L2 between timed steps (the line says so)")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.envs is not None:
        cfg["envs"] = a.envs
    if a.rollout is not None:
        cfg["T"] = a.rollout
    if a.burnin is not None:
        cfg["burnin"] = a.burnin
    if a.large_envs is not None:
        cfg["large"] = a.large_envs
    cfg.setdefault("launches_per_step", 1)
    return a, cfg


def make_instances(n, seed, M, S, profile="DA3C", breakdowns=False):
    from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance
    return [FJSPInstance.generate(seed * 100003 + i, [0.5, 1.0, 1.5][i % 3], M, S, profile, breakdowns=breakdowns)
            for i in range(n)]


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank's host thread on the CPUs of its GPU's NUMA node, so that the page-locked buffers it
    allocates afterwards (first touch) and the link traffic stay on that node.  Only narrows the CPU set the
    process already has; does nothing (and says why) when the box exposes no such topology."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)                 # torch: three integers
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        except Exception as e:
            return {"bound": False, "why": "no PCI bus id (%s)" % type(e).__name__}
    try:
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                                    # NVML: 8-digit domain
            bus = bus[4:]
        if not os.path.exists("/sys/bus/pci/devices/%s/numa_node" % bus):
            return {"bound": False, "why": "no /sys/bus/pci/devices/%s/numa_node" % bus}
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"bound": False, "why": "numa_node = -1 (single node)", "pci": bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        local = cpus & allowed
        if not local:
            return {"bound": False, "why": "the process may not run on node %d's CPUs" % node, "pci": bus, "node": node, "allowed_cpus": len(allowed)}
        os.sched_setaffinity(0, local)
        return {"bound": True, "pci": bus, "node": node, "cpus": len(local), "_previous": sorted(allowed)}
    except Exception as e:
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


def config_blobs(cfg, seed, rank=0):
    """(instance blobs, env -> instance map) of one rank's shard."""
    B = cfg["envs"]
    if cfg.get("brandimarte"):
        z = np.load(os.path.join(ROOT, "tests", "golden", "brandimarte_blobs.npz"))   # the reference's data/benchmark/Brandimarte_Data
        blobs = [z["Mk%02d" % k] for k in range(1, 11)]
        return blobs, (np.arange(B) % 10).astype(np.int32)
    n = min(B, cfg.get("distinct", B))
    insts = make_instances(n, seed + 7919 * rank, cfg["machines"], cfg["orders"], cfg["profile"], cfg.get("breakdowns", False))
    return [i.to_blob() for i in insts], (np.arange(B) % n).astype(np.int32)


def make_actions(rng, T, B, variant):
    nt, nm = (6, 5) if variant in ("SO_DFJSP", "SO_FJSSP") else (12, 10)
    a = np.stack([rng.integers(0, nt, (T, B)), rng.integers(0, nm, (T, B))], -1).astype(np.int32)
    r = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
    return a, r


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region: NVML in-process (a query takes microseconds, so a
    25 ms timed region still gets samples), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML indexes the physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                    try:
                        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    flags = [bool(r & getattr(n, k, 0)) for k in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                                                  "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
                    self.samples.append([str(sm), str(self.max_sm)] + ["Active" if f else "Not Active" for f in flags])
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def oracle_mod():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    return oracle_py


def cpu_port_throughput(blobs, env_inst, variant, seconds, T, seed, burnin):
    """The oracle (C port of the reference env) on the host cores, a bounded sample of the same workload: 64
    environment copies per host thread (enough work per thread that starting the threads does not show), a
    burn-in of up to a third of the budget so that the sample sits in the steady mix of episode phases like the
    GPU arm's timed region, then `T`-step rollouts with the same action distribution until the budget."""
    oracle_py = oracle_mod()
    threads = oracle_py.lib().fjsp_oracle_max_threads()
    B = min(len(env_inst), max(threads * 64, 64))
    pick = np.unique(np.linspace(0, len(env_inst) - 1, B).astype(np.int64))
    B = len(pick)
    Tc = max(T, 32, 65536 // B)      # steps per call: the per-call host work (threads, Python) must not show either
    envs = [oracle_py.OracleEnv(blobs[env_inst[i]], variant) for i in pick]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(seed)

    pool = [make_actions(rng, Tc, B, variant) for _ in range(4)]      # generated outside the timed region
    count = [0]

    def one():
        a, r = pool[count[0] % len(pool)]
        count[0] += 1
        oracle_py.batch_rollout(envs, a, r, 1, want_state=True, want_rec=False, threads=threads)
    burned, t0 = 0, time.perf_counter()
    while burned < burnin and time.perf_counter() - t0 < seconds / 3.0:
        one()
        burned += Tc
    steps, t0 = 0, time.perf_counter()
    while steps == 0 or time.perf_counter() - t0 < seconds * 2.0 / 3.0:
        one()
        steps += Tc * B
    dt = time.perf_counter() - t0
    used = min(threads, B)
    return {"value": steps / dt, "unit": UNIT, "cores": used, "kind": "port",
            "sample": f"{B} of the batch's environment copies x {steps // B} steps of the same workload in {dt:.1f}s on {used} host "
                      f"threads after {burned} burn-in steps per copy (oracle/fjsp_oracle.c, the C port pinned bit-exact to the Python reference)"}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU implementation of the path (its C port, oracle/: the Python
    reference needs CPLEX and cannot travel) on ALL host threads, on this config's own workload: the same
    instances, T and action distribution; a step = T env steps of up to 4096 of the config's copies."""
    if rank != 0:
        return
    oracle_py = oracle_mod()
    variant, T, lps = cfg["variant"], cfg["T"], cfg["launches_per_step"]
    blobs, env_inst = config_blobs(cfg, args.seed, 0)
    B = min(cfg["envs"], 4096)
    threads = min(oracle_py.lib().fjsp_oracle_max_threads(), B)     # a thread per environment copy at most
    envs = [oracle_py.OracleEnv(blobs[env_inst[i]], variant) for i in range(B)]
    for e in envs:
        e.reset()
    rng = np.random.default_rng(args.seed)
    Tc = T * lps

    pool = [make_actions(rng, Tc, B, variant) for _ in range(8)]      # generated outside the timed region, like the GPU arm's
    count = [0]

    def one():
        a, r = pool[count[0] % len(pool)]
        count[0] += 1
        oracle_py.batch_rollout(envs, a, r, 1, want_rec=False, threads=threads)
    # untimed burn-in so the sample sits in the same mix of episode phases as the GPU arm's timed region
    burn = min(cfg["burnin"], 1024)
    for _ in range((burn + Tc - 1) // Tc):
        one()
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    value = args.steps * Tc * B / dt
    sample = (f"{B} of the config's {cfg['envs']} environment copies x {args.steps * Tc} steps in {dt:.1f}s on {threads} host "
              f"threads after {burn} burn-in steps per copy (C port of the reference env, oracle/fjsp_oracle.c; the Python "
              f"reference itself needs CPLEX and runs 150-630 steps/s on one core, see BASELINE.md)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config, "envs_per_gpu": cfg["envs"], "envs_sampled": B,
                       "env_steps_per_step": Tc, "launches_per_step": lps, "machines": cfg.get("machines"),
                       "orders": cfg.get("orders"), "variant": variant, "distinct_instances_per_gpu": len(blobs),
                       "parallelism": f"host threads x{threads}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def parity_sample(vec, blobs, env_inst, variant, history, sample, dev, rng):
    """Replays `sample` environments of the timed batch on the CPU oracle from reset() through every
    launch the GPU ran (`history`: per launch the [T, len(sample), 2] actions / draws they were fed), then
    runs ONE more launch on both with the schedule records requested and compares everything."""
    import torch
    oracle_py = oracle_mod()
    envs = [oracle_py.OracleEnv(blobs[env_inst[e]], variant) for e in sample]
    for e in envs:
        e.reset()
    steps = 0
    for a, r in history:
        oracle_py.batch_rollout(envs, a, r, 1, want_state=False, want_rec=False)
        steps += a.shape[0]
    T, B = history[-1][0].shape[0], vec.n_envs
    a, r = make_actions(rng, T, B, variant)
    o = vec.rollout(torch.from_numpy(a).to(dev), torch.from_numpy(r.view(np.int32)).to(dev), reward_policy=1, want_rec=True)
    torch.cuda.synchronize(dev)
    ref = oracle_py.batch_rollout(envs, a[:, sample], r[:, sample], 1)
    st = o["state"][:, sample].cpu().numpy()
    rec_eq = bool(np.array_equal(o["rec"][:, sample].cpu().numpy(), ref["rec"]))
    rew_eq = bool(np.array_equal(o["reward"][:, sample].cpu().numpy(), ref["reward"]))
    done_eq = bool(np.array_equal(o["done"][:, sample].cpu().numpy(), ref["done"]))
    # float features: 1e-5 relative (the north star's bar) with an absolute floor of 1e-9 for features that are
    # (numerically) zero in the reference, e.g. the spread of identical values
    rs = ref["state"]
    diff = np.abs(st - rs)
    big = np.abs(rs) >= 1e-6
    rel = float(np.max(diff[big] / np.abs(rs[big]))) if big.any() else 0.0
    mabs = float(diff.max())
    state_ok = bool(np.all(diff <= 1e-5 * np.abs(rs) + 1e-9))
    w = np.unravel_index(int(np.argmax(diff - 1e-5 * np.abs(rs))), diff.shape)
    info = vec.info()
    oi = [e.info() for e in envs]
    time_eq = bool(np.array_equal(info["step_time"][sample], [x["step_time"] for x in oi]))
    return {"envs": len(sample), "env_steps_replayed_per_env": steps + T, "schedule_records_equal": rec_eq,
            "rewards_equal": rew_eq, "dones_equal": done_eq, "clock_equal": time_eq, "state_max_rel_err": rel,
            "state_max_abs_err": mabs, "state_within_tolerance": state_ok,
            "state_worst_entry": {"t": int(w[0]), "env": int(sample[w[1]]), "feature": int(w[2]), "b200": float(st[w]), "oracle": float(rs[w])},
            "ok": rec_eq and rew_eq and done_eq and time_eq and state_ok,
            "how": "oracle/fjsp_oracle.c replayed from reset() through the burn-in, warm-up and timed launches, "
                   "then one more launch compared output for output (schedule records, rewards, dones, clocks: equal; "
                   "float64 states: |diff| <= 1e-5 |ref| + 1e-9; state_max_rel_err is over |ref| >= 1e-6)"}


def policy_in_loop(blobs, env_inst, variant, local_rank, world, n_steps=200, seed=0):
    """One policy.forward + one step() per environment step (the loop every reference agent runs),
    device-timed: `PolicyRollout` replays policy MLP + Gumbel-max sampling + packing + step kernel as
    one CUDA graph.  Also the agent-side NCCL use: a PPO-shaped gradient all-reduce of the policy."""
    import torch
    import torch.distributed as dist
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    from deep_reinforcement_learning_for_fjsp_b200.agent_loop import PolicyRollout, make_mlp
    from deep_reinforcement_learning_for_fjsp_b200 import sharding
    dev = torch.device("cuda", local_rank)
    vec = FJSPVecEnv(None, env_inst, variant, device=local_rank, blobs=blobs)
    nt, nm = vec.actions_size
    net = make_mlp(vec.state_size, nt * nm, device=dev, seed=seed)       # same weights on every rank
    res = {"envs_per_gpu": vec.n_envs, "policy": f"MLP {vec.state_size}-200-200-200-{nt * nm} (DDQN ActorNet shape), fp32"}
    for graph in (True, False):
        pr = PolicyRollout(vec, net, reward_policy=1, use_graph=graph)
        pr.run(64)                                                       # warm-up / desynchronise a little
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = n_steps if graph else max(20, n_steps // 4)
        e0.record(torch.cuda.current_stream(dev))
        pr.run(n)
        e1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        ms, total = sharding.reduce_timing(e0.elapsed_time(e1), vec.n_envs * n, dev)
        res["cuda_graph" if graph else "eager_launches"] = {"value": total / (ms / 1e3), "unit": UNIT, "steps": n,
                                                            "ms_per_step": ms / n}
        if graph:
            # a PPO-shaped update on the last step's batch: loss -> backward -> gradient all-reduce (NCCL)
            st = pr.state.clone()
            logits = net(st)
            loss = -(torch.log_softmax(logits, -1).max(-1).values * pr.out["reward"][0].float()).mean()
            loss.backward()
            torch.cuda.synchronize(dev)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(torch.cuda.current_stream(dev))
            nel = sharding.allreduce_gradients(net.parameters())
            g1.record(torch.cuda.current_stream(dev))
            torch.cuda.synchronize(dev)
            res["gradient_allreduce"] = {"elements": int(nel), "ms": g0.elapsed_time(g1),
                                         "backend": dist.get_backend() if world > 1 else "single rank (no collective)"}
            net.zero_grad(set_to_none=True)
        res["env_errors"] = int((vec.info()["error"] != 0).sum())
    vec.close()
    return res


def main():
    args, cfg = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    # rank 0 prints ONE JSON line on stdout: everything else a library writes to fd 1 (NCCL's version banner is a
    # plain printf) goes to stderr for the whole run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    from deep_reinforcement_learning_for_fjsp_b200 import _lib, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    host_numa = bind_to_gpu_numa_node(torch, local_rank)
    prev_affinity = host_numa.pop("_previous", None)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    variant, B, T, LPS = cfg["variant"], cfg["envs"], cfg["T"], cfg["launches_per_step"]
    K, W = args.steps, max(args.warmup, 3)
    # every rank plays its own shard: no data-path collective
    blobs, env_inst = config_blobs(cfg, args.seed, rank)
    vec = FJSPVecEnv(None, env_inst, variant, device=local_rank, blobs=blobs)
    q = vec.query()
    vec.reset()
    rng = np.random.default_rng(args.seed + rank)
    sample = np.unique(np.linspace(0, B - 1, min(args.parity_envs, B)).astype(np.int64))
    history = []                                    # what the sample environments were fed, launch by launch
    # inputs for every launch, resident in HBM before the timed region (a pool that the burn-in cycles through)
    NP_ = min(W + K, 8) if B * T > 2_000_000 else W + K
    acts, rnds, acts_h, rnds_h, acts_full, rnds_full = [], [], [], [], [], []
    for _ in range(NP_):
        a, r = make_actions(rng, T, B, variant)
        acts.append(torch.from_numpy(a).to(dev))
        rnds.append(torch.from_numpy(r.view(np.int32)).to(dev))
        acts_h.append(a[:, sample].copy())
        acts_full.append(a)
        rnds_full.append(r)
        rnds_h.append(r[:, sample].copy())
    out = {"state": torch.empty((T, B, vec.state_size), dtype=torch.float32, device=dev),
           "reward": torch.empty((T, B), dtype=torch.float64, device=dev),
           "done": torch.empty((T, B), dtype=torch.int32, device=dev)}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def launch(ia, ir):
        vec.rollout(acts[ia], rnds[ir], reward_policy=1, out=out, state_dtype=torch.float32)
        history.append((acts_h[ia], rnds_h[ir]))

    for i in range((cfg["burnin"] + T - 1) // T):   # untimed burn-in: reach the steady mix of episode phases
        launch(i % NP_, (i * 7 + 3) % NP_)
    for i in range(W * LPS):
        launch(i % NP_, i % NP_)
    barrier()
    info0 = vec.info()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = vec.query()["launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for i in range(K):
        if not args.no_flush:
            flush.fill_(i & 0xff)                   # evict L2 between timed steps (not timed)
        ev[i][0].record(stream)
        for j in range(LPS):
            launch((W + i * LPS + j) % NP_, (W + i * LPS + j) % NP_)
        ev[i][1].record(stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.summary()                      # (stops the sampler: nvidia-smi queries stall the driver's launch path)
    sampler.join(timeout=2.0)
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(per_step_ms))
    launches = vec.query()["launches"] - launches0
    info1 = vec.info()
    errors = int((info1["error"] != 0).sum())
    lp_solves = int((info1["lp_solves"] - info0["lp_solves"]).sum())
    episodes = int((info1["episodes"] - info0["episodes"]).sum())
    # ---- parity of the timed batch itself: a sample of its environments against the CPU oracle
    parity = parity_sample(vec, blobs, env_inst, variant, history, sample, dev, rng) if len(sample) else None
    # ---- e2e: the host-buffer C-ABI calls, pinned host buffers, every copy inside the timed region.  Each
    # call copies that step's actions / draws host-to-device and delivers state, reward and done in page-locked
    # host memory.  The launch time of this workload depends on which environments meet their order arrivals in
    # the launch (1.1-2.3 ms over a few hundred launches: tools/e2e_probe.py), so every e2e mode runs on a TWIN
    # batch -- the same instances, brought through the same burn-in and warm-up launches -- and is fed the very
    # launches of the device-timed region, from host memory: `e2e` and `value` time the same env steps.
    Te = T * LPS
    nburn = (cfg["burnin"] + T - 1) // T
    ha = [torch.from_numpy(a).pin_memory() for a in acts_full]
    hr = [torch.from_numpy(r.view(np.int32)).pin_memory() for r in rnds_full]
    hs = [torch.empty((LPS, T, B, vec.state_size), dtype=torch.float32).pin_memory() for _ in range(2)]
    hrw = [torch.empty((LPS, T, B), dtype=torch.float64).pin_memory() for _ in range(2)]
    hdn = [torch.empty((LPS, T, B), dtype=torch.int32).pin_memory() for _ in range(2)]
    L = vec._L

    def twin():
        tv = FJSPVecEnv(None, env_inst, variant, device=local_rank, blobs=blobs)
        tv.reset()
        for i in range(nburn):
            tv.rollout(acts[i % NP_], rnds[(i * 7 + 3) % NP_], reward_policy=1, out=out, state_dtype=torch.float32)
        torch.cuda.synchronize(dev)
        return tv

    def host_call(tv, n, k, fn, j=0):
        """the launch that used input set n % NP_ as one host-buffer call into slice j of buffer set k"""
        s_ = n % NP_
        _lib.check(fn(tv._h, T, ha[s_].data_ptr(), hr[s_].data_ptr(), 1, 1.0, 1.0, 1.0, 1,
                      None, hs[k][j].data_ptr(), hrw[k][j].data_ptr(), hdn[k][j].data_ptr(), None))

    def e2e_blocking(no_kernel_stores):
        if no_kernel_stores:
            os.environ["FJSP_ZEROCOPY"] = "0"
        tv = twin()
        for n in range(W * LPS):
            host_call(tv, n, 0, L.fjsp_vec_step_host)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            for j in range(LPS):
                host_call(tv, W + i * LPS + j, i % 2, L.fjsp_vec_step_host, j)
        barrier()
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(tv.info()["step_time"], info1["step_time"]))
        tv.close()
        if no_kernel_stores:
            del os.environ["FJSP_ZEROCOPY"]
        return dt, same

    e2e_sync_s, same_sync = e2e_blocking(False)
    # the same call with the outputs copied out in chunks while the kernel runs (the route buffers the device cannot map take)
    e2e_zc_s = None
    if args.e2e_all and LPS == 1 and T >= 16 and "FJSP_ZEROCOPY" not in os.environ:
        e2e_zc_s, _ = e2e_blocking(True)
    same_pipe = None
    if LPS == 1:
        tv = twin()
        for n in range(W):                                           # warm-up (the pipeline's staging buffers are allocated on first use)
            host_call(tv, n, 0, L.fjsp_vec_step_host_begin)
            _lib.check(L.fjsp_vec_step_host_wait(tv._h))
        barrier()
        t0 = time.perf_counter()
        checksum = 0.0
        dbg = []
        host_call(tv, W, 0, L.fjsp_vec_step_host_begin)
        for i in range(1, K):
            ta = time.perf_counter()
            host_call(tv, W + i, i % 2, L.fjsp_vec_step_host_begin)
            tb = time.perf_counter()
            _lib.check(L.fjsp_vec_step_host_wait(tv._h))           # call i - 1 is complete: its outputs are in host memory
            tc = time.perf_counter()
            checksum += float(hrw[(i - 1) % 2][0, -1, 0])            # the device-to-host read of the step's result
            dbg.append((round((tb - ta) * 1e3, 3), round((tc - tb) * 1e3, 3), round((time.perf_counter() - tc) * 1e3, 3)))
        if os.environ.get("FJSP_BENCH_DEBUG"):
            print("pipelined calls (begin, wait, read) ms:", dbg, file=sys.stderr)
        _lib.check(L.fjsp_vec_step_host_wait(tv._h))
        checksum += float(hrw[(K - 1) % 2][0, -1, 0])
        barrier()
        e2e_s = time.perf_counter() - t0
        same_pipe = bool(np.array_equal(tv.info()["step_time"], info1["step_time"]))
        tv.close()
        e2e_api = ("fjsp_vec_step_host_begin / _wait (C ABI, two calls in flight, pinned host buffers: input copy, kernels and "
                   "output copy of consecutive calls overlap; float32 state out)")
    else:
        e2e_s, e2e_api = e2e_sync_s, "fjsp_vec_step_host (C ABI, one blocking call per step(), pinned host buffers)"
    h2d = (ha[0].numel() * 4 + hr[0].numel() * 4) * LPS
    d2h = hs[0].numel() * 4 + hrw[0].numel() * 8 + hdn[0].numel() * 4
    # what the link of THIS box gives a plain device-to-host copy into page-locked memory (the e2e numbers are
    # bound by it: 148 bytes cross the link per env step)
    dsrc = torch.empty(hs[0].numel(), dtype=torch.float32, device=dev)
    lk = []
    for _ in range(4):
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        hs[0].view(-1).copy_(dsrc, non_blocking=True)
        l1.record(stream)
        torch.cuda.synchronize(dev)
        lk.append(dsrc.numel() * 4 / (l0.elapsed_time(l1) / 1e3) / 1e9)
    link_d2h_gbs = max(lk[1:])
    del ha, hr, hs, hrw, hdn, dsrc
    # ---- other rollout lengths on the same batch (not the headline): T = 1 is one reference
    # step() per launch, T = 128 a PPO-style rollout
    sweep = []
    if args.sweep and B * 128 <= 4096 * 128 * 4:
        for T2, n2 in ((1, 64), (128, max(3, K // 4))):
            if T2 == T:
                continue
            a2 = [torch.from_numpy(make_actions(rng, T2, B, variant)[0]).to(dev) for _ in range(4)]
            r2 = [torch.from_numpy(make_actions(rng, T2, B, variant)[1].view(np.int32)).to(dev) for _ in range(4)]
            o2 = {"state": torch.empty((T2, B, vec.state_size), dtype=torch.float32, device=dev),
                  "reward": torch.empty((T2, B), dtype=torch.float64, device=dev),
                  "done": torch.empty((T2, B), dtype=torch.int32, device=dev)}
            for i in range(3):
                vec.rollout(a2[i % 4], r2[(i + 1) % 4], reward_policy=1, out=o2, state_dtype=torch.float32)
            lp_a = int(vec.info()["lp_solves"].sum())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(n2 + 1)]
            torch.cuda.synchronize(dev)
            e0.record(stream)
            evs[0].record(stream)
            for i in range(n2):
                vec.rollout(a2[i % 4], r2[(i + 1) % 4], reward_policy=1, out=o2, state_dtype=torch.float32)
                evs[i + 1].record(stream)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms2 = e0.elapsed_time(e1)
            each = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n2))
            sweep.append({"env_steps_per_launch": T2, "launches": n2, "value": B * T2 * n2 / (ms2 / 1e3), "unit": UNIT,
                          "ms_per_launch": ms2 / n2, "ms_min_median_max": [each[0], each[n2 // 2], each[-1]],
                          "fluid_lp_solves": int(vec.info()["lp_solves"].sum()) - lp_a})
            del a2, r2, o2
    # ---- the same kernels with enough copies to fill the machine (not the headline)
    large = None
    if cfg["large"] and cfg["large"] > B:
        BL = cfg["large"]
        vecL = FJSPVecEnv(None, np.arange(BL) % len(blobs), variant, device=local_rank, blobs=blobs)
        vecL.reset()
        aL = [torch.from_numpy(make_actions(rng, T, BL, variant)[0]).to(dev) for _ in range(2)]
        rL = [torch.from_numpy(make_actions(rng, T, BL, variant)[1].view(np.int32)).to(dev) for _ in range(2)]
        outL = {"state": torch.empty((T, BL, vec.state_size), dtype=torch.float32, device=dev),
                "reward": torch.empty((T, BL), dtype=torch.float64, device=dev),
                "done": torch.empty((T, BL), dtype=torch.int32, device=dev)}
        for i in range(min(cfg["burnin"], 1024) // T + 3):
            vecL.rollout(aL[i % 2], rL[(i + 1) % 2], reward_policy=1, out=outL, state_dtype=torch.float32)
        barrier()
        KL = max(3, K // 4)
        evL = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KL)]
        for i in range(KL):
            evL[i][0].record(stream)
            vecL.rollout(aL[i % 2], rL[(i + 1) % 2], reward_policy=1, out=outL, state_dtype=torch.float32)
            evL[i][1].record(stream)
        torch.cuda.synchronize(dev)
        msL, stepsL = sharding.reduce_timing(sum(a.elapsed_time(b) for a, b in evL), BL * T * KL, dev)
        infL = vecL.info()
        badL = np.nonzero(infL["error"])[0]
        large = {"envs_per_gpu": BL, "value": stepsL / (msL / 1e3), "unit": UNIT, "launches": KL,
                 "ms_per_launch": msL / KL, "env_errors": int(len(badL)),
                 "env_error_flags": [(int(e), int(infL["error"][e]), int(infL["step_count"][e]), int(infL["lp_solves"][e]), int(infL["lp_iterations"][e])) for e in badL[:8]],
                 "note": "instances replicated; env table > L2, no flush needed"}
        vecL.close()
        del vecL, outL, aL, rL
    # ---- the agent loop: one policy forward + one step() per env step, CUDA-graphed (not the headline)
    pol = []
    if args.policy and B >= 64:
        for Bp in sorted({min(B, 4096), cfg["large"] or B, B} - {0}):
            try:
                pol.append(policy_in_loop(blobs, (np.arange(Bp) % len(blobs)).astype(np.int32), variant, local_rank, world,
                                          seed=args.seed))
            except Exception as e:   # the headline must survive a failure of a side line
                pol.append({"envs_per_gpu": Bp, "error": repr(e)[:300]})
    # ---- rollout statistics of every copy, gathered over the ranks (NCCL at N > 1; the only collectives
    # besides the timing reduction: there is none on the data path)
    stats = sharding.gather_episode_stats(info1["completion_time"], info1["delay_time_sum"], info1["energy_consumption"], dev)
    # ---- max over ranks
    dev_ms_max, total_steps = sharding.reduce_timing(dev_ms, B * T * LPS * K, dev)
    e2e_ms_max, _ = sharding.reduce_timing(e2e_s * 1e3, 0, dev)
    e2e_sync_ms_max, _ = sharding.reduce_timing(e2e_sync_s * 1e3, 0, dev)
    value = total_steps / (dev_ms_max / 1e3)
    e2e_value = total_steps / (e2e_ms_max / 1e3)
    e2e_sync_value = total_steps / (e2e_sync_ms_max / 1e3)
    blocking_api = "fjsp_vec_step_host (one blocking call per bench step; the kernel stores its outputs into the page-locked buffers while it runs)"
    e2e_modes = {"pipelined_calls": {"value": e2e_value, "unit": UNIT, "api": e2e_api},
                 "blocking_call": {"value": e2e_sync_value, "unit": UNIT, "api": blocking_api}}
    if e2e_zc_s is not None:
        e2e_zc_ms_max, _ = sharding.reduce_timing(e2e_zc_s * 1e3, 0, dev)
        e2e_modes["blocking_call_chunked_copies"] = {"value": total_steps / (e2e_zc_ms_max / 1e3), "unit": UNIT,
                                                     "api": "fjsp_vec_step_host with FJSP_ZEROCOPY=0 (device staging; chunks of 8 steps leave through the copy "
                                                            "engine while the kernel runs, the host polling the kernel's progress words)"}
    if e2e_sync_value > e2e_value:
        e2e_value, e2e_api, e2e_s = e2e_sync_value, e2e_modes["blocking_call"]["api"], e2e_sync_s
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # algorithmic bytes of one launch (DESIGN.md "roofline"): each env record read and
        # written once, plus per env-step inputs (2 x int32 action + 2 x uint32 draw) and
        # outputs (float32 state, float64 reward, int32 done)
        per_env_step_io = 16 + vec.state_size * 4 + 8 + 4
        algo_bytes = B * (2 * q["env_record_bytes"] + T * per_env_step_io)
        ncu = {}
        npath = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
        if os.path.exists(npath):
            ncu = json.load(open(npath)).get(args.config, {}) if (B, T) == (CONFIGS[args.config]["envs"], CONFIGS[args.config]["T"]) else {}
        launch_s = (dev_ms / (K * LPS)) / 1e3
        achieved = algo_bytes / launch_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "state_out_dtype": "f32 (timed launches and e2e write the observation as float32; clocks and counters are int32/int64, rule keys and features float64)",
                "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config, "envs_per_gpu": B, "env_steps_per_step": T * LPS,
                           "launches_per_step": LPS, "machines": cfg.get("machines"), "orders": cfg.get("orders"), "variant": variant,
                           "distinct_instances_per_gpu": len(blobs),
                           "l2": "NOT flushed (diagnostic run)" if args.no_flush else "flushed between timed steps (256 MiB fill)",
                           "kernels_per_step": "flag + pack kernels (LP-aware env-to-warp map), step kernel (env CTAs of lockstep warps + LP-server CTAs)",
                           "parallelism": f"shard{world}", "env_record_bytes": q["env_record_bytes"], "grid": q["grid"],
                           "block": q["block"], "env_warps": q["env_warps"], "lp_server_ctas": q["lp_server_ctas"],
                           "step_kernel_dynamic_smem": q["step_smem_bytes"]},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "link_gbs": (h2d + d2h) * K / e2e_s / 1e9, "api": e2e_api,
                        "link_d2h_copy_gbs_measured": link_d2h_gbs,
                        "link_bound_value": link_d2h_gbs * 1e9 / (d2h / (B * Te)) * world,
                        "modes": e2e_modes,
                        "same_env_steps_as_value": {"blocking_call": same_sync, "pipelined_calls": same_pipe,
                                                    "how": "each mode runs on a twin batch (same instances, same burn-in and warm-up launches) and "
                                                           "replays the device-timed region's launches from host memory; true = the twin's clocks "
                                                           "after its timed region equal the device-timed batch's"},
                        "note": "value = the faster of the two public host-buffer APIs on this run; both move every input and output over the link inside the timed region (the link of a shared box is the noisy part: see link_gbs)"},
                "host_numa": host_numa,
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu.get("dram_bytes_per_launch"), "peak_source": peak_src,
                             "kernel": "fjsp_step_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                             "launch_ms": launch_s * 1e3,
                             "note": "latency-bound discrete-event kernel, not HBM-bound (see roofline_issue and DESIGN.md section 4)"},
                "roofline_issue": {"bound": "warp-instruction issue", "inst_per_env_step": ncu.get("warp_inst_per_env_step"),
                                   "issue_pct_of_peak": ncu.get("inst_executed_pct_of_peak"),
                                   "source": ncu.get("source", "no ncu capture committed for this config")},
                "parity_sample": parity,
                "wall_s_timed_region": wall, "env_errors": errors,
                "timed_region_events": {"fluid_lp_solves": lp_solves, "episodes_finished": episodes,
                                        "burnin_env_steps_per_copy": cfg["burnin"]},
                "episode_stats_all_ranks": {"copies": int(stats.shape[0]), "mean_completion_time": float(stats[:, 0].mean()),
                                            "mean_delay_time_sum": float(stats[:, 1].mean()),
                                            "gathered_with": "sharding.gather_episode_stats (" + (dist.get_backend() if world > 1 else "single rank") + ")"},
                "step_ms_min_max": [min(per_step_ms), max(per_step_ms)], "step_ms_all": [round(x, 4) for x in per_step_ms],
                "large_batch": large,
                "rollout_sweep": sweep, "policy_in_loop": pol}
        if not args.no_cpu_baseline:
            if host_numa.get("bound"):
                os.sched_setaffinity(0, prev_affinity)             # the CPU baseline gets every core the process had
            line["cpu_baseline"] = cpu_port_throughput(blobs, env_inst, variant, args.cpu_seconds, T, args.seed, cfg["burnin"])
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
