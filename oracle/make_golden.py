"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by stepping the UNMODIFIED
reference environments (imported from /root/reference through oracle/refshim).

Run in the build container (the reference does not exist on the GPU box):
    python oracle/make_golden.py [case ...]      (no names: every case)
Each fixture holds the instance blob, the action / random-word sequences and what the
reference returned: states, rewards, dones, and per step the dispatched operation
(kind, stage, job number), its machine and its begin / end times.  The reference's
`random.choice` is replaced by  seq[word % len(seq)]  with the recorded words (see
oracle/fjsp_oracle.c header); its CPLEX solve is served by oracle/refshim/docplex with
the deterministic simplex of oracle/fjsp_lp.c ("fluid LP specification").
"""
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))

import ref_loader  # noqa: E402
from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance  # noqa: E402
import oracle_py  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FMT = {"SO_DFJSP": "SO", "SO_FJSSP": "SO", "MO_DFJSP": "MO", "MO_DFJSP_breakdown": "MO"}
NRULES = {"SO_DFJSP": (6, 5), "SO_FJSSP": (6, 5), "MO_DFJSP": (12, 10), "MO_DFJSP_breakdown": (12, 10)}


class Chooser:
    def __init__(self):
        self.words = []

    def __call__(self, seq):
        return seq[self.words.pop(0) % len(seq)]


def run_reference(variant, inst, actions, rnd, reward_policy, episodes=1, max_steps=None, fmt_dir=None):
    """Step the reference; returns dict of arrays.  actions/rnd: callables (t)->pair or arrays."""
    Env = ref_loader.load(variant)
    tmp = fmt_dir or tempfile.mkdtemp(prefix="fjsp_golden_")
    inst.write_csv(tmp, "case", FMT[variant])
    env = Env(use_instance=False, path=tmp, file_name="case")
    chooser = Chooser()
    saved = random.choice
    random.choice = chooser
    task_random = (6,) if variant.startswith("SO") else (11, 12)
    mach_random = (5,) if variant.startswith("SO") else (9, 10)
    out = dict(states=[], rewards=[], dones=[], recs=[], resets=[], infos=[])
    try:
        t = 0
        for ep in range(episodes):
            out["resets"].append(np.array(env.reset(), dtype=np.float64))
            while not env.done:
                if max_steps is not None and t >= max_steps:
                    break
                a, w = actions[t], rnd[t]
                chooser.words = []
                if a[0] + 1 in task_random:
                    chooser.words.append(int(w[0]))
                if a[1] + 1 in mach_random:
                    chooser.words.append(int(w[1]))
                before = {m: len(mo.task_list) for m, mo in env.machine_dict.items()}
                if variant.startswith("MO"):
                    st, rw, dn = env.step(tuple(int(x) for x in a), reward_policy=reward_policy,
                                          completion=1.0, tardiness=1.0, energy_consumption=1.0)
                else:
                    st, rw, dn = env.step(tuple(int(x) for x in a))
                assert not chooser.words
                grown = [m for m, mo in env.machine_dict.items() if len(mo.task_list) != before[m]]
                assert len(grown) == 1
                m = grown[0]
                task = env.machine_dict[m].task_list[-1]
                rj = env.kind_task_tuple.index((task.kind, task.task))
                out["recs"].append([rj, task.kind, task.task, task.number, m, task.time_begin, task.time_end,
                                    env.machine_dict[m].time_end])
                out["states"].append(np.array(st, dtype=np.float64))
                out["rewards"].append(float(rw))
                out["dones"].append(int(dn))
                t += 1
            out["infos"].append([env.step_time, env.step_count, getattr(env, "completion_time", 0),
                                 env.delay_time_sum, getattr(env, "energy_consumption", 0),
                                 max(mo.time_end for mo in env.machine_dict.values())])
            if max_steps is not None and t >= max_steps:
                break
    finally:
        random.choice = saved
    return {k: np.array(v) for k, v in out.items()}, t


ONLY = set(sys.argv[1:])      # python oracle/make_golden.py [case ...]: only these cases (default: all)


def make_case(name, variant, inst, seed, episodes=1, max_steps=None, reward_policy=1, fixed_action=None):
    if ONLY and name not in ONLY:
        return None
    # the reference reads DDT back through its integer regex (0.5 -> 0, 1.5 -> 1); keep
    # the blob consistent with what the reference saw
    inst.ddt = float(int(inst.ddt))
    rng = np.random.default_rng(seed)
    cap = (max_steps or inst.total_operations) * episodes + 8
    nt, nm = NRULES[variant]
    actions = np.stack([rng.integers(0, nt, cap), rng.integers(0, nm, cap)], 1).astype(np.int32)
    if fixed_action is not None:
        actions[:] = fixed_action
    rnd = rng.integers(0, 2**32, (cap, 2), dtype=np.uint64).astype(np.uint32)
    ref, T = run_reference(variant, inst, actions, rnd, reward_policy, episodes, max_steps)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, blob=inst.to_blob(), variant=variant, reward_policy=reward_policy,
                        actions=actions[:T], rnd=rnd[:T], episodes=episodes, **ref)
    print(f"{name}: variant={variant} steps={T} episodes={len(ref['infos'])} "
          f"KT={len(inst.kind_task_tuple)} M={inst.machine_count} S={inst.order_count} "
          f"size={os.path.getsize(path) / 1024:.0f} KiB")
    return path


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    oracle_py.build()
    G = FJSPInstance.generate
    ref_data = os.path.join(ref_loader.REFERENCE_ROOT, "data")
    # small generated instances: whole episodes, two episodes on one env object (re-reset quirks)
    make_case("so_dfjsp_small_a", "SO_DFJSP", G(1, 1.0, 5, 2, "DA3C", scale=0.12), 11, episodes=2)
    make_case("so_dfjsp_small_b", "SO_DFJSP", G(2, 0.5, 10, 3, "DA3C", scale=0.1), 12, episodes=2)
    make_case("so_dfjsp_rule30", "SO_DFJSP", G(3, 1.5, 8, 3, "DA3C", scale=0.1), 13, fixed_action=(2, 0))
    make_case("so_fjssp_small", "SO_FJSSP", G(4, 1.0, 6, 2, "DA3C", scale=0.15), 14, episodes=2)
    make_case("mo_dfjsp_small_a", "MO_DFJSP", G(5, 1.0, 6, 3, "HMPSAC", scale=0.5), 15, episodes=2, reward_policy=1)
    make_case("mo_dfjsp_small_b", "MO_DFJSP", G(6, 0.5, 10, 2, "DA3C", scale=0.1), 16, reward_policy=0)
    make_case("mo_dfjsp_small_c", "MO_DFJSP", G(7, 1.5, 12, 4, "HMPSAC", scale=0.4), 17, reward_policy=2)
    make_case("mo_dfjsp_reward3", "MO_DFJSP", G(8, 1.0, 5, 2, "DA3C", scale=0.1), 18, reward_policy=3)
    make_case("mo_breakdown_small_a", "MO_DFJSP_breakdown", G(9, 1.0, 6, 3, "HMPSAC", breakdowns=True, scale=0.5), 19,
              episodes=2)
    make_case("mo_breakdown_small_b", "MO_DFJSP_breakdown", G(10, 0.5, 10, 2, "DA3C", breakdowns=True, scale=0.1), 20)
    # the reference's own data files (first steps of an episode)
    make_case("so_dfjsp_DA3C_DDT0.5_M10_S3", "SO_DFJSP",
              FJSPInstance.from_csv(os.path.join(ref_data, "DA3C"), "DDT0.5_M10_S3", "SO"), 21, max_steps=400)
    make_case("so_dfjsp_Mk01", "SO_DFJSP",
              FJSPInstance.from_csv(os.path.join(ref_data, "benchmark", "Brandimarte_Data"), "Mk01", "SO"), 22)
    make_case("mo_dfjsp_HMPSAC_DDT1.0_M10_S3", "MO_DFJSP",
              FJSPInstance.from_csv(os.path.join(ref_data, "HMPSAC"), "DDT1.0_M10_S3", "MO"), 23, max_steps=250)
    # full-size instances of the bench workloads (Instance_generate.py distributions, scale = 1): the first steps of an
    # episode, which hold its order arrivals (fluid LPs of 50-110 rows; ~300 rows for 20 machines / 5 orders)
    make_case("mo_dfjsp_bench_scale", "MO_DFJSP", G(7000, 0.5, 10, 3, "DA3C"), 24, max_steps=500, reward_policy=1)
    make_case("mo_breakdown_bench_scale", "MO_DFJSP_breakdown", G(7001, 1.0, 10, 3, "DA3C", breakdowns=True), 25,
              max_steps=400, reward_policy=2)
    make_case("mo_dfjsp_bench_scale_two_episodes", "MO_DFJSP", G(7002, 1.5, 10, 3, "DA3C"), 29, episodes=2, reward_policy=1)
    make_case("mo_dfjsp_m20_s5", "MO_DFJSP", G(8000, 1.0, 20, 5, "HMPSAC"), 26, max_steps=300, reward_policy=1)
    make_case("so_fjssp_small_b", "SO_FJSSP", G(11, 0.5, 8, 3, "DA3C", scale=0.12), 27, episodes=2)
    make_case("so_dfjsp_Mk03", "SO_DFJSP",
              FJSPInstance.from_csv(os.path.join(ref_data, "benchmark", "Brandimarte_Data"), "Mk03", "SO"), 28, max_steps=200)


def brandimarte_blobs():
    """tests/golden/brandimarte_blobs.npz: instance blobs of the reference's Mk01-Mk10 csv
    directories (data/benchmark/Brandimarte_Data), for the replicated-copies tests."""
    base = os.path.join(ref_loader.REFERENCE_ROOT, "data", "benchmark", "Brandimarte_Data")
    out = {"Mk%02d" % k: FJSPInstance.from_csv(base, "Mk%02d" % k, "SO").to_blob() for k in range(1, 11)}
    np.savez_compressed(os.path.join(GOLDEN, "brandimarte_blobs.npz"), **out)


if __name__ == "__main__":
    main()
    if not ONLY:
        brandimarte_blobs()
