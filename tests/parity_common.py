"""Shared checks: a vector-environment backend (the CUDA library through its C ABI, or the
one-lane host build of the same kernel source) against the golden reference trajectories
and against the CPU oracle."""
import os

import numpy as np

import oracle_py
from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance

# state features: the task's bar is 1e-5 relative to the reference's float64; the kernel's
# observation-only sums use a warp tree instead of CPython's compensated sum, which moves
# the last bits only, so the tests hold it to a far tighter bound.
STATE_RTOL, STATE_ATOL = 1e-9, 1e-12
NRULES = {"SO_DFJSP": (6, 5), "MO_DFJSP": (12, 10), "MO_DFJSP_breakdown": (12, 10), "SO_FJSSP": (6, 5)}


def assert_states_close(got, want, what):
    ok = np.isclose(got, want, rtol=STATE_RTOL, atol=STATE_ATOL)
    assert ok.all(), f"{what}: state feature mismatch at {np.argwhere(~ok)[:5].tolist()} " \
                     f"got {got[~ok][:5]} want {want[~ok][:5]}"


def replay_golden(make_vec, golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"), allow_pickle=False)
    variant = str(g["variant"])
    vec = make_vec([g["blob"]], [0], variant)
    actions, rnd, rp = g["actions"], g["rnd"], int(g["reward_policy"])
    T = len(actions)
    s0 = vec.reset_host()
    assert_states_close(s0[0], g["resets"][0], "reset")
    # the whole trajectory, auto-reset between the recorded episodes, in launches of up to 64 steps
    t = 0
    while t < T:
        n = min(64, T - t)
        st, rw, dn, rec = vec.step_host(actions[t:t + n, None, :], rnd[t:t + n, None, :], rp, 1.0, 1.0, 1.0, True)
        assert np.array_equal(rec[:, 0], g["recs"][t:t + n]), f"{case}: schedule records differ in steps {t}..{t + n}"
        assert np.array_equal(rw[:, 0], g["rewards"][t:t + n]), f"{case}: rewards differ in steps {t}..{t + n}"
        assert np.array_equal(dn[:, 0], g["dones"][t:t + n]), f"{case}: done flags differ in steps {t}..{t + n}"
        assert_states_close(st[:, 0], g["states"][t:t + n], f"{case} steps {t}..{t + n}")
        t += n
    info = vec.info()
    assert int(info["error"][0]) == 0
    return vec


def random_batch(seed, variant, n_inst, copies, profile_mix=True, breakdowns=False, M=None):
    rng = np.random.default_rng(seed)
    insts = []
    for i in range(n_inst):
        prof = ("DA3C", "HMPSAC")[i % 2] if profile_mix else "DA3C"
        m = M or int(rng.integers(4, 13))
        inst = FJSPInstance.generate(seed * 1000 + i, float(rng.choice([0.5, 1.0, 1.5])), m, int(rng.integers(1, 4)),
                                     prof, breakdowns=breakdowns, scale=0.12 if prof == "DA3C" else 0.4)
        inst.ddt = float(int(inst.ddt))
        insts.append(inst)
    env_instance = np.repeat(np.arange(n_inst), copies)
    return insts, env_instance


def compare_with_oracle(make_vec, variant, seed, n_inst=6, copies=3, T=48, launches=4, reward_policy=1,
                        breakdowns=False, sum_mode=1):
    """Random rule actions on a mixed batch with auto-reset; every output of every step must
    agree with the oracle stepping the same instances one by one."""
    insts, env_instance = random_batch(seed, variant, n_inst, copies, breakdowns=breakdowns)
    blobs = [i.to_blob() for i in insts]
    B = len(env_instance)
    vec = make_vec(blobs, env_instance, variant) if sum_mode == 1 else make_vec(blobs, env_instance, variant, sum_mode)
    envs = [oracle_py.OracleEnv(blobs[k], variant, sum_mode) for k in env_instance]
    s0 = vec.reset_host()
    o0 = np.stack([e.reset() for e in envs])
    assert_states_close(s0, o0, "reset")
    rng = np.random.default_rng(seed + 77)
    nt, nm = NRULES[variant]
    ok = np.ones(B, bool)       # environments the reference has not raised on (it raises where a rule finds nothing
                                # to dispatch -- e.g. right after the reset() of a used object whose machines are still
                                # flagged busy; both sides flag the error, outputs are undefined from there on)
    for L in range(launches):
        actions = np.stack([rng.integers(0, nt, (T, B)), rng.integers(0, nm, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, reward_policy, 1.0, 1.0, 1.0, True)
        ref = oracle_py.batch_rollout(envs, actions, rnd, reward_policy, allow_errors=True)
        if ref["error_flags"]:
            ok &= np.array([e.info()["error"] == 0 for e in envs])
        assert np.array_equal(rec[:, ok], ref["rec"][:, ok]), f"launch {L}: schedule records differ"
        assert np.array_equal(dn[:, ok], ref["done"][:, ok]), f"launch {L}: done flags differ"
        assert np.array_equal(rw[:, ok], ref["reward"][:, ok]), f"launch {L}: rewards differ"
        assert_states_close(st[:, ok], ref["state"][:, ok], f"launch {L}")
    info = vec.info()
    oi = [e.info() for e in envs]
    assert np.array_equal(info["error"], [x["error"] for x in oi]), "error flags differ from the reference's"
    assert ok.sum() >= (B + 1) // 2, "the reference raised on most of the batch: not a useful case"
    assert np.array_equal(info["step_time"][ok], np.array([x["step_time"] for x in oi])[ok])
    assert np.array_equal(info["lp_solves"] >= 1, np.ones(B, bool))
    return vec, envs


def schedule_is_feasible(rec, dn, ntask_of_env):
    """Size-independent property: inside an episode no machine runs two operations at once
    and a job's stages run in order, each starting after the previous stage ended."""
    T, B = rec.shape[:2]
    for b in range(B):
        mach_free, job_end, job_stage = {}, {}, {}
        for t in range(T):
            q, r, j, n, m, tb, te, me = rec[t, b]
            assert tb >= mach_free.get(m, 0), (b, t, "machine overlap")
            assert j == job_stage.get((r, n), 0), (b, t, "stage order")
            assert tb >= job_end.get((r, n), 0), (b, t, "job precedence")
            assert te > tb and me >= te
            mach_free[m] = me
            job_end[(r, n)] = te
            job_stage[(r, n)] = j + 1
            if dn[t, b]:
                mach_free, job_end, job_stage = {}, {}, {}
    return True


def edge_case_instances():
    """Smallest and widest shapes the blob format allows in practice."""
    one = FJSPInstance(machine_count=1, ntask=[3], machine_rj={(0, 0): (0,), (0, 1): (0,), (0, 2): (0,)},
                       time_rjm={(0, 0): {0: 5}, (0, 1): {0: 7}, (0, 2): {0: 3}}, arrive=[0], due=[9], count=[(1,)],
                       ddt=1.0, power_rjm={(0, 0): {0: 10}, (0, 1): {0: 20}, (0, 2): {0: 30}}, idle_power=[2],
                       breakdowns={0: [(6, 8)]}, name="one_machine_one_job")
    wide = FJSPInstance.generate(77, 1.0, 32, 2, "DA3C", scale=0.08)       # 32 machines: every mask bit used
    wide.ddt = 1.0
    late = FJSPInstance.generate(78, 0.5, 3, 4, "HMPSAC", scale=0.3)       # few machines, four orders, tight due dates
    late.ddt = 0.0
    return [one, wide, late]


def check_edge_cases(make_vec, variant):
    """Tiny / wide / overloaded instances in ONE batch, run past the end of their episodes twice:
    first with auto-reset (against the oracle), then without (a finished copy must repeat
    done=1, reward 0, an empty record and its terminal observation with zero differences)."""
    insts = edge_case_instances()
    blobs = [i.to_blob() for i in insts]
    env_instance = np.array([0, 1, 2, 0, 2], np.int32)       # 5 copies: not a multiple of anything
    B = len(env_instance)
    nt, nm = NRULES[variant]
    vec = make_vec(blobs, env_instance, variant)
    envs = [oracle_py.OracleEnv(blobs[k], variant) for k in env_instance]
    assert_states_close(vec.reset_host(), np.stack([e.reset() for e in envs]), "reset")
    rng = np.random.default_rng(5)
    T = 40
    for L in range(3):
        actions = np.stack([rng.integers(0, nt, (T, B)), rng.integers(0, nm, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, 1, 1.0, 1.0, 1.0, True)
        ref = oracle_py.batch_rollout(envs, actions, rnd, 1)
        assert np.array_equal(rec, ref["rec"]) and np.array_equal(rw, ref["reward"]) and np.array_equal(dn, ref["done"])
        assert_states_close(st, ref["state"], f"edge launch {L}")
    assert dn.sum() + 1 > 0 and (vec.info()["episodes"][[0, 3]] >= 1).all()     # the 3-operation job finished many times
    # without auto-reset
    vec2 = make_vec(blobs, env_instance, variant)
    vec2.reset_host()
    T2 = 12
    actions = np.zeros((T2, B, 2), np.int32)
    st, rw, dn, rec = vec2.step_host(actions, None, 1, 1.0, 1.0, 1.0, False)
    for b in (0, 3):                       # the one-job instance is done after its 3 operations
        assert list(dn[:, b]) == [0, 0, 1] + [1] * (T2 - 3)
        assert (rw[3:, b] == 0).all() and (rec[3:, b] == -1).all()
        n = st.shape[-1] // 2
        assert np.array_equal(st[3:, b, :n], np.broadcast_to(st[2, b, :n], (T2 - 3, n)))
        assert (st[3:, b, n:] == 0).all()
    assert (vec2.info()["error"] == 0).all()


def check_reset_of_used_env(make_vec, variant, seed=91):
    """reset() on a USED environment object -- after a finished episode, and in the middle of one (agents only do
    the former; the latter must still do what the reference's reset() does to a half-played object: busy flags,
    order_arrive_time and the `done` seen by the first observation survive) -- against the oracle doing the same,
    error flags included."""
    insts, env_instance = random_batch(seed, variant, 3, 2)
    blobs = [i.to_blob() for i in insts]
    B = len(env_instance)
    vec = make_vec(blobs, env_instance, variant)
    envs = [oracle_py.OracleEnv(blobs[k], variant) for k in env_instance]
    rng = np.random.default_rng(seed)
    nt, nm = NRULES[variant]
    for phase, T in enumerate((25, 40, 30)):                   # reset, 25 steps, reset mid-episode, 40 steps, reset, 30 steps
        s0 = vec.reset_host()
        o0 = np.stack([e.reset() for e in envs])
        assert_states_close(s0, o0, f"reset {phase}")
        actions = np.stack([rng.integers(0, nt, (T, B)), rng.integers(0, nm, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, 1, 1.0, 1.0, 1.0, False)
        # the oracle steps one by one so that an error flag (a rule that finds nothing to dispatch) does not abort the run
        for t in range(T):
            for b, e in enumerate(envs):
                if e.info()["done"]:
                    continue
                try:
                    so, ro, do, reco = e.step(actions[t, b], rnd[t, b], 1)
                except AssertionError:
                    continue
                assert np.array_equal(rec[t, b], reco), (phase, t, b)
                assert rw[t, b] == ro and dn[t, b] == int(do), (phase, t, b)
                assert_states_close(st[t, b], so, f"phase {phase} step {t} env {b}")
    info = vec.info()
    assert np.array_equal(info["error"], [e.info()["error"] for e in envs])
    assert np.array_equal(info["step_time"], [e.info()["step_time"] for e in envs])
