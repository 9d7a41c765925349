"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
of libfjsp_b200.so (ctypes, host buffers or torch device pointers); the oracle and the
committed golden trajectories are the checkers."""
import numpy as np
import pytest

import parity_common as pc
from conftest import golden_cases

pytestmark = pytest.mark.gpu
DEVICE_CASES = golden_cases()


def make_vec(blobs, env_instance, variant, sum_mode=1):
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    return FJSPVecEnv(None, env_instance, variant, blobs=blobs, sum_mode=sum_mode)


@pytest.mark.parametrize("case", DEVICE_CASES)
def test_golden_reference_trajectories(case, golden_dir):
    pc.replay_golden(make_vec, golden_dir, case)


@pytest.mark.parametrize("variant,seed,rp,bd", [("SO_DFJSP", 11, 1, False), ("MO_DFJSP", 12, 0, False),
                                                 ("MO_DFJSP", 13, 1, False), ("MO_DFJSP", 14, 3, False),
                                                 ("MO_DFJSP_breakdown", 15, 2, True), ("SO_FJSSP", 16, 1, False)])
def test_random_batch_vs_oracle(variant, seed, rp, bd):
    pc.compare_with_oracle(make_vec, variant, seed, n_inst=8, copies=8, T=64, launches=4, reward_policy=rp,
                           breakdowns=bd)


def test_single_steps_equal_fused_rollout():
    """T steps in one launch == T launches of one step (state lives in HBM between launches)."""
    insts, env_instance = pc.random_batch(21, "MO_DFJSP", 4, 16)
    blobs = [i.to_blob() for i in insts]
    a, b = make_vec(blobs, env_instance, "MO_DFJSP"), make_vec(blobs, env_instance, "MO_DFJSP")
    B, T = len(env_instance), 96
    rng = np.random.default_rng(3)
    actions = np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32)
    rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(a.reset_host(), b.reset_host())
    fused = a.step_host(actions, rnd, 1)
    for t in range(T):
        one = b.step_host(actions[t:t + 1], rnd[t:t + 1], 1)
        for x, y in zip(fused, one):
            assert np.array_equal(x[t], y[0])


def test_full_size_properties():
    """BASELINE configs[1] size (4096 copies): replicas of one instance fed the same actions
    stay bit-identical, reruns are deterministic, schedules are feasible, and a sample of
    environments agrees with the oracle."""
    import torch
    from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance
    n_inst, copies, T = 64, 64, 128
    insts = []
    for i in range(n_inst):
        inst = FJSPInstance.generate(500 + i, [0.5, 1.0, 1.5][i % 3], 10, 3, "DA3C", scale=0.3)
        inst.ddt = float(int(inst.ddt))
        insts.append(inst)
    blobs = [i.to_blob() for i in insts]
    env_instance = np.repeat(np.arange(n_inst), copies)
    B = len(env_instance)
    rng = np.random.default_rng(9)
    act1 = np.stack([rng.integers(0, 12, (T, n_inst)), rng.integers(0, 10, (T, n_inst))], -1).astype(np.int32)
    rnd1 = rng.integers(0, 2**32, (T, n_inst, 2), dtype=np.uint64).astype(np.uint32)
    actions = np.repeat(act1, copies, axis=1)
    rnd = np.repeat(rnd1, copies, axis=1)
    outs = []
    for rep in range(2):
        vec = make_vec(blobs, env_instance, "MO_DFJSP")
        vec.reset()
        o = vec.rollout(torch.from_numpy(actions).cuda(), torch.from_numpy(rnd.view(np.int32)).cuda(),
                        reward_policy=1, want_rec=True)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items()})
        assert (vec.info()["error"] == 0).all()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), f"{k}: rerun differs"
    rec = outs[0]["rec"].reshape(T, n_inst, copies, 8)
    assert (rec == rec[:, :, :1]).all(), "replicas of one instance diverged"
    st = outs[0]["state"].reshape(T, n_inst, copies, -1)
    assert np.array_equal(st, np.broadcast_to(st[:, :, :1], st.shape))
    assert pc.schedule_is_feasible(outs[0]["rec"][:, ::copies], outs[0]["done"][:, ::copies], None)
    # oracle on a sample of instances
    import oracle_py
    sample = [0, 17, 63]
    envs = [oracle_py.OracleEnv(blobs[k], "MO_DFJSP") for k in sample]
    for e in envs:
        e.reset()
    ref = oracle_py.batch_rollout(envs, act1[:, sample], rnd1[:, sample], 1)
    assert np.array_equal(ref["rec"], rec[:, sample, 0])
    assert np.array_equal(ref["reward"], outs[0]["reward"].reshape(T, n_inst, copies)[:, sample, 0])
    pc.assert_states_close(st[:, sample, 0], ref["state"], "full size sample")


def test_error_behaviour():
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv, FJSPEnv, MyError
    with pytest.raises(RuntimeError):
        FJSPVecEnv(None, [0], "MO_DFJSP", blobs=[np.zeros(64, np.int32)])
    with pytest.raises(MyError):
        FJSPVecEnv(None, [0], "NOT_AN_ENV", blobs=[np.zeros(64, np.int32)])
    insts, _ = pc.random_batch(31, "MO_DFJSP", 1, 1)
    env = FJSPEnv(variant="MO_DFJSP", instance=insts[0])
    env.reset()
    with pytest.raises(MyError):
        env.step((12, 0), reward_policy=1)      # the reference raises MyError for an undefined rule
    with pytest.raises(MyError):
        env.step((0, 0), reward_policy=7)       # ... and for an undefined reward function


def test_single_env_drop_in_episode():
    """FJSPEnv: the reference's reset()/step(action) loop, one DDQN-style rule episode."""
    import oracle_py
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPEnv
    insts, _ = pc.random_batch(41, "SO_DFJSP", 1, 1)
    env = FJSPEnv(variant="SO_DFJSP", instance=insts[0])
    ora = oracle_py.OracleEnv(insts[0].to_blob(), "SO_DFJSP")
    s, so = env.reset(), ora.reset()
    pc.assert_states_close(s, so, "reset")
    n = 0
    while not env.done:
        a = (n % 5, (n // 5) % 4)
        s, r, d = env.step(a)
        so, ro, do, _ = ora.step(a)
        assert r == ro and d == do
        pc.assert_states_close(s, so, f"step {n}")
        n += 1
    assert env.delay_time_sum == ora.info()["delay_sum"]


def _brandimarte(golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, "brandimarte_blobs.npz"))
    return [z["Mk%02d" % k] for k in range(1, 11)]


def test_brandimarte_replicated_copies(golden_dir):
    """BASELINE configs[3] shape: Brandimarte mk01-mk10 replicated (here 10 x 2048 copies, the
    DA3C environment class SO_DFJSP): replicas fed the same rules stay bit-identical through
    several episodes with auto-reset, and every instance agrees with the oracle."""
    import torch
    import oracle_py
    blobs = _brandimarte(golden_dir)
    copies, T, launches = 2048, 64, 3
    env_instance = np.repeat(np.arange(10), copies)
    B = len(env_instance)
    vec = make_vec(blobs, env_instance, "SO_DFJSP")
    envs = [oracle_py.OracleEnv(b, "SO_DFJSP") for b in blobs]
    s0 = vec.reset().cpu().numpy().reshape(10, copies, -1)
    o0 = np.stack([e.reset() for e in envs])
    pc.assert_states_close(s0[:, 0], o0, "reset")
    assert np.array_equal(s0, np.broadcast_to(s0[:, :1], s0.shape))
    rng = np.random.default_rng(4)
    for L in range(launches):
        act1 = np.stack([rng.integers(0, 6, (T, 10)), rng.integers(0, 5, (T, 10))], -1).astype(np.int32)
        rnd1 = rng.integers(0, 2**32, (T, 10, 2), dtype=np.uint64).astype(np.uint32)
        o = vec.rollout(torch.from_numpy(np.repeat(act1, copies, axis=1)).cuda(),
                        torch.from_numpy(np.repeat(rnd1, copies, axis=1).view(np.int32)).cuda(), want_rec=True)
        torch.cuda.synchronize()
        rec = o["rec"].cpu().numpy().reshape(T, 10, copies, 8)
        st = o["state"].cpu().numpy().reshape(T, 10, copies, -1)
        assert (rec == rec[:, :, :1]).all(), "replicas diverged"
        ref = oracle_py.batch_rollout(envs, act1, rnd1, 1)
        assert np.array_equal(ref["rec"], rec[:, :, 0]), f"launch {L}"
        assert np.array_equal(ref["done"], o["done"].cpu().numpy().reshape(T, 10, copies)[:, :, 0])
        assert np.array_equal(ref["reward"], o["reward"].cpu().numpy().reshape(T, 10, copies)[:, :, 0])
        pc.assert_states_close(st[:, :, 0], ref["state"], f"launch {L}")
    info = vec.info()
    assert (info["error"] == 0).all() and (info["episodes"] >= 1).any()


def test_large_dynamic_instances():
    """BASELINE configs[4] shape: 20 machines, five orders (about 200 inserted jobs), MO_DFJSP:
    a batch of copies against the oracle, LP rows in the hundreds."""
    from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance
    import oracle_py
    insts = []
    for i in range(3):
        inst = FJSPInstance.generate(900 + i, 1.0, 20, 5, "HMPSAC", scale=0.6)
        inst.ddt = float(int(inst.ddt))
        insts.append(inst)
    blobs = [i.to_blob() for i in insts]
    env_instance = np.repeat(np.arange(3), 16)
    B, T = len(env_instance), 64
    vec = make_vec(blobs, env_instance, "MO_DFJSP")
    envs = [oracle_py.OracleEnv(blobs[k], "MO_DFJSP") for k in env_instance]
    pc.assert_states_close(vec.reset_host(), np.stack([e.reset() for e in envs]), "reset")
    rng = np.random.default_rng(8)
    for L in range(3):
        actions = np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, 2)
        ref = oracle_py.batch_rollout(envs, actions, rnd, 2)
        assert np.array_equal(rec, ref["rec"]) and np.array_equal(rw, ref["reward"]) and np.array_equal(dn, ref["done"])
        pc.assert_states_close(st, ref["state"], f"launch {L}")
    assert (vec.info()["error"] == 0).all()


@pytest.mark.parametrize("variant", ["SO_DFJSP", "MO_DFJSP", "MO_DFJSP_breakdown", "SO_FJSSP"])
def test_edge_cases(variant):
    """one machine / one job, 32 machines, overloaded shop; 5 copies; with and without auto-reset"""
    pc.check_edge_cases(make_vec, variant)


def test_left_to_right_sum_mode_vs_oracle():
    pc.compare_with_oracle(make_vec, "SO_DFJSP", 19, n_inst=4, copies=4, T=48, launches=2, sum_mode=0)


@pytest.mark.parametrize("n_inst,copies", [(3, 1), (5, 7), (16, 40), (8, 600)])
def test_lp_aware_packing_places_every_env_once(n_inst, copies):
    """The warp-slot map the packing kernel writes before each launch holds every environment
    exactly once, whatever the batch size (one partial CTA, several CTAs, more than one round of
    CTAs), and the launch that used it agrees with the oracle on a sample of environments."""
    import oracle_py
    insts, env_instance = pc.random_batch(77 + n_inst, "MO_DFJSP", n_inst, copies)
    blobs = [i.to_blob() for i in insts]
    vec = make_vec(blobs, env_instance, "MO_DFJSP")
    B, T = len(env_instance), 48
    rng = np.random.default_rng(5)
    vec.reset_host()
    sample = list(range(0, B, max(1, B // 12)))
    envs = [oracle_py.OracleEnv(blobs[env_instance[e]], "MO_DFJSP") for e in sample]
    for e in envs:
        e.reset()
    for launch in range(6):
        actions = np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, 1)
        sl = vec.slots()
        placed = sl[sl >= 0]
        assert placed.size == B and np.array_equal(np.sort(placed), np.arange(B))
        ref = oracle_py.batch_rollout(envs, actions[:, sample], rnd[:, sample], 1)
        assert np.array_equal(rec[:, sample], ref["rec"]) and np.array_equal(dn[:, sample], ref["done"])
        assert np.array_equal(rw[:, sample], ref["reward"])
        assert np.allclose(st[:, sample], ref["state"], rtol=1e-9, atol=1e-12)
    assert (vec.info()["error"] == 0).all()


def test_bench_scale_instances_three_rounds_whole_episodes():
    """The bench workload itself (Instance_generate.py distributions at full size, scale = 1.0) on a batch of
    more than three rounds of the step kernel's warp slots (multi-round packing, TMA staging with the next
    round's prefetch, LP servers under load), from reset() through whole episodes into the next ones: a sample
    of environments is compared with the oracle output for output, every launch."""
    import torch
    import oracle_py
    from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance
    n_inst, T, launches = 96, 64, 44                      # 2816 steps: episodes of this profile take ~1500-3500
    insts = [FJSPInstance.generate(7000 + i, [0.5, 1.0, 1.5][i % 3], 10, 3, "DA3C") for i in range(n_inst)]
    blobs = [i.to_blob() for i in insts]
    B = 3 * 148 * 32 + 777                                # > 3 full rounds whatever the CTA shape
    env_instance = (np.arange(B) % n_inst).astype(np.int32)
    vec = make_vec(blobs, env_instance, "MO_DFJSP")
    q = vec.query()
    assert q["n_slots"] >= 3 * (q["grid"] - q["lp_server_ctas"]) * q["env_warps"]
    sample = np.unique(np.concatenate([np.linspace(0, B - 1, 10).astype(np.int64), [1, B // 2 + 5]]))
    envs = [oracle_py.OracleEnv(blobs[env_instance[e]], "MO_DFJSP") for e in sample]
    s0 = vec.reset().cpu().numpy()
    pc.assert_states_close(s0[sample], np.stack([e.reset() for e in envs]), "reset")
    rng = np.random.default_rng(12)
    dev = vec.dev
    idx = torch.from_numpy(sample).to(dev)
    episodes = 0
    for L in range(launches):
        actions = np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32)
        o = vec.rollout(torch.from_numpy(actions).to(dev), torch.from_numpy(rnd.view(np.int32)).to(dev), reward_policy=1, want_rec=True)
        ref = oracle_py.batch_rollout(envs, actions[:, sample], rnd[:, sample], 1)
        assert np.array_equal(o["rec"][:, idx].cpu().numpy(), ref["rec"]), f"launch {L}: schedule records differ"
        assert np.array_equal(o["done"][:, idx].cpu().numpy(), ref["done"]), f"launch {L}"
        assert np.array_equal(o["reward"][:, idx].cpu().numpy(), ref["reward"]), f"launch {L}"
        pc.assert_states_close(o["state"][:, idx].cpu().numpy(), ref["state"], f"launch {L}")
        episodes += int(ref["done"].sum())
    info = vec.info()
    assert (info["error"] == 0).all()
    assert episodes >= len(sample) // 2, "the run must carry the sample through whole episodes"
    assert np.array_equal(info["step_time"][sample], [e.info()["step_time"] for e in envs])


@pytest.mark.parametrize("knobs", [{"FJSP_LP_SLOTS": "5"}, {"FJSP_NO_CTA_LP": "1"}, {"FJSP_LP_SLOTS": "3", "FJSP_NO_CTA_LP": "1"},
                                   {"FJSP_LP_SERVERS": "1", "FJSP_LP_GROUPS": "1"}, {"FJSP_NO_STAGE": "1"},
                                   {"FJSP_SRV_JOIN": "1", "FJSP_LP_SERVERS": "1"}])
def test_fallback_paths_vs_oracle(knobs, monkeypatch):
    """The paths a default run does not take: fewer LP solution slots than environments (an env without a
    slot solves its LP in line on its warp's scratch slab, in reset() and in the resume kernel; no cached
    order-0 solution for some instances), parking for the LP / resume kernels instead of the LP servers, a
    single one-group server, no shared-memory staging, and env CTAs that serve LPs once their envs are done."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    pc.compare_with_oracle(make_vec, "MO_DFJSP", 23, n_inst=6, copies=4, T=48, launches=3, reward_policy=1)


def test_pipelined_host_calls_equal_blocking_calls():
    """fjsp_vec_step_host_begin / _wait (two calls in flight, outputs written by the kernel into page-locked
    host memory) deliver exactly what the blocking fjsp_vec_step_host delivers."""
    import torch
    from deep_reinforcement_learning_for_fjsp_b200 import _lib
    insts, env_instance = pc.random_batch(33, "MO_DFJSP", 5, 30)
    blobs = [i.to_blob() for i in insts]
    a, b = make_vec(blobs, env_instance, "MO_DFJSP"), make_vec(blobs, env_instance, "MO_DFJSP")
    B, T, calls = len(env_instance), 24, 5
    rng = np.random.default_rng(2)
    acts = [np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32) for _ in range(calls)]
    rnds = [rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32) for _ in range(calls)]
    assert np.array_equal(a.reset_host(), b.reset_host())
    want = [a.step_host(acts[k], rnds[k], 1) for k in range(calls)]
    L = b._L
    pa = [torch.from_numpy(x).pin_memory() for x in acts]
    pr = [torch.from_numpy(x.view(np.int32)).pin_memory() for x in rnds]
    st = [torch.empty((T, B, b.state_size), dtype=torch.float64).pin_memory() for _ in range(calls)]
    rw = [torch.empty((T, B), dtype=torch.float64).pin_memory() for _ in range(calls)]
    dn = [torch.empty((T, B), dtype=torch.int32).pin_memory() for _ in range(calls)]
    rec = [torch.empty((T, B, 8), dtype=torch.int32).pin_memory() for _ in range(calls)]

    def begin(k):
        _lib.check(L.fjsp_vec_step_host_begin(b._h, T, pa[k].data_ptr(), pr[k].data_ptr(), 1, 1.0, 1.0, 1.0, 1,
                                              st[k].data_ptr(), None, rw[k].data_ptr(), dn[k].data_ptr(), rec[k].data_ptr()))
    begin(0)
    for k in range(1, calls):
        begin(k)
        _lib.check(L.fjsp_vec_step_host_wait(b._h))
        assert np.array_equal(rec[k - 1].numpy(), want[k - 1][3]) and np.array_equal(st[k - 1].numpy(), want[k - 1][0])
    with pytest.raises(RuntimeError):
        begin(0); begin(1)                  # a third call in flight is refused
    _lib.check(L.fjsp_vec_step_host_wait(b._h))
    _lib.check(L.fjsp_vec_step_host_wait(b._h))
    assert np.array_equal(rw[calls - 1].numpy(), want[calls - 1][1]) and np.array_equal(dn[calls - 1].numpy(), want[calls - 1][2])
    assert np.array_equal(rec[calls - 1].numpy(), want[calls - 1][3])


@pytest.mark.parametrize("T", [16, 43, 600])
def test_host_output_routes_agree(T, monkeypatch):
    """fjsp_vec_step_host delivers its outputs by one of three routes: kernel stores into page-locked host memory
    (the default for such buffers), chunks copied out while the kernel runs (progress words polled by the host: other
    buffers from T = 16 on), device staging copied after the launch (FJSP_PROGRESSIVE=0).  All three must deliver the same
    bytes -- on a batch of more than one round of warp slots, with page-locked and with pageable buffers, for a
    ragged last chunk (T = 43) and for more steps than progress words at 8 steps per chunk (T = 600)."""
    import torch
    from deep_reinforcement_learning_for_fjsp_b200 import _lib
    copies = 1300 if T < 100 else 40
    insts, env_instance = pc.random_batch(35, "MO_DFJSP", 4, copies)
    blobs = [i.to_blob() for i in insts]
    B = len(env_instance)
    rng = np.random.default_rng(3)
    calls = 3
    acts = [np.stack([rng.integers(0, 12, (T, B)), rng.integers(0, 10, (T, B))], -1).astype(np.int32) for _ in range(calls)]
    rnds = [rng.integers(0, 2**32, (T, B, 2), dtype=np.uint64).astype(np.uint32) for _ in range(calls)]
    got = {}
    for route, knobs in (("progressive", {"FJSP_ZEROCOPY": "0"}), ("zero_copy", {}), ("staged", {"FJSP_ZEROCOPY": "0", "FJSP_PROGRESSIVE": "0"})):
        for k in ("FJSP_ZEROCOPY", "FJSP_PROGRESSIVE"):
            monkeypatch.delenv(k, raising=False)
        for k, x in knobs.items():
            monkeypatch.setenv(k, x)
        vec = make_vec(blobs, env_instance, "MO_DFJSP")
        vec.reset_host()
        out = []
        for k in range(calls):
            pa, pr = torch.from_numpy(acts[k]).pin_memory(), torch.from_numpy(rnds[k].view(np.int32)).pin_memory()
            st = torch.full((T, B, vec.state_size), -7.0, dtype=torch.float32).pin_memory()
            rw = torch.full((T, B), -7.0, dtype=torch.float64).pin_memory()
            dn = torch.full((T, B), -7, dtype=torch.int32).pin_memory()
            rec = torch.full((T, B, 8), -7, dtype=torch.int32).pin_memory()
            _lib.check(vec._L.fjsp_vec_step_host(vec._h, T, pa.data_ptr(), pr.data_ptr(), 2, 1.0, 1.0, 1.0, 1,
                                                 None, st.data_ptr(), rw.data_ptr(), dn.data_ptr(), rec.data_ptr()))
            out.append([x.numpy().copy() for x in (st, rw, dn, rec)])
        # pageable buffers, float64 state
        out.append(list(vec.step_host(acts[0], rnds[0], 2)))
        assert (vec.info()["error"] == 0).all()
        got[route] = out
        vec.close()
    for route in ("zero_copy", "staged"):
        for k in range(calls + 1):
            for a, b in zip(got["progressive"][k], got[route][k]):
                assert np.array_equal(a, b), (route, k)
    assert (got["progressive"][0][2] >= 0).all() and not (got["progressive"][0][0] == -7.0).all(axis=-1).any()


@pytest.mark.parametrize("variant", ["SO_DFJSP", "MO_DFJSP"])
def test_reset_of_a_used_environment(variant):
    pc.check_reset_of_used_env(make_vec, variant)


def test_input_the_reference_raises_on():
    """Two copies of this batch restart an episode with a machine still flagged busy (reset() of a used object,
    DESIGN.md section 1) and the first rule finds nothing to dispatch: the reference raises there.  Both sides
    must flag exactly those copies (error 2 at step 0 of the new episode) and agree on every other copy."""
    vec, envs = pc.compare_with_oracle(make_vec, "MO_DFJSP_breakdown", 1908, n_inst=7, copies=80, T=32, launches=4,
                                       reward_policy=2, breakdowns=True)
    info = vec.info()
    bad = np.nonzero(info["error"])[0]
    assert len(bad) == 2 and (info["error"][bad] == 2).all() and (info["step_count"][bad] == 0).all()
