"""CPU parity of the KERNEL SOURCE: csrc/fjsp_core.cuh compiled by g++ as a one-lane program
(tests/hostsim) must reproduce the golden reference trajectories and agree with the oracle
on random mixed batches.  This covers the compressed device state (per-order counters,
linked queues, rule cache, fluid slots), the set-order emulation, the in-kernel simplex
and the table builder; the 32-lane cooperation is covered by tests/test_gpu_parity.py."""
import sys, os

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as hostsim  # noqa: E402
import parity_common as pc  # noqa: E402
from conftest import golden_cases  # noqa: E402

DEVICE_CASES = golden_cases()


def make_vec(blobs, env_instance, variant, sum_mode=1):
    return hostsim.HostSimVec(blobs, env_instance, variant, sum_mode)


@pytest.mark.parametrize("case", DEVICE_CASES)
def test_golden(case, golden_dir):
    pc.replay_golden(make_vec, golden_dir, case)


@pytest.mark.parametrize("variant,seed,rp,bd", [("SO_DFJSP", 1, 1, False), ("MO_DFJSP", 2, 0, False),
                                                 ("MO_DFJSP", 3, 3, False), ("MO_DFJSP_breakdown", 4, 2, True), ("SO_FJSSP", 5, 1, False)])
def test_random_batch_vs_oracle(variant, seed, rp, bd):
    pc.compare_with_oracle(make_vec, variant, seed, n_inst=4, copies=2, T=40, launches=3, reward_policy=rp,
                           breakdowns=bd)


def test_table_builder_set_order_matches_cpython():
    L = hostsim.lib()
    rng = np.random.default_rng(5)
    out = np.zeros(32, np.int32)
    for _ in range(5000):
        M = int(rng.integers(1, 33))
        tup = np.ascontiguousarray(rng.permutation(M)[:int(rng.integers(1, M + 1))], np.int32)
        k = L.fjsp_hostsim_pyset_order(tup.ctypes.data, len(tup), out.ctypes.data)
        ord_tup = list(set(int(v) for v in tup))
        assert list(out[:k]) == ord_tup
        idle = [m for m in range(M) if rng.random() < 0.5]
        mask = sum(1 << m for m in idle)
        o = np.ascontiguousarray(ord_tup, np.int32)
        k = L.fjsp_hostsim_selectable(mask, o.ctypes.data, len(o), out.ctypes.data)
        assert list(out[:k]) == list(set(idle) & set(int(v) for v in tup))


def test_bad_blob_is_rejected():
    with pytest.raises(RuntimeError):
        hostsim.HostSimVec([np.zeros(32, np.int32)], [0], "SO_DFJSP")


def test_brandimarte_instances(golden_dir):
    """Mk01 / Mk06 / Mk10 of the reference's benchmark data (static, one job per kind, up to 240
    operation types and 475 LP rows): one episode and a half with auto-reset against the oracle."""
    import oracle_py
    z = np.load(os.path.join(golden_dir, "brandimarte_blobs.npz"))
    for name in ("Mk01", "Mk06", "Mk10"):
        blob = z[name]
        vec = make_vec([blob], [0], "SO_DFJSP")
        env = oracle_py.OracleEnv(blob, "SO_DFJSP")
        pc.assert_states_close(vec.reset_host()[0], env.reset(), "reset")
        T = int(blob[4] * 3 // 2)
        rng = np.random.default_rng(1)
        actions = np.stack([rng.integers(0, 6, (T, 1)), rng.integers(0, 5, (T, 1))], -1).astype(np.int32)
        rnd = rng.integers(0, 2**32, (T, 1, 2), dtype=np.uint64).astype(np.uint32)
        st, rw, dn, rec = vec.step_host(actions, rnd, 1)
        ref = oracle_py.batch_rollout([env], actions, rnd, 1)
        assert np.array_equal(rec, ref["rec"]) and np.array_equal(rw, ref["reward"]) and np.array_equal(dn, ref["done"])
        pc.assert_states_close(st, ref["state"], name)
        assert dn.sum() >= 1


@pytest.mark.parametrize("variant", ["SO_DFJSP", "MO_DFJSP", "MO_DFJSP_breakdown", "SO_FJSSP"])
def test_edge_cases(variant):
    pc.check_edge_cases(make_vec, variant)


def test_left_to_right_sum_mode_vs_oracle():
    """sum_mode 0 (CPython <= 3.11 builtin sum) is a separate template instantiation: it must
    agree with the oracle run in the same mode."""
    pc.compare_with_oracle(make_vec, "MO_DFJSP", 9, n_inst=3, copies=2, T=40, launches=2, sum_mode=0)


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


@pytest.mark.parametrize("chunk", range(4))
def test_random_shapes_vs_oracle(chunk):
    """The first draws of tools/fuzz_parity.py (which runs them on the GPU): batch shapes from one copy to a few
    hundred, rollouts of 1 / 7 / 32 / 64 steps, all four environment classes, every reward policy."""
    variants = [("SO_DFJSP", False), ("MO_DFJSP", False), ("MO_DFJSP_breakdown", True), ("SO_FJSSP", False)]
    rng = np.random.default_rng(7)
    for n in range(32):
        n_inst, copies = int(rng.integers(1, 9)), int(rng.choice([1, 3, 17, 80]))
        T, launches = int(rng.choice([1, 7, 32, 64])), int(rng.integers(2, 5))
        if T == 1:
            launches = 40
        rp = int(rng.integers(0, 4))
        if n % 4 != chunk or copies == 80:
            continue
        variant, bd = variants[(n // 3) % 4]
        pc.compare_with_oracle(make_vec, variant, 1001 + n, n_inst=n_inst, copies=copies, T=T, launches=launches,
                               reward_policy=rp, breakdowns=bd)
