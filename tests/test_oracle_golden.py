"""The CPU oracle must reproduce, bit for bit, what the UNMODIFIED reference
environments returned for the committed golden trajectories (tests/golden/*.npz,
made by oracle/make_golden.py): schedules (operation, job, machine, begin, end),
float64 states, rewards, done flags, including re-reset of a used environment."""
import os

import numpy as np
import pytest

import oracle_py
from conftest import golden_cases


def replay(case, golden_dir, sum_mode=1):
    g = np.load(os.path.join(golden_dir, case + ".npz"), allow_pickle=False)
    variant = str(g["variant"])
    env = oracle_py.OracleEnv(g["blob"], variant, sum_mode)
    actions, rnd, rp = g["actions"], g["rnd"], int(g["reward_policy"])
    t, T = 0, len(actions)
    for ep in range(len(g["resets"])):
        s0 = env.reset()
        assert np.array_equal(s0, g["resets"][ep]), f"reset state, episode {ep}"
        while t < T:
            st, rw, dn, rec = env.step(actions[t], rnd[t], rp)
            assert np.array_equal(rec, g["recs"][t]), f"schedule record at step {t}"
            assert np.array_equal(st, g["states"][t]), f"state at step {t}"
            assert rw == g["rewards"][t], f"reward at step {t}"
            assert int(dn) == int(g["dones"][t]), f"done at step {t}"
            t += 1
            if dn:
                info = env.info()
                ref = g["infos"][ep]
                assert info["step_time"] == ref[0] and info["step_count"] == ref[1]
                assert info["delay_sum"] == ref[3]
                if variant.startswith("MO"):
                    assert info["completion"] == ref[2] and info["energy"] == ref[4]
                assert int(env.machine_end().max()) == ref[5]
                break
    assert t == T
    return env


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_reproduces_reference(case, golden_dir):
    replay(case, golden_dir)


def test_naive_sum_mode_differs_only_in_floats(golden_dir):
    """sum_mode=0 (CPython <= 3.11 left-to-right sum) is a different float stream; it must
    still run to completion without error flags on a golden instance."""
    g = np.load(os.path.join(golden_dir, "so_dfjsp_small_a.npz"))
    env = oracle_py.OracleEnv(g["blob"], "SO_DFJSP", 0)
    env.reset()
    done, n = False, 0
    while not done:
        _, _, done, _ = env.step((n % 5, n % 4))
        n += 1
    assert env.info()["error"] == 0 and n == env.info()["step_count"]
