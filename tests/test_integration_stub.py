"""The ctypes stub INTEGRATION.md shows a maintainer of the reference is executed VERBATIM
(extracted from the markdown) and stepped through the C ABI next to the CPU oracle: SO
environments (no reward_policy argument) and MO environments, flat and pair actions."""
import os
import re

import numpy as np
import pytest

import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert blocks, "INTEGRATION.md lost its stub"
    return blocks[0]


def test_stub_declares_every_restype():
    src = stub_source()
    for fn in ("fjsp_vec_create", "fjsp_vec_reset_host", "fjsp_vec_step_host", "fjsp_vec_info", "fjsp_last_error"):
        assert re.search(fn + r"\.(argtypes, _L\." + fn + r"\.)?restype", src), fn
    assert "reward_policy or 0" in src     # SO environments call step(action) without it
    compile(src, "INTEGRATION.md", "exec")


@pytest.mark.gpu
@pytest.mark.parametrize("cls,variant,rp", [("SO_DFJSP_Environment", "SO_DFJSP", None), ("MO_DFJSP_Environment", "MO_DFJSP", 1),
                                            ("MO_DFJSP_breakdown_Environment", "MO_DFJSP_breakdown", 2),
                                            ("SO_FJSSP_Environment", "SO_FJSSP", None)])
def test_stub_steps_through_the_c_abi(cls, variant, rp):
    import oracle_py
    ns = {}
    exec(compile(stub_source(), "INTEGRATION.md", "exec"), ns)
    insts, _ = pc.random_batch(51, variant, 1, 1, breakdowns=variant.endswith("breakdown"))
    env = ns[cls](instance=insts[0])
    ora = oracle_py.OracleEnv(insts[0].to_blob(), variant)
    pc.assert_states_close(env.reset(), ora.reset(), "reset")
    nt, nm = pc.NRULES[variant]
    rng = np.random.default_rng(7)
    n = 0
    while not env.done and n < 400:
        a = (int(rng.integers(0, nt)), int(rng.integers(0, nm)))
        draws = rng.integers(0, 2**32, (1, 1, 2), dtype=np.uint64).astype(np.uint32)
        act = [a[0] * nm + a[1]] if n % 3 == 0 else a           # the reference accepts the flat index too
        kw = {} if rp is None else dict(reward_policy=rp, completion=1.0, tardiness=1.0, energy_consumption=1.0)
        s, r, d = env.step(act, draws=draws, **kw)
        so, ro, do, _ = ora.step(a, draws[0, 0], rp if rp is not None else 1)
        assert r == ro and d == do, n
        pc.assert_states_close(s, so, f"step {n}")
        n += 1
    assert n > 20
    assert env.delay_time_sum == ora.info()["delay_sum"]
