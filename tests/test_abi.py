"""The C-ABI shared library loads on a machine without a GPU and exports every function that
include/fjsp_b200.h declares; without a CUDA device it refuses to create an environment
(there is no CPU fallback in the product)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from deep_reinforcement_learning_for_fjsp_b200 import build, _lib
    build.build_cuda()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fjsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fjsp_[a-z_0-9]+)\s*\(", text)))


def test_header_functions_are_exported(lib):
    names = declared_functions()
    assert {"fjsp_vec_create", "fjsp_vec_reset", "fjsp_vec_step", "fjsp_vec_step_host", "fjsp_vec_info",
            "fjsp_vec_destroy", "fjsp_vec_query", "fjsp_vec_reset_host", "fjsp_last_error",
            "fjsp_abi_version"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fjsp_b200.h but not exported"
    assert lib.fjsp_abi_version() == 2


def test_create_argument_errors(lib):
    h = ctypes.c_void_p()
    assert lib.fjsp_vec_create(None, None, 0, None, 0, 1, 1, 0, ctypes.byref(h)) != 0
    assert b"null or empty" in lib.fjsp_last_error()
    blob = np.zeros(64, np.int32)
    offs = np.zeros(1, np.int64)
    ei = np.zeros(1, np.int32)
    assert lib.fjsp_vec_create(blob.ctypes.data, offs.ctypes.data, 1, ei.ctypes.data, 1, 9, 1, 0, ctypes.byref(h)) != 0
    assert b"variant" in lib.fjsp_last_error()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from deep_reinforcement_learning_for_fjsp_b200.instance import FJSPInstance
    blob = FJSPInstance.generate(1, 1.0, 5, 1, "DA3C", scale=0.1).to_blob()
    offs = np.zeros(1, np.int64)
    ei = np.zeros(1, np.int32)
    h = ctypes.c_void_p()
    rc = lib.fjsp_vec_create(blob.ctypes.data, offs.ctypes.data, 1, ei.ctypes.data, 1, 0, 1, 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.fjsp_last_error()
    from deep_reinforcement_learning_for_fjsp_b200.vec_env import FJSPVecEnv
    with pytest.raises(RuntimeError):
        FJSPVecEnv(None, [0], "SO_DFJSP", blobs=[blob])


def test_product_never_imports_the_oracle():
    """No Python import and no #include in the package reaches oracle/ or tests/."""
    pkg = os.path.join(ROOT, "deep_reinforcement_learning_for_fjsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for line in open(path):
                    if re.match(r"\s*(import|from)\s", line):
                        assert "oracle" not in line and "hostsim" not in line and "tests" not in line, (f, line)
            elif f.endswith((".cu", ".cuh", ".h")):
                for line in open(path):
                    if line.lstrip().startswith("#include"):
                        assert "oracle" not in line and "tests/" not in line, (f, line)
