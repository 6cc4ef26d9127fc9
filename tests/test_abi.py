"""CPU checks of the drop-in boundary: libgmpc.so builds, loads and exports every symbol that
include/gmpc.h declares (no compute calls -- there is no GPU here)."""

import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "gmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmpc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    fns = header_functions()
    for name in ("gmpc_create", "gmpc_destroy", "gmpc_set_weights", "gmpc_rollout",
                 "gmpc_objective_grad", "gmpc_plan", "gmpc_plan_host", "gmpc_critic_loss_grad",
                 "gmpc_clip_adam_step", "gmpc_last_error"):
        assert name in fns


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in gmpc.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from gan_mpc_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()


def test_error_path_without_gpu(built_lib):
    """Argument validation happens before any CUDA call; messages come through gmpc_last_error."""
    from gan_mpc_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.gmpc_create(None, ctypes.byref(h)) == -1
    assert b"null" in lib.gmpc_last_error()
    cfg = _lib.Config(3, 1, 5, 4, 200, 3, 128, 10, 0, 1, 1, 0)
    cfg.dyn_hidden = 4096
    assert lib.gmpc_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
    assert b"hidden <= 512" in lib.gmpc_last_error()
    assert lib.gmpc_launch_count(None) == 0


def test_product_never_imports_oracle():
    """The shipped package must not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gan_mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
