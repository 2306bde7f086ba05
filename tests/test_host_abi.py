"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/snp_b200.h declares; host-side tables agree with the reference's recorded values."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, traj_names, load_traj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "snp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from social_navigation_pyenvs_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/snp_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes signature table and header disagree"
    lib = _lib.lib()
    assert lib.snp_abi_version() == 1


def test_ctypes_struct_layouts_match_the_header():
    """Field order / sizes of the ctypes mirrors (a drift here would silently corrupt every call)."""
    from social_navigation_pyenvs_b200 import _lib as L
    assert ctypes.sizeof(L.SnpCrowd) == 4 * 4 + 6 * 8 + 20 * 8 + 2 * 8 + 3 * 4 + 4  # ints, ptrs, params, ptrs, ints, tail padding
    assert L.SnpCrowd.params.offset == 64 and L.SnpCrowd.robot.offset == 224 and L.SnpCrowd.W.offset == 240
    assert L.SnpStepOpts.dt.offset == 24 and L.SnpStepOpts.action.offset == 32 and L.SnpStepOpts.consts.offset == 56
    assert L.SnpStepOpts.time_now.offset == 104 and L.SnpStepOpts.respawn_bounds.offset == 128 and L.SnpStepOpts.robot_params.offset == 152 and ctypes.sizeof(L.SnpStepOpts) == 312
    assert L.SnpLaserArgs.pose.offset == 64 and L.SnpLaserArgs.ranges.offset == 96 and ctypes.sizeof(L.SnpLaserArgs) == 112


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from social_navigation_pyenvs_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.lib()


def test_model_parameter_rows_match_the_reference():
    """model_parameters(title) == Agent.get_parameters(title) as recorded in every golden fixture (agent.py:268-388)."""
    from social_navigation_pyenvs_b200 import model_parameters, SFMS
    import oracle
    seen = set()
    for name in traj_names():
        d = load_traj(name)
        title = SFMS[int(d["type"])]
        seen.add(title)
        assert np.array_equal(model_parameters(title), d["params"][0]), title
        assert np.array_equal(oracle.default_params(title), d["params"][0]), title
    assert seen == set(SFMS)
    with pytest.raises(Exception):
        model_parameters("orca")


def test_pack_walls_matches_reference_segment_order():
    """Vertex lists -> sorted-endpoint NaN-padded segments == what the reference's Obstacle built (obstacle.py:26-32)."""
    from social_navigation_pyenvs_b200 import scenarios
    d = load_traj("walls7_sfm_helbing")
    assert np.array_equal(scenarios.pack_walls(scenarios.EXAMPLE_WALLS), d["walls"], equal_nan=True)


def test_operator_argument_validation_needs_no_gpu():
    from social_navigation_pyenvs_b200 import update_humans_parallel
    with pytest.raises(ValueError):
        update_humans_parallel(-1, np.zeros((2, 13)), np.zeros((2, 1, 2)), None, np.zeros((2, 20)), 0.01, np.zeros(2))
    with pytest.raises(ValueError):
        update_humans_parallel(0, np.zeros((2, 12)), np.zeros((2, 1, 2)), None, np.zeros((2, 20)), 0.01, np.zeros(2))
    with pytest.raises(ValueError):
        update_humans_parallel(0, np.zeros((3, 13)), np.zeros((2, 1, 2)), None, np.zeros((2, 20)), 0.01, np.zeros(2))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under social_navigation_pyenvs_b200/ may reference it."""
    pkg = os.path.join(ROOT, "social_navigation_pyenvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "libsnp_oracle" not in src and "oracle/_build" not in src, f  # no dlopen / link of the checker
