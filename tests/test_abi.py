"""CPU tier: the C-ABI library loads, exports every symbol include/scs_b200.h declares, its
struct layouts match the reference's, and the Python front end mirrors the reference's
validation / error behaviour (no compute calls: there is no GPU in this tier)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from scs_python_b200 import _scs_b200 as B
    return B


def test_every_declared_symbol_is_exported():
    B = _lib()
    hdr = open(os.path.join(ROOT, "include", "scs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    names = set(re.findall(r"\b(scs(?:_\w+)?)\s*\(", hdr))
    names = {n for n in names if not n.startswith("scs_int") and not n.startswith("scs_float")}
    assert {"scs_init", "scs_solve", "scs_update", "scs_finish", "scs", "scs_set_default_settings", "scs_version",
            "scs_init_lin_sys_work", "scs_solve_lin_sys", "scs_update_lin_sys_diag_r", "scs_free_lin_sys_work",
            "scs_get_lin_sys_method", "scs_b200_proj_dual_cone"} <= names
    for n in sorted(names):
        assert hasattr(B.lib, n), "libscsb200.so does not export %s" % n


def test_struct_layouts_match_reference_abi():
    """sizes/offsets printed by a C program compiled against the reference's own scs.h
    (non-DLONG, non-SFLOAT; tests/golden/make_golden.py:reference_abi) vs the ctypes mirror; the
    C++ side includes the same include/scs_b200.h declarations."""
    from tests import helpers
    abi = helpers.golden("kat.json")["abi"]
    B = _lib()
    for name, val in abi.items():
        if name.startswith("sizeof_scs_"):
            continue
        if name.startswith("sizeof_"):
            assert C.sizeof(getattr(B, name[len("sizeof_"):])) == val, name
        else:
            _, T, f = name.split("_", 2)
            assert getattr(getattr(B, T), f).offset == val, name
    assert abi["sizeof_scs_int"] == 4 and abi["sizeof_scs_float"] == 8


def test_defaults_and_version():
    B = _lib()
    st = B.ScsSettings()
    B.lib.scs_set_default_settings(C.byref(st))
    assert (st.max_iters, st.eps_abs, st.eps_rel, st.eps_infeas) == (100000, 1e-4, 1e-4, 1e-7)  # glbopts.h:35-41
    assert (st.alpha, st.rho_x, st.scale, st.normalize, st.adaptive_scale) == (1.5, 1e-6, 0.1, 1, 1)
    assert (st.acceleration_lookback, st.acceleration_interval, st.acceleration_type_1) == (10, 10, 1)
    assert (st.acceleration_regularization, st.acceleration_relaxation) == (1e-8, 1.0)
    assert B.version() == "3.2.11" and B.sizeof_int() == 4 and B.sizeof_float() == 8
    assert B.lib.scs_get_lin_sys_method().decode().startswith("sparse-indirect-b200")


def _tiny():
    A = sp.csc_matrix(np.array([[1.0, 1.0], [1.0, -1.0]]))
    return dict(A=A, b=np.array([3.0, 1.0]), c=np.array([1.0, 1.0])), dict(z=2)


def test_front_end_validation_mirrors_reference():
    import scs_python_b200 as scs
    data, cone = _tiny()
    with pytest.raises(ValueError):
        scs.SCS({}, cone)
    with pytest.raises(ValueError):
        scs.SCS(dict(A=data["A"], b=data["b"]), cone)
    with pytest.raises(TypeError):
        scs.SCS(dict(A=data["A"].toarray(), b=data["b"], c=data["c"]), cone)
    with pytest.raises(ValueError):
        scs.SCS(dict(A=data["A"], b=data["b"][:1], c=data["c"]), cone)
    for bad in (dict(max_iters=0), dict(alpha=2.5), dict(scale=-1.0), dict(eps_abs=float("nan")),
                dict(acceleration_lookback=-1), dict(rho_x=0.0), dict(time_limit_secs=-1.0)):
        with pytest.raises(ValueError):
            scs.SCS(data, cone, verbose=False, **bad)
    with pytest.raises(TypeError):
        scs.SCS(data, cone, verbose=1)            # bool settings must be bool (scsobject.h:537-539)
    with pytest.raises(TypeError):
        scs.SCS(data, cone, not_a_setting=3)
    with pytest.raises(ValueError):
        scs.SCS(data, dict(z=2, q=[-1]), verbose=False)
    with pytest.raises(ImportError):
        scs.SCS(data, cone, linear_solver=scs.LinearSolver.QDLDL)
    assert scs.LinearSolver("b200") is scs.LinearSolver.B200
    assert (scs.SOLVED, scs.INFEASIBLE, scs.UNBOUNDED, scs.FAILED) == (1, -2, -1, -4)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device scs_init must fail loudly (NULL -> ValueError), never compute on CPU."""
    B = _lib()
    if B.lib.scs_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import scs_python_b200 as scs
    data, cone = _tiny()
    with pytest.raises(ValueError, match="ScsWork allocation error"):
        scs.SCS(data, cone, verbose=False)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scs_python_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "scs_oracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


def test_upstream_patch_adds_linear_solver_b200():
    """integration/b200_upstream.patch applied to the reference's front end (make -C oracle ref_frontend_b200):
    `scs.LinearSolver.B200` exists and resolves to `scs._scs_b200`, the reference's scspy.c compiled with
    -DPY_B200 and linked against libscsb200.so (32-bit scs_int).  Runs in a subprocess: the patched package is
    also called `scs`.  No device needed (module import and the three module functions only)."""
    import subprocess
    pkg = os.path.join(ROOT, "oracle", "_ref", "scs_b200")
    if not os.path.exists(os.path.join(pkg, "scs", "__init__.py")):
        pytest.skip("oracle/_ref/scs_b200 not built (make -C oracle ref_frontend_b200)")
    code = ("import sys; sys.path.insert(0, %r); import scs; m = scs._SOLVER_DISPATCH[scs.LinearSolver.B200](); "
            "import os; print(scs.LinearSolver.B200.value, os.path.basename(m.__file__).split('.')[0], m.sizeof_int(), m.sizeof_float(), m.version())" % pkg)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-1500:]
    val, name, si, sf, ver = r.stdout.split()
    assert val == "b200" and name == "_scs_b200" and si == "4" and sf == "8" and ver == "3.2.11"
    patch = open(os.path.join(ROOT, "integration", "b200_upstream.patch")).read()
    for needle in ("link_b200", "PY_B200", "PyInit__scs_b200", 'B200 = "b200"', "_scs_b200"):
        assert needle in patch
