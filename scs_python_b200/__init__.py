"""scs_python_b200 -- B200-native (sm_100a) backend for the SCS ADMM iteration.

Python-side mirror of the reference front end `scs/py/__init__.py` for the one backend this
package provides: `scs.SCS(data, cone, linear_solver=LinearSolver.B200, **settings)`.
Same validation, CSC / upper-triangular preparation of A and P, `solve(warm_start, x, y, s)`,
`update(b, c)`, legacy `solve(data, cone, **settings)` and status constants
(scs/py/__init__.py:14-25, 87-230).  The only extension module behind it is `_scs_b200`
(libscsb200.so); selecting any other LinearSolver raises ImportError exactly as the
reference does for a backend that was not built.

INTEGRATION.md shows the three-line patch that adds `LinearSolver.B200` and the
`-Dlink_b200=true` meson target to the reference itself.
"""
from __future__ import annotations

import enum
import warnings

import numpy as np
from scipy import sparse

# SCS return codes (scs/py/__init__.py:14-25 ; S/include/scs.h:33-42)
INFEASIBLE_INACCURATE = -7
UNBOUNDED_INACCURATE = -6
SIGINT = -5
FAILED = -4
INDETERMINATE = -3
INFEASIBLE = -2
UNBOUNDED = -1
UNFINISHED = 0
SOLVED = 1
SOLVED_INACCURATE = 2


class LinearSolver(enum.Enum):
    """Linear system solver backend (scs/py/__init__.py:28-37 plus the new member)."""
    AUTO = "auto"
    QDLDL = "qdldl"
    CPU_INDIRECT = "cpu_indirect"
    MKL = "mkl"
    ACCELERATE = "accelerate"
    CPU_DENSE = "cpu_dense"
    GPU_INDIRECT = "gpu_indirect"
    CUDSS = "cudss"
    B200 = "b200"


def _load_b200():
    from . import _scs_b200  # raises ImportError loudly when libscsb200.so is missing
    return _scs_b200


def _select_scs_module(stgs):
    linear_solver = stgs.pop("linear_solver", LinearSolver.B200)
    if isinstance(linear_solver, str):
        linear_solver = LinearSolver(linear_solver)
    if linear_solver in (LinearSolver.B200, LinearSolver.AUTO):
        return _load_b200()
    raise ImportError("scs_python_b200 only ships the B200 backend; "
                      "%s is provided by the reference package `scs`" % linear_solver)


def _has_lower_tri(P):
    nnz_per_col = np.diff(P.indptr)
    nonempty = nnz_per_col > 0
    if not nonempty.any():
        return False
    last_row = P.indices[P.indptr[1:][nonempty] - 1]
    return bool(np.any(last_row > np.where(nonempty)[0]))


def _prepare(data, cone):
    """Validation and CSC / upper-triangle conversion of scs.SCS.__init__ (scs/py/__init__.py:89-184);
    returns the positional arguments of the extension type's constructor."""
    if not data or not cone:
        raise ValueError("Missing data or cone information")
    if "b" not in data or "c" not in data:
        raise ValueError("Missing one of b, c from data dictionary")
    if "A" not in data:
        raise ValueError("Missing A from data dictionary")
    A, b, c = data["A"], data["b"], data["c"]
    if A is None or b is None or c is None:
        raise ValueError("Incomplete data specification")
    if not sparse.issparse(A):
        raise TypeError("A is required to be a sparse matrix")
    if not A.format == "csc":
        warnings.warn("Converting A to a CSC (compressed sparse column) matrix; may take a while.")
        A = A.tocsc()
    if sparse.issparse(b):
        b = np.asarray(b.todense()).ravel()
    if sparse.issparse(c):
        c = np.asarray(c.todense()).ravel()
    m, n = len(b), len(c)
    if not A.has_sorted_indices:
        A = A.sorted_indices()
    if A.shape != (m, n):
        raise ValueError("A shape not compatible with b,c")
    Pdata = Pindices = Pcolptr = None
    P = data.get("P", None)
    if P is not None:
        if not sparse.issparse(P):
            raise TypeError("P is required to be a sparse matrix")
        if P.shape != (n, n):
            raise ValueError("P shape not compatible with A,b,c")
        if not P.format == "csc":
            warnings.warn("Converting P to a CSC (compressed sparse column) matrix; may take a while.")
            P = P.tocsc()
        if not P.has_sorted_indices:
            P = P.sorted_indices()
        if _has_lower_tri(P):
            P = sparse.triu(P, format="csc")
        Pdata, Pindices, Pcolptr = P.data, P.indices, P.indptr
    return ((m, n), A.data, A.indices, A.indptr, Pdata, Pindices, Pcolptr, np.asarray(b), np.asarray(c), cone)


class SCS(object):
    def __init__(self, data, cone, **settings):
        """Same contract as scs.SCS.__init__ (scs/py/__init__.py:89-184)."""
        self._settings = settings
        args = _prepare(data, cone)
        _scs = _select_scs_module(self._settings)
        self._solver = _scs.SCS(*args, **self._settings)

    def solve(self, warm_start=True, x=None, y=None, s=None):
        """scs/py/__init__.py:186-203."""
        return self._solver.solve(warm_start, x, y, s)

    def update(self, b=None, c=None):
        """scs/py/__init__.py:205-214."""
        self._solver.update(b, c)


def dist_init(rank=None, world=None):
    """Row-partitioned mode (one process per GPU, e.g. under torchrun): call once per process
    after selecting the GPU.  The NCCL id is shipped with torch.distributed when it is
    initialised (any backend), which is plumbing only -- the data path is NCCL inside
    libscsb200.so."""
    import os
    mod = _load_b200()
    rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
    if world == 1:
        mod.dist_init(0, 1, b"")
        return
    import torch.distributed as td
    if not td.is_initialized():
        raise RuntimeError("dist_init: initialise torch.distributed first (it only carries the 128-byte NCCL id)")
    box = [mod.dist_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    mod.dist_init(rank, world, box[0])


def dist_finalize():
    _load_b200().dist_finalize()


def solve_batch(problems, **settings):
    """Solve independent problems in one call: `problems` is a sequence of (data, cone) pairs with
    the meaning `SCS(data, cone)` gives them; every problem uses the same settings and a cold start.
    Small problems with zero / nonneg / second-order cones run in the batch engine (one CTA per
    problem, one kernel launch for the whole batch; BASELINE.json configs[4]); anything else is
    solved by the streaming engine one after another.  The reference API is single-problem
    (scs/py/__init__.py:89-230), so this entry point has no reference counterpart; each result is
    what `SCS(data, cone, **settings).solve(warm_start=False)` returns.  Shard a batch across GPUs by
    giving rank r the problems r, r + world, ... (no communication)."""
    settings = dict(settings)
    _scs = _select_scs_module(settings)
    return _scs.solve_batch([_prepare(d, k) for d, k in problems], **settings)


def read_data(filename):
    """Read a problem file in the reference's binary layout (written by `write_data_filename=...` of either
    library, or S/test/problems/*): returns (data, cone, settings) ready for `SCS(data, cone, **settings)`.
    Reference: SCS(read_data), S/src/rw.c:262-315, used by S/test/run_from_file.c."""
    return _load_b200().read_data(filename)


def write_data(filename, data, cone, **settings):
    """Write (data, cone, settings) in that layout without building a workspace (SCS(write_data),
    S/src/rw.c:240-260).  `SCS(data, cone, write_data_filename=...)` does the same at construction."""
    settings = dict(settings)
    settings.pop("linear_solver", None)
    return _load_b200().write_data(filename, *_prepare(data, cone), **settings)


def solve(data, cone, **settings):
    """Legacy one-shot API (scs/py/__init__.py:217-230)."""
    solver = SCS(data, cone, **settings)
    return solver.solve(warm_start=True, x=data.get("x"), y=data.get("y"), s=data.get("s"))


def __getattr__(name):
    if name in ("__version__", "__sizeof_int__", "__sizeof_float__"):
        mod = _load_b200()
        return {"__version__": mod.version(), "__sizeof_int__": mod.sizeof_int(),
                "__sizeof_float__": mod.sizeof_float()}[name]
    raise AttributeError(name)
