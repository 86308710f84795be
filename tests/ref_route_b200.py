"""pytest plugin (-p ref_route_b200) for running the REFERENCE's own test files against the B200 backend.

The reference's tests ask for LinearSolver.AUTO / QDLDL / CPU_INDIRECT.  With the patched front end of
integration/b200_upstream.patch on the path (oracle/_ref/scs_b200, built by `make -C oracle
ref_frontend_b200`), this plugin re-points those three dispatch entries at `scs._scs_b200`, so every solve in
the reference's files goes through scs.LinearSolver.B200 -> libscsb200.so.  Nothing else is touched."""
import scs

assert hasattr(scs.LinearSolver, "B200"), "the patched reference front end is not on the path"
_b200 = scs._SOLVER_DISPATCH[scs.LinearSolver.B200]
for _k in (scs.LinearSolver.AUTO, scs.LinearSolver.QDLDL, scs.LinearSolver.CPU_INDIRECT):
    scs._SOLVER_DISPATCH[_k] = _b200
_mod = _b200()
assert _mod.__name__.endswith("_scs_b200") or _mod.sizeof_int() == 4


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    terminalreporter.write_line("reference tests routed to scs.LinearSolver.B200 (%s, %s)" % (_mod.__file__, _mod.version()))
