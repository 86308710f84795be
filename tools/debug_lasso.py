import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scs_python_b200 as scsb
from scs_python_b200 import problems
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n0 = int(1_000_000 * scale)
data, cone, aux = problems.lasso(n0, 2 * n0, 100, 0)
eps_infeas = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-7
s = scsb.SCS(data, cone, verbose=True, max_iters=iters, eps_infeas=eps_infeas)
sol = s.solve(warm_start=False)
print({k: v for k, v in sol["info"].items() if k != "aa_stats"})
