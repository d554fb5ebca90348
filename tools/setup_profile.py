"""Where does osqp_setup time go? (development aid)"""
import sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
import ctypes as C
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels
k = kernels(); assert k.b200_init(0) == 0
pb = problems.lasso(int(1e5), int(1e6))
kw = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5, verbose=0)
for rep in range(3):
    t0 = time.perf_counter()
    s = OSQP()
    P = sp.triu(sp.csc_matrix(pb["P"]), format="csc"); A = sp.csc_matrix(pb["A"])
    t1 = time.perf_counter()
    s.setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
    t2 = time.perf_counter()
    print(f"rep {rep}: python csc prep {1e3*(t1-t0):.0f} ms; OSQP.setup total {1e3*(t2-t1):.0f} ms; C-level info.setup_time {1e3*s.info.setup_time:.0f} ms")
    s.cleanup()
# scaling=0 to see Ruiz cost
for sc in (0, 10):
    s = OSQP(); t1 = time.perf_counter(); s.setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], scaling=sc, **kw); t2 = time.perf_counter()
    print(f"scaling={sc}: setup {1e3*(t2-t1):.0f} ms (C {1e3*s.info.setup_time:.0f} ms)"); s.cleanup()
