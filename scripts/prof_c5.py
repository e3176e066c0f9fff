"""C5 workload run (block QL, 512 energies) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ila_b200
xs = np.concatenate([np.linspace(-0.95, 0.95, 256), np.linspace(2.25, 4.0, 256)])   # the bench.py C5 energies
for _ in range(3):
    L11, iters, st = ila_b200.ql_blocktri_l11(*ila_b200.periodic_schrodinger_blocks(), energies=xs)
print("ok", int(iters.sum()), int((st != 0).sum()))
