"""Run one kernel id at one size a few times (target for ncu / compute-sanitizer).
usage: run_one.py ID N [reps] [key=value debug knobs ...]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "scripts"))
import __graft_entry__ as ge
import cuda_rt as cu
pkg = ge.load_package()
kid, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 and '=' not in sys.argv[3] else 3
for a in sys.argv[3:]:
    if '=' in a:
        pkg.debug_set(a.split('=')[0], int(a.split('=')[1]))
rng = np.random.default_rng(0)
A = (rng.integers(-9, 10, n * n) * 0.1).astype(np.float32)
B = (rng.integers(-9, 10, n * n) * 0.1).astype(np.float32)
dA, dB, dC = cu.DevBuf.from_numpy(A), cu.DevBuf.from_numpy(B), cu.DevBuf(4 * n * n)
dC.zero()
ft = pkg.FtSgemm()
for _ in range(reps):
    ft.run(kid, n, n, n, dA, dB, dC, 1.0, 0.0, None)
cu.sync()
print("done", kid, n)
