"""Probe 7: new double-float encode, FT dissection after the producer fix, plain vs cuBLAS at 8192."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    run_case({"kind": "numerics", "id": 16, "M": 1056, "N": 1120, "K": 520, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 8192, "tau_abs": 1e9, "oracle": False})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "dist": "normal", "tau_abs": 1e9})
    for n in (4096, 8192):
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "tag": "full"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "tag": "reuse-encode"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 1}, "tag": "no-check"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 2}, "tau_abs": 1e30, "tag": "no-chk-tiles"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 3}, "tag": "neither"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"splitk": 0}, "tag": "no-split"}, timeout=600)

if __name__ == "__main__":
    main()
