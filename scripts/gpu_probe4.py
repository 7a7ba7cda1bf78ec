"""Probe 4: 3-D single-instruction TMA stage loads; numerics (incl. 2-D fallback shapes) + timings."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 21, "M": 1024, "N": 768, "K": 1000, "alpha": 0.75, "beta": -1.5})
    run_case({"kind": "numerics", "id": 5, "M": 512, "N": 512, "K": 512})
    run_case({"kind": "numerics", "id": 6, "M": 200, "N": 136, "K": 100})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    run_case({"kind": "numerics", "id": 16, "M": 1056, "N": 1120, "K": 520, "selftest": [10000.0, 17, 5]})
    for n in (1024, 2048, 4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 1, 2, 6, 5, 21, 22, 12, 16, 15, 31, 32], "reps": 10 if n <= 4096 else 4}, timeout=600)
    run_case({"kind": "timing", "M": 16384, "N": 16384, "K": 16384, "ids": [7, 21, 31], "reps": 2}, timeout=600)
    for n in (4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [5, 21, 31], "reps": 10 if n <= 4096 else 4, "dbg": {"tma3d": 0}, "tag": "2d"}, timeout=600)

if __name__ == "__main__":
    main()
