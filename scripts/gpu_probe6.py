"""Probe 6: stream-K head + specialised producer loop; numerics on split-heavy shapes, timings."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 21, "M": 1024, "N": 1024, "K": 1024})
    run_case({"kind": "numerics", "id": 21, "M": 1024, "N": 768, "K": 1000, "alpha": 0.75, "beta": -1.5})
    run_case({"kind": "numerics", "id": 6, "M": 2048, "N": 2048, "K": 512})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    run_case({"kind": "numerics", "id": 31, "M": 2048, "N": 2048, "K": 2048, "oracle": False, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 31, "M": 4096, "N": 4096, "K": 4096, "oracle": False, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 16, "M": 1056, "N": 1120, "K": 520, "selftest": [10000.0, 17, 5]})
    for n in (1024, 2048, 3072, 4096, 6144, 8192):
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 6, 5, 21, 16, 15, 31], "reps": reps, "tag": "sk"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [5, 21, 31], "reps": reps, "dbg": {"splitk": 0}, "tag": "no-sk"}, timeout=600)
    for n in (4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": 10 if n <= 4096 else 4, "reuse": 1, "tag": "reuse-encode"}, timeout=600)

if __name__ == "__main__":
    main()
