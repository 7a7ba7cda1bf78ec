"""Third probe: is the CTA-pair kernel fixed; v1 library vs current on the same box; clocks."""
import os, sys, json, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 21, "M": 1024, "N": 768, "K": 1000, "alpha": 0.75, "beta": -1.5})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    for n in (4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 6, 5, 21, 22, 16, 15, 31], "reps": 10 if n <= 4096 else 4}, timeout=600)
    os.environ["FTSGEMM_LIB"] = str(Path(__file__).resolve().parent / "libftsgemm_v1.so")
    for n in (4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 6, 5, 16, 15], "reps": 10 if n <= 4096 else 4, "tag": "v1lib"}, timeout=600)
    del os.environ["FTSGEMM_LIB"]
    for n in (4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 6, 5, 21, 31], "reps": 10 if n <= 4096 else 4, "tag": "again"}, timeout=600)

if __name__ == "__main__":
    main()
