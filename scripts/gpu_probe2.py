"""Second bring-up probe: CTA-pair (cta_group::2) kernels, checksum tile-column ABFT, timings."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for kid, n in ((21, 256), (22, 256), (21, 1024), (22, 1024)):
        run_case({"kind": "numerics", "id": kid, "M": n, "N": n, "K": 256})
    run_case({"kind": "numerics", "id": 21, "M": 1024, "N": 768, "K": 1000, "alpha": 0.75, "beta": -1.5})
    run_case({"kind": "numerics", "id": 21, "M": 200, "N": 136, "K": 100})
    for kid in (11, 12, 16, 15, 31, 32):
        run_case({"kind": "numerics", "id": kid, "M": 1024, "N": 1024, "K": 1024, "tau_abs": 1e9})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 8192, "tau_abs": 1e9, "oracle": False})
    run_case({"kind": "numerics", "id": 31, "M": 4096, "N": 4096, "K": 4096, "oracle": False})
    run_case({"kind": "numerics", "id": 16, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024,
              "faults": [{"row": 5, "col": 7, "add": 1.0}, {"row": 300, "col": 900, "xor": 1 << 30},
                         {"row": 777, "col": 333, "xor": 1 << 22}, {"row": 1000, "col": 64, "xor": 1 << 31}]})
    for n in (1024, 2048, 4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 2, 6, 5, 21, 22, 12, 16, 15, 31, 32],
                  "reps": 10 if n <= 4096 else 4}, timeout=600)
    run_case({"kind": "timing", "M": 16384, "N": 16384, "K": 16384, "ids": [7, 21, 31], "reps": 2}, timeout=600)

if __name__ == "__main__":
    main()
