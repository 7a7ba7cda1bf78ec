"""Probe 15: programmatic dependent launch of the GEMM behind the encode pre-pass, on/off."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]:
        reps = 10 if n <= 4096 else 4
        for rep in range(2):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"pdl": 1}, "tag": "pdl-on"}, timeout=600)
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"pdl": 0}, "tag": "pdl-off"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21], "reps": reps, "tag": "ref"}, timeout=600)

if __name__ == "__main__":
    main()
