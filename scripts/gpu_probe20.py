"""Probe 20: library variants (FTSGEMM_LIB) in the sustained regime."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case

def main():
    tag = os.path.basename(os.environ.get("FTSGEMM_LIB", "default"))
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        reps = 400 if n <= 4096 else 60
        base = {"kind": "timing", "M": n, "N": n, "K": n, "reps": reps, "sustain": 1, "beta": -1.5}
        run_case({**base, "ids": [7, 21, 31], "tag": tag}, timeout=600)

if __name__ == "__main__":
    main()
