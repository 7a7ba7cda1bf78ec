"""Probe 10: pipeline depth A/B (run with FTSGEMM_LIB=scripts/libftsgemm_s6.so etc.)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case

def main():
    tag = os.path.basename(os.environ.get("FTSGEMM_LIB", "default"))
    for n in (4096, 8192):
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31, 7], "reps": reps, "dbg": {"enc_mode": 1}, "tag": tag}, timeout=600)

if __name__ == "__main__":
    main()
