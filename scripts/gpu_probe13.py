"""Probe 13: A/B of a library variant (FTSGEMM_LIB) -- timing of cuBLAS / plain / ABFT in the encode modes."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case

def main():
    tag = os.path.basename(os.environ.get("FTSGEMM_LIB", "default"))
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "tag": tag + ":tiles"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": tag + ":prepass"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 2}, "tag": tag + ":items"}, timeout=600)

if __name__ == "__main__":
    main()
