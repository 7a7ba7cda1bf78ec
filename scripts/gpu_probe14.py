"""Probe 14: encoder-tile cost model sweep vs pre-pass."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": "prepass"}, timeout=600)
        for c in (1030, 1500, 1850, 2200):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 3, "enc_tile_cost_permille": c}, "tag": f"tiles-{c}"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": "prepass-again"}, timeout=600)

if __name__ == "__main__":
    main()
