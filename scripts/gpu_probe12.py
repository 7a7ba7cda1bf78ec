"""Probe 12: encoder items vs pre-pass, encoder cost model sweep."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
    for n in sizes:
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": "prepass"}, timeout=600)
        for c in (450, 800, 1200, 1600):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 2, "enc_cost_permille": c}, "tag": f"items-{c}"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": "prepass-again"}, timeout=600)

if __name__ == "__main__":
    main()
