"""Probe 11: planner variants (default / uncut / lockstep off) over sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 6144, 8192, 12288]
    for n in sizes:
        reps = 10 if n <= 4096 else (4 if n <= 8192 else 2)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "tag": "default"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"splitk": 0}, "tag": "uncut"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"lockstep": 0}, "tag": "free-cut"}, timeout=600)

if __name__ == "__main__":
    main()
