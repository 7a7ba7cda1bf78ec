"""Probe 23: reduced vs full planner search."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [8192, 10240, 7168]:
        reps = 6 if n <= 8192 else 3
        for rnd in range(2):
            for full in (0, 1):
                run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"plan_full_search": full}, "tag": f"full_search={full}"}, timeout=600)

if __name__ == "__main__":
    main()
