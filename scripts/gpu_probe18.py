"""Probe 18: raster group width 8 vs 16 over sizes, plain and ABFT."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [6144, 8192, 10240, 12288, 16384]:
        reps = 4 if n <= 8192 else 2
        for g in (8, 16, 12):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"group_n": g}, "tag": f"group_n={g}"}, timeout=600)

if __name__ == "__main__":
    main()
