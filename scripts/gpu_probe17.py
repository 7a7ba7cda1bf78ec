"""Probe 17: large sizes -- raster group width of the persistent kernel vs cuBLAS."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [16384]:
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21], "reps": 2, "tag": "default(group 8)"}, timeout=600)
        for g in (2, 4, 16, 32):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21], "reps": 2, "dbg": {"group_n": g}, "tag": f"group_n={g}"}, timeout=600)

if __name__ == "__main__":
    main()
