"""Probe 21: wave re-synchronisation on/off at large sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [16384, 12288, 8192]:
        reps = 4 if n <= 8192 else 2
        for rnd in range(2):
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"wave_sync": 0}, "tag": "sync-off"}, timeout=600)
            run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"wave_sync": 1}, "tag": "sync-on"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7], "reps": reps, "tag": "cublas"}, timeout=600)

if __name__ == "__main__":
    main()
