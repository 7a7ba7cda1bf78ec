"""Probe 24: helper-assisted epilogue on/off."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 3072, 4096, 8192]:
        reps = 20 if n <= 4096 else 5
        for rnd in range(2):
            for on in (1, 0):
                run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31, 16], "reps": reps, "beta": -1.5, "dbg": {"epi_assist": on}, "tag": f"assist={on}"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7], "reps": reps, "beta": -1.5, "tag": "cublas"}, timeout=600)

if __name__ == "__main__":
    main()
