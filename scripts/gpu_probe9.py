"""Probe 9: seeded cut tiles + in-kernel encode vs the pre-pass / uncut variants."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 8192, "tau_abs": 1e9, "oracle": False})
    run_case({"kind": "numerics", "id": 31, "M": 4096, "N": 4096, "K": 4096, "tau_abs": 1e9, "oracle": False})
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192]
    for n in sizes:
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31], "reps": reps, "tag": "default"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_mode": 1}, "tag": "prepass"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [21, 31], "reps": reps, "dbg": {"splitk": 0}, "tag": "uncut"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "tag": "reuse-encode"}, timeout=600)

if __name__ == "__main__":
    main()
