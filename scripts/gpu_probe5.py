"""Probe 5: dissect the ABFT overhead (encode / checksum tiles / epilogue check), narrowed checksum tiles, new encode."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    run_case({"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200]})
    run_case({"kind": "numerics", "id": 16, "M": 1056, "N": 1120, "K": 520, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 15, "M": 512, "N": 2304, "K": 256, "selftest": [10000.0, 17, 5]})
    run_case({"kind": "numerics", "id": 31, "M": 4096, "N": 4096, "K": 4096, "oracle": False, "selftest": [10000.0, 17, 5]})
    for n in (4096, 8192):
        reps = 10 if n <= 4096 else 4
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [7, 21, 31, 5, 15], "reps": reps, "tag": "full"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31, 15], "reps": reps, "reuse": 1, "tag": "reuse-encode"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31, 15], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 1}, "tag": "no-check"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31, 15], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 2}, "tau_abs": 1e30, "tag": "no-chk-tiles"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31, 15], "reps": reps, "reuse": 1, "dbg": {"ft_dbg": 3}, "tag": "neither"}, timeout=600)

if __name__ == "__main__":
    main()
