"""Probe 8: where does the larger fault-free residual at K=8192 come from (split-K slices? encode arithmetic?)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    base = {"kind": "numerics", "id": 31, "M": 1024, "N": 1024, "K": 8192, "tau_abs": 1e9, "oracle": False}
    run_case(dict(base, tag="auto"))
    run_case(dict(base, dbg={"splitk": 0}, tag="nosplit"))
    for s in (2, 3, 4, 8):
        run_case(dict(base, dbg={"splitk": s}, tag="S=%d" % s))
    run_case(dict(base, id=16, dbg={"splitk": 0}, tag="huge-nosplit"))
    run_case(dict(base, id=16, dbg={"splitk": 4}, tag="huge-S4"))
    run_case(dict(base, dist="normal", dbg={"splitk": 0}, tag="normal-nosplit"))
    run_case(dict(base, dist="normal", dbg={"splitk": 4}, tag="normal-S4"))
    run_case(dict(base, M=4096, N=4096, K=4096, tag="4096-auto"))
    run_case(dict(base, M=8192, N=8192, K=8192, tag="8192-auto"), timeout=300)

if __name__ == "__main__":
    main()
