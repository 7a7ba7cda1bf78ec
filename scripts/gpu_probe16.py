"""Probe 16: where the ABFT time goes in the SUSTAINED (power-limited) regime: 3 x 400 back-to-back launches, last batch."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    reps = 400 if n <= 4096 else 60
    base = {"kind": "timing", "M": n, "N": n, "K": n, "reps": reps, "sustain": 1, "beta": -1.5}
    run_case({**base, "ids": [7, 21, 31], "tag": "sustained"}, timeout=600)
    run_case({**base, "ids": [31], "reuse": 1, "tag": "reuse"}, timeout=600)
    run_case({**base, "ids": [31], "reuse": 1, "dbg": {"ft_dbg": 1}, "tag": "reuse,no-check"}, timeout=600)
    run_case({**base, "ids": [31], "reuse": 1, "tau_abs": 1e30, "dbg": {"ft_dbg": 2}, "tag": "reuse,no-chk-tiles"}, timeout=600)
    run_case({**base, "ids": [31], "reuse": 1, "tau_abs": 1e30, "dbg": {"ft_dbg": 3}, "tag": "reuse,neither"}, timeout=600)
    run_case({**base, "ids": [21, 31], "dbg": {"splitk": 0}, "tag": "uncut"}, timeout=600)
    run_case({**base, "ids": [7, 21], "tag": "sustained-again"}, timeout=600)

if __name__ == "__main__":
    main()
