"""Probe 19: encode modes in the SUSTAINED (power-limited) regime: 3 x N back-to-back launches, last batch."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        reps = 400 if n <= 4096 else 60
        base = {"kind": "timing", "M": n, "N": n, "K": n, "reps": reps, "sustain": 1, "beta": -1.5}
        run_case({**base, "ids": [7, 21], "tag": "ref"}, timeout=600)
        for rnd in range(2):
            for mode in (1, 3, 2):
                run_case({**base, "ids": [31], "dbg": {"enc_mode": mode}, "tag": f"enc_mode={mode}"}, timeout=600)
        run_case({**base, "ids": [7], "tag": "ref-again"}, timeout=600)

if __name__ == "__main__":
    main()
