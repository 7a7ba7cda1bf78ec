"""Probe 22: encode pre-pass item size / grid."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gpu_probe import run_case, OUT

def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
        reps = 20 if n <= 4096 else 6
        for rnd in range(2):
            for kr, bps in ((8, 4), (4, 2), (4, 4), (8, 2)):
                run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "dbg": {"enc_kr": kr, "enc_blocks_per_sm": bps}, "tag": f"kr={kr},blocks/sm={bps}"}, timeout=600)
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [31], "reps": reps, "reuse": 1, "tag": "reuse"}, timeout=600)

if __name__ == "__main__":
    main()
