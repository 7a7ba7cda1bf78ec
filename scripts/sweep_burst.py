"""Burst-regime sweep (the driver's regime: 20 back-to-back launches per cell at boost clocks, idle gaps in between).

usage: sweep_burst.py [steps=20] [rounds=3] [ids=31,21,7] [sizes=1024,2048,...] [key=value debug knobs]
Per size and round: every engine gets `steps` launches between two CUDA events (after 3 warm-up launches); engines are
interleaved and the GPU idles 50 ms between cells, so that all cells are measured in the same clock state.  Prints one
JSON line per size: median TFLOP/s per engine and the overhead of each engine vs id 7 (cuBLAS-TF32).
"""
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def main():
    pkg = ge.load_package()
    kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    steps = int(kv.pop("steps", 20))
    rounds = int(kv.pop("rounds", 3))
    ids = [int(x) for x in kv.pop("ids", "31,21,7").split(",")]
    sizes = [int(x) for x in kv.pop("sizes", ",".join(str(s) for s in range(1024, 16385, 1024))).split(",")]
    beta = float(kv.pop("beta", -1.5))
    pause = float(kv.pop("pause", 0.05))
    reuse = int(kv.pop("reuse", 0))
    for k, v in kv.items():
        pkg.debug_set(k, int(v))
    big = max(sizes)
    g = torch.Generator(device="cuda").manual_seed(7)
    def ref_dist(count):
        return (torch.randint(0, 10, (count,), generator=g, device="cuda").float() * 0.1) * \
               (torch.randint(0, 2, (count,), generator=g, device="cuda").float() * 2 - 1)
    dA, dB = ref_dist(big * big), ref_dist(big * big)
    dC = torch.zeros(big * big, device="cuda")
    ft = pkg.FtSgemm()
    stream = torch.cuda.current_stream().cuda_stream
    opts = pkg.make_opts(stream=stream, reuse_b_checksums=bool(reuse))
    out = []
    for n in sizes:
        res = {i: [] for i in ids}
        for r in range(rounds):
            for kid in ids:
                dC[: n * n].zero_()
                for _ in range(3):
                    ft.run(kid, n, n, n, dA, dB, dC, 1.0, beta, opts)
                torch.cuda.synchronize()
                time.sleep(pause)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    ft.run(kid, n, n, n, dA, dB, dC, 1.0, beta, opts)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[kid].append(2.0 * n ** 3 / ms / 1e9)
        med = {i: statistics.median(v) for i, v in res.items()}
        line = {"n": n, "steps": steps, "tflops": {str(i): round(med[i], 1) for i in ids},
                "us": {str(i): round(2.0 * n ** 3 / med[i] / 1e6, 2) for i in ids}}
        if 7 in med:
            line["overhead_pct_vs_7"] = {str(i): round(100.0 * (med[7] / med[i] - 1.0), 2) for i in ids if i != 7}
        print(json.dumps(line), flush=True)
        out.append(line)
    st = ft.stats()
    print(json.dumps({"stats": {k: st[k] for k in ("tiles", "rows_checked", "detected", "max_rel_residual")}}))


if __name__ == "__main__":
    main()
