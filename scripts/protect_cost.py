"""Cost of opts.protect_epilogue in the burst regime: us per launch with / without, interleaved, id 31 (and 16)."""
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
ft = pkg.FtSgemm()
stream = torch.cuda.current_stream().cuda_stream
big = 8192
g = torch.Generator(device="cuda").manual_seed(7)
dA = torch.randint(-9, 10, (big * big,), generator=g, device="cuda").float() * 0.1
dB = torch.randint(-9, 10, (big * big,), generator=g, device="cuda").float() * 0.1
dC = torch.zeros(big * big, device="cuda")
for kid, n in ((31, 2048), (31, 4096), (31, 8192), (16, 4096)):
    res = {0: [], 1: []}
    for _ in range(3):
        for prot in (0, 1):
            o = pkg.make_opts(stream=stream, protect_epilogue=bool(prot))
            dC.zero_()  # (beta = -1.5 on the same C: it grows 1.5x per launch)
            for _ in range(3):
                ft.run(kid, n, n, n, dA, dB, dC, 1.0, -1.5, o)
            torch.cuda.synchronize()
            time.sleep(0.05)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ft.run(kid, n, n, n, dA, dB, dC, 1.0, -1.5, o)
            e1.record()
            torch.cuda.synchronize()
            res[prot].append(e0.elapsed_time(e1) / 20 * 1e3)
    a, b = statistics.median(res[0]), statistics.median(res[1])
    print(json.dumps({"id": kid, "n": n, "us": round(a, 2), "us_protected": round(b, 2), "cost_pct": round(100 * (b / a - 1), 2)}), flush=True)
st = ft.stats()
print(json.dumps({k: st[k] for k in ("rows_checked", "detected", "epilogue_faults", "uncorrectable")}))
