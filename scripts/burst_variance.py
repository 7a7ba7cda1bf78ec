"""Burst-to-burst variance of one size: prints every burst (20 launches) of ids 31 / 21 / 7, interleaved.
usage: burst_variance.py [n=8192] [bursts=10]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n, bursts = int(kv.get("n", 8192)), int(kv.get("bursts", 10))
ft = pkg.FtSgemm()
g = torch.Generator(device="cuda").manual_seed(7)
dA = torch.randint(-9, 10, (n * n,), generator=g, device="cuda").float() * 0.1
dB = torch.randint(-9, 10, (n * n,), generator=g, device="cuda").float() * 0.1
dC = torch.zeros(n * n, device="cuda")
o = pkg.make_opts(stream=torch.cuda.current_stream().cuda_stream)
out = {31: [], 21: [], 7: []}
for b in range(bursts):
    for kid in (7, 31, 21):
        dC.zero_()
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
        out[kid].append(round(e0.elapsed_time(e1) / 20 * 1e3, 1))
print(json.dumps({"n": n, "us_per_launch": {str(k): v for k, v in out.items()}}))
