"""Bring-up probe for the B200 box: numerics of the tcgen05 path (descriptor sweep), ABFT residual floor,
fault injection, and first timings against cuBLAS.  Each case runs in its own subprocess (a faulting kernel kills
only that case).  Writes gpurun_out/probe.jsonl.

  python scripts/gpu_probe.py            # orchestrator
  python scripts/gpu_probe.py case ...   # one case (internal)
"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def run_case(spec, timeout=120):
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, __file__, "case", json.dumps(spec)], capture_output=True, text=True,
                           timeout=timeout)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        res = json.loads(lines[-1]) if lines else {"error": "no output", "stderr": p.stderr[-600:], "rc": p.returncode}
    except subprocess.TimeoutExpired:
        res = {"error": "timeout"}
    res["spec"] = spec
    res["wall_s"] = round(time.time() - t0, 1)
    with open(OUT / "probe.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")
    print(json.dumps(res), flush=True)
    return res


def case_main(spec):
    import numpy as np
    import __graft_entry__ as ge
    import cuda_rt as cu
    from oracle import oracle as O
    pkg = ge.load_package()
    for k, v in spec.get("dbg", {}).items():
        pkg.debug_set(k, v)
    kind = spec["kind"]
    M, N, K = spec["M"], spec["N"], spec["K"]
    rng = np.random.default_rng(spec.get("seed", 0))
    A = (rng.integers(0, 10, M * K) * 0.1 * rng.choice([-1, 1], M * K)).astype(np.float32)
    B = (rng.integers(0, 10, N * K) * 0.1 * rng.choice([-1, 1], N * K)).astype(np.float32)
    if spec.get("dist") == "normal":
        A = rng.standard_normal(M * K).astype(np.float32)
        B = rng.standard_normal(N * K).astype(np.float32)
    alpha, beta = spec.get("alpha", 1.0), spec.get("beta", 0.0)
    C0 = rng.standard_normal(M * N).astype(np.float32) if beta != 0 else np.zeros(M * N, np.float32)
    dA, dB, dC = cu.DevBuf.from_numpy(A), cu.DevBuf.from_numpy(B), cu.DevBuf.from_numpy(C0)
    ft = pkg.FtSgemm()
    res = {}
    if kind == "numerics":
        opts = None
        if spec.get("selftest"):
            opts = pkg.make_opts(selftest=tuple(spec["selftest"]), tau_abs=spec.get("tau_abs", 0), tau_rel=spec.get("tau_rel", 0))
        elif spec.get("faults"):
            opts = pkg.make_opts(faults=spec["faults"], tau_abs=spec.get("tau_abs", 0), tau_rel=spec.get("tau_rel", 0))
        elif spec.get("tau_abs") or spec.get("tau_rel"):
            opts = pkg.make_opts(tau_abs=spec.get("tau_abs", 0), tau_rel=spec.get("tau_rel", 0))
        ft.run(spec["id"], M, N, K, dA, dB, dC, alpha, beta, opts)
        cu.sync()
        got = dC.to_numpy(np.float32, M * N)
        if spec.get("oracle", True) and M * N * K <= 3e9:
            want = O.sgemm_nt(M, N, K, alpha, A, B, beta, C0.copy())
            em = O.error_metrics(want, got)
            res["vs_fp32_oracle"] = {k: float(f"{v:.3e}") for k, v in em.items()}
            res["verify_matrix_first_bad"] = O.verify_matrix(want, got, M, N)
        if M * N * K <= 2e10:
            for mode in ("trunc", "rna"):
                model = O.sgemm_nt_tf32_model(M, N, K, alpha, A, B, beta, C0, mode)
                res["vs_tf32_" + mode] = float(f"{O.error_metrics(model, got)['rel_fro']:.3e}")
        res["finite"] = bool(np.isfinite(got).all())
        res["got_head"] = [float(x) for x in got[:4]]
        info = [k for k in pkg.kernel_table() if k["id"] == spec["id"]][0]
        if info["fault_tolerant"] and info["engine"] == 1:
            res["stats"] = ft.stats()
    elif kind == "timing":
        ids = spec["ids"]
        reps = spec.get("reps", 10)
        tm = cu.Timer()
        res["gflops"] = {}
        for kid in ids:
            opts = None
            if spec.get("reuse") or spec.get("tau_abs"):
                opts = pkg.make_opts(reuse_b_checksums=bool(spec.get("reuse")), tau_abs=spec.get("tau_abs", 0))
            try:
                for _ in range(3):
                    ft.run(kid, M, N, K, dA, dB, dC, alpha, beta, opts)
                cu.sync()
                best = 1e30
                for _ in range(3):
                    tm.start()
                    for _ in range(reps):
                        ft.run(kid, M, N, K, dA, dB, dC, alpha, beta, opts)
                    ms = tm.stop() / reps
                    best = ms if spec.get("sustain") else min(best, ms)  # sustain: the LAST (power-limited) batch
                res["gflops"][str(kid)] = round(2.0 * M * N * K / best / 1e6, 1)
            except Exception as e:  # noqa
                res["gflops"][str(kid)] = "ERR " + str(e)[:80]
    print(json.dumps(res))


def main():
    if (OUT / "probe.jsonl").exists():
        (OUT / "probe.jsonl").unlink()
    # ---- 1. numerics of the plain kernel, default descriptors, single tile then multi-tile
    ok_default = True
    for kid, bn in ((6, 128), (2, 64), (1, 32), (5, 256)):
        r = run_case({"kind": "numerics", "id": kid, "M": 128, "N": bn, "K": 64})
        e = r.get("vs_tf32_trunc", 1)
        if not (min(e, r.get("vs_tf32_rna", 1)) < 1e-4):
            ok_default = False
    if not ok_default:
        # ---- descriptor sweep on one tile (128x128x64)
        sweeps = []
        for layout, sw in ((1, 4), (2, 3), (1, 3), (2, 4), (1, 5), (1, 6)):
            for sbo in (512, 1024, 256):
                for lbo in (4096, 1024, 512):
                    for kstep in (1024, 512):
                        sweeps.append({"layout_type": layout, "tma_swizzle": sw, "sbo": sbo, "lbo": lbo, "kstep": kstep})
        best = None
        for d in sweeps:
            r = run_case({"kind": "numerics", "id": 6, "M": 128, "N": 128, "K": 64, "dbg": d, "oracle": False}, timeout=60)
            e = min(r.get("vs_tf32_trunc", 1), r.get("vs_tf32_rna", 1))
            if best is None or e < best[0]:
                best = (e, d)
            if e < 1e-4:
                break
        print("BEST", best, flush=True)
        if best[0] >= 1e-4:
            return
        dflt = best[1]
    else:
        dflt = {}
    # ---- 2. bigger numerics (multi-tile, multi-k, ragged, alpha/beta)
    for spec in (
        {"kind": "numerics", "id": 6, "M": 512, "N": 512, "K": 512},
        {"kind": "numerics", "id": 5, "M": 1024, "N": 1024, "K": 1024},
        {"kind": "numerics", "id": 6, "M": 1024, "N": 1024, "K": 1024, "alpha": 0.75, "beta": -1.5},
        {"kind": "numerics", "id": 6, "M": 200, "N": 136, "K": 100},
        {"kind": "numerics", "id": 2, "M": 1024, "N": 1024, "K": 1024, "dist": "normal"},
        {"kind": "numerics", "id": 1, "M": 4096, "N": 4096, "K": 512, "oracle": False},
    ):
        spec["dbg"] = dflt
        run_case(spec, timeout=300)
    # ---- 3. ABFT: residual floor for the three encode roundings, then injection
    for enc in (0, 1, 2):
        for kid, K in ((16, 1024), (15, 4096), (16, 8192)):
            run_case({"kind": "numerics", "id": kid, "M": 1024, "N": 1024, "K": K, "oracle": False,
                      "tau_abs": 1e9, "dbg": dict(dflt, enc_rounding=enc)}, timeout=300)
    run_case({"kind": "numerics", "id": 16, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 17, 5], "dbg": dflt})
    run_case({"kind": "numerics", "id": 15, "M": 1024, "N": 1024, "K": 1024, "selftest": [10000.0, 100, 200], "dbg": dflt})
    run_case({"kind": "numerics", "id": 16, "M": 1024, "N": 1024, "K": 1024, "dbg": dflt,
              "faults": [{"row": 5, "col": 7, "add": 1.0}, {"row": 300, "col": 900, "xor": 1 << 30},
                         {"row": 777, "col": 333, "xor": 1 << 22}, {"row": 1000, "col": 64, "xor": 1 << 31}]})
    # ---- 4. timings
    for n in (1024, 2048, 4096, 8192):
        run_case({"kind": "timing", "M": n, "N": n, "K": n, "ids": [0, 7, 1, 2, 6, 5, 11, 12, 16, 15, 10, 30],
                  "reps": 10 if n <= 4096 else 4, "dbg": dflt}, timeout=600)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "case":
        case_main(json.loads(sys.argv[2]))
    else:
        main()
