"""Host-side mirror of the reference's ft_sgemm interface over the C ABI of libftsgemm.so (include/ftsgemm.h).

The reference (shixun404/Fault-Tolerant-SGEMM-on-NVIDIA-GPUs) has no Python API; its operator surface is
  * the kernel contract  k(M, N, K, A, B, C, alpha, beta)  selected by a kernel id
    (kernel/ft_sgemm/sgemm.cu:110-199, ids/names at :235-237, tiles at code_gen/main.py:8-16),
  * baseline_ft_sgemm(...)  (kernel/ft_sgemm/include/baseline_ft_sgemm.cuh:1),
  * verify_matrix(...)      (utils/utils.cu:61-77),
  * the CLI  ft_sgemm START END GAP ST_KERNEL END_KERNEL  (sgemm.cu:13-19).
This module mirrors exactly those, binding the shared library with ctypes (plain pointers and sizes; torch is only
used by callers to own device memory).  There is NO CPU fallback: if the library or an sm_100 device is missing,
every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libftsgemm.so"
CLI_PATH = HERE / "ft_sgemm"

MAX_FAULTS = 8
MAX_EVENTS = 16

# kernel ids (include/ftsgemm.h; reference sgemm.cu:235-237)
ID_CUBLAS, ID_CUBLAS_TF32, ID_ABFT_BASELINE, ID_ABFT_BASELINE_TF32 = 0, 7, 10, 30
ID_SGEMM_AUTO, ID_ABFT_AUTO = 20, 40  # per-shape choice among the tcgen05 variants (select_kernel)
SGEMM_IDS = {"small": 1, "medium": 2, "large": 3, "tall": 4, "wide": 5, "huge": 6, "giant": 21, "pair128": 22}
ABFT_IDS = {"small": 11, "medium": 12, "large": 13, "tall": 14, "wide": 15, "huge": 16, "giant": 31, "pair128": 32}


class FtsgemmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ftsgemm error {code}: {msg}")
        self.code = code


class KernelInfo(C.Structure):
    _fields_ = [("id", C.c_int), ("name", C.c_char * 24), ("fault_tolerant", C.c_int), ("engine", C.c_int),
                ("ref_tile_m", C.c_int), ("ref_tile_n", C.c_int), ("ref_tile_k", C.c_int),
                ("tile_m", C.c_int), ("tile_n", C.c_int), ("tile_k", C.c_int)]


class Fault(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("mode", C.c_int), ("add_value", C.c_float),
                ("xor_mask", C.c_uint32)]


class Opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("stream", C.c_void_p), ("inject_mode", C.c_int),
                ("selftest_value", C.c_float), ("selftest_row", C.c_int), ("selftest_col", C.c_int),
                ("n_faults", C.c_int), ("faults", Fault * MAX_FAULTS), ("tau_abs", C.c_float),
                ("tau_rel", C.c_float), ("detect_only", C.c_int), ("reuse_b_checksums", C.c_int),
                ("baseline_host_sync", C.c_int), ("precision", C.c_int), ("check_segments", C.c_int),
                ("no_recompute", C.c_int), ("protect_epilogue", C.c_int)]


class Event(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("residual", C.c_float), ("corrected_value", C.c_float),
                ("status", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("tiles", C.c_ulonglong), ("rows_checked", C.c_ulonglong), ("detected", C.c_ulonglong),
                ("corrected", C.c_ulonglong), ("uncorrectable", C.c_ulonglong), ("checksum_faults", C.c_ulonglong),
                ("max_abs_residual", C.c_float), ("max_rel_residual", C.c_float), ("n_events", C.c_int),
                ("events", Event * MAX_EVENTS), ("recomputed", C.c_ulonglong),
                ("epilogue_faults", C.c_ulonglong)]

    def as_dict(self):
        return {"tiles": self.tiles, "rows_checked": self.rows_checked, "detected": self.detected,
                "corrected": self.corrected, "uncorrectable": self.uncorrectable, "recomputed": self.recomputed, "epilogue_faults": self.epilogue_faults,
                "checksum_faults": self.checksum_faults, "max_abs_residual": self.max_abs_residual,
                "max_rel_residual": self.max_rel_residual,
                "events": [{"row": e.row, "col": e.col, "residual": e.residual,
                            "corrected_value": e.corrected_value, "status": e.status}
                           for e in list(self.events)[:self.n_events]]}


# every symbol include/ftsgemm.h declares (checked by tests/test_abi.py without a GPU)
EXPORTED_SYMBOLS = [
    "ftsgemm_create", "ftsgemm_destroy", "ftsgemm_abi_version", "ftsgemm_error_string", "ftsgemm_last_cuda_error",
    "ftsgemm_default_opts", "ftsgemm_kernel_table", "ftsgemm_kernel_lookup", "ftsgemm_run", "ftsgemm_get_stats",
    "ftsgemm_run_host", "ftsgemm_baseline", "ftsgemm_verify", "ftsgemm_debug_set", "ftsgemm_debug_schedule",
    "ftsgemm_verify_bad_count", "ftsgemm_debug_trace", "ftsgemm_stats_device", "ftsgemm_select_kernel", "ftsgemm_launch_count", "ftsgemm_peer_export", "ftsgemm_peer_connect",
    "ftsgemm_peer_verdict",
]

_lib = None


def build(force: bool = False, verbose: bool = False) -> None:
    from importlib import util
    spec = util.spec_from_file_location("_ftsgemm_build", HERE / "build.py")
    mod = util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build(force=force, verbose=verbose)


def lib():
    """Load libftsgemm.so (raises if it has not been built: the product path never falls back to the CPU)."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("FTSGEMM_LIB", str(LIB_PATH)))  # override only for A/B experiments
        if not path.exists():
            raise FtsgemmError(-4, f"{path} not built (run __graft_entry__.build())")
        L = C.CDLL(str(path))
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.ftsgemm_create.argtypes = [C.POINTER(vp)]
        L.ftsgemm_destroy.argtypes = [vp]
        L.ftsgemm_error_string.argtypes = [ip]
        L.ftsgemm_error_string.restype = C.c_char_p
        L.ftsgemm_last_cuda_error.argtypes = [vp]
        L.ftsgemm_default_opts.argtypes = [C.POINTER(Opts)]
        L.ftsgemm_default_opts.restype = None
        L.ftsgemm_kernel_table.argtypes = [C.POINTER(KernelInfo), ip]
        L.ftsgemm_kernel_lookup.argtypes = [ip, C.POINTER(KernelInfo)]
        L.ftsgemm_run.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, fp, fp, C.POINTER(Opts)]
        L.ftsgemm_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.ftsgemm_run_host.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp, fp, fp, C.POINTER(Opts)]
        L.ftsgemm_stats_device.argtypes = [vp, vp, vp]
        L.ftsgemm_select_kernel.argtypes = [ip, ip, ip, ip]
        L.ftsgemm_peer_export.argtypes = [vp, vp]
        L.ftsgemm_peer_connect.argtypes = [vp, ip, ip, vp]
        L.ftsgemm_peer_verdict.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), ip]
        L.ftsgemm_launch_count.argtypes = [vp]
        L.ftsgemm_launch_count.restype = C.c_ulonglong
        L.ftsgemm_baseline.argtypes = [vp, ip, ip, ip, vp, vp, vp, fp, fp, ip, C.POINTER(Opts), vp]
        L.ftsgemm_verify.argtypes = [vp, vp, vp, ip, ip, C.POINTER(C.c_longlong), C.POINTER(C.c_double), vp]
        L.ftsgemm_debug_set.argtypes = [C.c_char_p, C.c_longlong]
        L.ftsgemm_verify_bad_count.argtypes = [vp]
        L.ftsgemm_verify_bad_count.restype = C.c_longlong
        L.ftsgemm_debug_schedule.argtypes = [ip, ip, ip, ip, ip, C.POINTER(C.c_int), C.POINTER(C.c_int), ip]
        L.ftsgemm_debug_trace.argtypes = [vp, C.POINTER(C.c_ulonglong), ip]
        _lib = L
    return _lib


def _check(code: int):
    if code != 0:
        raise FtsgemmError(code, lib().ftsgemm_error_string(code).decode())


def kernel_table():
    """The kernel-variant table (reference: sgemm.cu:235-237 + code_gen/main.py:8-16)."""
    n = lib().ftsgemm_kernel_table(None, 0)
    arr = (KernelInfo * n)()
    lib().ftsgemm_kernel_table(arr, n)
    return [{"id": k.id, "name": k.name.decode(), "fault_tolerant": bool(k.fault_tolerant), "engine": k.engine,
             "ref_tile": (k.ref_tile_m, k.ref_tile_n, k.ref_tile_k), "tile": (k.tile_m, k.tile_n, k.tile_k)}
            for k in arr]


def select_kernel(M: int, N: int, K: int, fault_tolerant: bool) -> int:
    """The concrete kernel id the AUTO ids (20 / 40) resolve to for this shape."""
    r = lib().ftsgemm_select_kernel(M, N, K, 1 if fault_tolerant else 0)
    if r < 0:
        _check(r)
    return r


def default_opts() -> Opts:
    o = Opts()
    lib().ftsgemm_default_opts(C.byref(o))
    return o


def make_opts(stream=None, selftest=None, faults=None, tau_abs=0.0, tau_rel=0.0, detect_only=False,
              reuse_b_checksums=False, baseline_host_sync=True, no_recompute=False, precision=0, check_segments=0,
              protect_epilogue=False) -> Opts:
    """selftest: None | (value, tile_row, tile_col)  -> the reference's always-on injector (ft_sgemm_huge.cuh:324-327)
    faults: list of dicts {row, col, add=float} or {row, col, xor=int}; with xor, where="epilogue_tmem" | "epilogue_value"
    places the upset after the accumulator check (ftsgemm_fault.mode 2 / 3)"""
    o = default_opts()
    o.stream = stream
    if selftest is not None:
        o.inject_mode = 1
        o.selftest_value, o.selftest_row, o.selftest_col = float(selftest[0]), int(selftest[1]), int(selftest[2])
    if faults:
        assert selftest is None and len(faults) <= MAX_FAULTS
        o.inject_mode = 2
        o.n_faults = len(faults)
        for i, f in enumerate(faults):
            o.faults[i].row, o.faults[i].col = int(f["row"]), int(f["col"])
            if "xor" in f:
                o.faults[i].mode, o.faults[i].xor_mask = {"acc": 1, "epilogue_tmem": 2, "epilogue_value": 3}[f.get("where", "acc")], int(f["xor"]) & 0xFFFFFFFF
            else:
                o.faults[i].mode, o.faults[i].add_value = 0, float(f["add"])
    o.tau_abs, o.tau_rel = float(tau_abs), float(tau_rel)
    o.detect_only = int(detect_only)
    o.reuse_b_checksums = int(reuse_b_checksums)
    o.baseline_host_sync = int(baseline_host_sync)
    o.no_recompute = int(no_recompute)
    o.precision = int(precision)  # 0 = single-pass TF32, 1 = 3xTF32 (FP32-grade)
    o.check_segments = int(check_segments)  # > 1: intra-K checking (K-segments verified one after the other)
    o.protect_epilogue = int(protect_epilogue)
    return o


def debug_schedule(kernel_id: int, M: int, N: int, K: int, num_sms: int = 148):
    """Work decomposition of one launch, enumerated on the host (no GPU needed): (header dict, list of segment dicts)."""
    hdr = (C.c_int * 8)()
    n = lib().ftsgemm_debug_schedule(kernel_id, M, N, K, num_sms, hdr, None, 0)
    if n < 0:
        _check(n)
    rows = (C.c_int * (9 * max(n, 1)))()
    lib().ftsgemm_debug_schedule(kernel_id, M, N, K, num_sms, hdr, rows, n)
    keys = ("unit", "tile", "is_chk", "m_blk", "n_blk", "kb_begin", "kb_end", "kind", "slice")
    segs = [dict(zip(keys, rows[9 * i:9 * i + 9])) for i in range(n)]
    return dict(zip(("units", "num_tiles", "n_chk_tiles", "sk_tiles", "num_kb", "cta_group", "sk_slices", "chk_slices"), hdr)), segs


def debug_set(key: str, value: int) -> None:
    _check(lib().ftsgemm_debug_set(key.encode(), int(value)))


def _ptr(x) -> int:
    """Device/host address of a torch tensor, numpy array or raw int."""
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class FtSgemm:
    """One handle = one device context (cuBLAS handle, checksum workspace, fault counters)."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(lib().ftsgemm_create(C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ftsgemm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run_checked(self, code):
        if code != 0:
            msg = lib().ftsgemm_error_string(code).decode()
            raise FtsgemmError(code, f"{msg} (cuda/cublas status {lib().ftsgemm_last_cuda_error(self._h)})")

    # the reference kernel contract:  kernel<<<...>>>(M, N, K, dA, dB, dC, alpha, beta)  selected by id
    def run(self, kernel_id: int, M: int, N: int, K: int, dA, dB, dC, alpha: float = 1.0, beta: float = 0.0,
            opts: Opts | None = None) -> None:
        self._run_checked(lib().ftsgemm_run(self._h, kernel_id, M, N, K, _ptr(dA), _ptr(dB), _ptr(dC), alpha, beta,
                                            C.byref(opts) if opts is not None else None))

    def run_host(self, kernel_id: int, M: int, N: int, K: int, hA, hB, hC, alpha: float = 1.0, beta: float = 0.0,
                 opts: Opts | None = None) -> None:
        self._run_checked(lib().ftsgemm_run_host(self._h, kernel_id, M, N, K, _ptr(hA), _ptr(hB), _ptr(hC), alpha,
                                                 beta, C.byref(opts) if opts is not None else None))

    # baseline_ft_sgemm (include/baseline_ft_sgemm.cuh:1)
    def baseline(self, M, N, K, dA, dB, dC, alpha=1.0, beta=0.0, tf32=False, opts: Opts | None = None,
                 residual_out=None) -> None:
        self._run_checked(lib().ftsgemm_baseline(self._h, M, N, K, _ptr(dA), _ptr(dB), _ptr(dC), alpha, beta,
                                                 1 if tf32 else 0, C.byref(opts) if opts is not None else None,
                                                 _ptr(residual_out) if residual_out is not None else None))

    def stats(self) -> dict:
        s = Stats()
        self._run_checked(lib().ftsgemm_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # fused multi-GPU verdict exchange (include/ftsgemm.h: ftsgemm_peer_*)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._run_checked(lib().ftsgemm_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, rank: int, world: int, handles) -> None:
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._run_checked(lib().ftsgemm_peer_connect(self._h, rank, world, C.create_string_buffer(blob, len(blob))))

    def peer_verdict(self, world: int, timeout_ms: int = 10000):
        out, per = (C.c_double * 8)(), (C.c_double * (8 * world))()
        self._run_checked(lib().ftsgemm_peer_verdict(self._h, out, per, timeout_ms))
        return list(out), [list(per[8 * r:8 * r + 8]) for r in range(world)]

    def launch_count(self) -> int:
        """Kernels of this library launched through the handle so far (include/ftsgemm.h)."""
        return int(lib().ftsgemm_launch_count(self._h))

    def stats_device(self, d_out8, stream=None) -> None:
        """Device-side verdict vector (8 doubles, see include/ftsgemm.h), asynchronous on `stream`, counters not reset."""
        self._run_checked(lib().ftsgemm_stats_device(self._h, _ptr(d_out8), stream))

    def debug_trace(self):
        """Timeline of the last launch made under debug_set("trace", 1): list (per unit) of item dicts, times in ns."""
        cap = 160 * 512
        buf = (C.c_ulonglong * cap)()
        n = lib().ftsgemm_debug_trace(self._h, buf, cap)
        if n < 0:
            self._run_checked(n)
        keys = ("prod_start", "prod_end", "mma_start", "mma_end", "acc_done", "check_done", "epi_end")
        out = []
        for u in range(n):
            items = []
            for i in range(63):
                r = buf[(u * 64 + i) * 8:(u * 64 + i) * 8 + 8]
                if r[4] == 0:  # unused slot, or an encoder item (it has no accumulator / epilogue)
                    continue
                d = dict(zip(keys, r[:7]))
                d["tile"] = r[7] & 0xFFFFFF
                d["kind"] = r[7] >> 24
                items.append(d)
            if items:
                items[0]["enc_end"] = buf[(u * 64 + 63) * 8]  # helper warps finished their share of the in-kernel encode
                items[0]["enc_start"] = buf[(u * 64 + 63) * 8 + 1]
                items[0]["enc_worker0"] = (buf[(u * 64 + 63) * 8 + 2], buf[(u * 64 + 63) * 8 + 3], buf[(u * 64 + 63) * 8 + 5])
            out.append(items)
        return out

    # verify_matrix (utils/utils.cu:61-77) on device buffers
    def verify(self, d_ref, d_x, M: int, N: int, stream=None):
        bad, rel = C.c_longlong(), C.c_double()
        code = lib().ftsgemm_verify(self._h, _ptr(d_ref), _ptr(d_x), M, N, C.byref(bad), C.byref(rel), stream)
        if code not in (0, -6):
            self._run_checked(code)
        return code == 0, int(bad.value), float(rel.value)


def run_cli(args, timeout=None) -> subprocess.CompletedProcess:
    """ft_sgemm START END GAP ST_KERNEL END_KERNEL  (sgemm.cu:13-19)."""
    if not CLI_PATH.exists():
        raise FtsgemmError(-4, f"{CLI_PATH} not built")
    return subprocess.run([str(CLI_PATH), *map(str, args)], capture_output=True, text=True, timeout=timeout)
