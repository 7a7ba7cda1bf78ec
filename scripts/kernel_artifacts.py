"""Per-kernel artefacts (north_star: "each kernel committed with SASS listing and ncu capture").

  python scripts/kernel_artifacts.py sass          # here, no GPU: cuobjdump of every instantiated kernel of libftsgemm.so
                                                   #   -> profiles/r02_sass_<kernel>.txt (mnemonic histogram + the tcgen05 /
                                                   #      TMA / TMEM instruction lines) and profiles/r02_sass_summary.json
  python scripts/kernel_artifacts.py ncu [n=4096]  # under gpurun (1 GPU): one `ncu --set full` capture per kernel id at n^3
                                                   #   -> gpurun_out/r02_ncu_id<k>_<n>.ncu-rep
  python scripts/kernel_artifacts.py summarize     # here: reads the .ncu-rep files -> profiles/r02_ncu_summary.json (tensor
                                                   #   pipe %, DRAM bytes / %, duration, registers) and profiles/traffic.json
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "fault-tolerant-sgemm-on-nvidia-gpus_b200" / "libftsgemm.so"
PROF = ROOT / "profiles"
OUT = ROOT / "gpurun_out"
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMAPF", "SYNCS", "HMMA",
       "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "MEMBAR", "ATOM", "RED", "DADD", "ELECT")
# kernel id -> (template instantiation, role) : one id per distinct binary
IDS = {1: "128x64 plain cg1 (small)", 2: "256x64 plain cg2 (medium)", 3: "256x128 plain cg2 (large / pair128)", 4: "128x32 plain cg1 (tall)",
       5: "128x256 plain cg1 (wide)", 6: "128x128 plain cg1 (huge)", 21: "256x256 plain cg2 (giant)",
       11: "128x64 abft cg1", 12: "256x64 abft cg2", 13: "256x128 abft cg2", 14: "128x32 abft cg1", 15: "128x256 abft cg1",
       16: "128x128 abft cg1 (config 2 literal)", 31: "256x256 abft cg2 (bench)"}


def demangle(name):
    m = re.search(r"ftsgemm_tc_kernelILi(\d+)ELb(\d)ELi(\d)ELb(\d)", name)
    if m:
        return f"ftsgemm_tc_kernel_{m.group(1)}_{'ft' if m.group(2) == '1' else 'plain'}_cg{m.group(3)}{'_prot' if m.group(4) == '1' else ''}"
    m = re.search(r"encode_b_kernelILi(\d+)ELi(\d+)ELb(\d)", name)
    if m:
        return f"encode_b_kernel_{m.group(1)}_kr{m.group(2)}{'_stream' if m.group(3) == '1' else ''}"
    m = re.search(r"_ZN\d+_GLOBAL__N_[^\d]*\d+(\w+?)E", name)
    return re.sub(r"\W+", "_", name)[:60]


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    summary = {}
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        short = demangle(name)
        ops = Counter()
        keep = []
        for line in f.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if not m:
                continue
            op = m.group(1)
            ops[op.split(".")[0]] += 1
            if op.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLKCP")):
                keep.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.strip()))
        total = sum(ops.values())
        row = {k: ops.get(k, 0) for k in KEY if ops.get(k, 0)}
        row["instructions"] = total
        summary[short] = row
        if short.startswith(("ftsgemm_tc_kernel", "encode_b_kernel_256_kr8")):
            lines = [f"# {short}   ({name})", f"# {total} SASS instructions; histogram of the mnemonics that matter:",
                     "# " + ", ".join(f"{k} {v}" for k, v in row.items() if k != "instructions"), "#",
                     "# tcgen05 / TMA / tensor-memory instruction lines (cuobjdump -sass, sm_100a):"] + keep
            (PROF / f"r02_sass_{short}.txt").write_text("\n".join(lines) + "\n")
    (PROF / "r02_sass_summary.json").write_text(json.dumps(summary, indent=1, sort_keys=True))
    for k, v in sorted(summary.items()):
        if k.startswith("ftsgemm_tc_kernel"):
            print(k, {x: v[x] for x in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "instructions") if x in v})


def ncu(n):
    OUT.mkdir(exist_ok=True)
    ids = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else sorted(IDS)
    for kid in ids:
        rep = OUT / f"r02_ncu_id{kid}_{n}"
        cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:ftsgemm_tc_kernel",
               "-s", "2", "-c", "1", "-f", "-o", str(rep), sys.executable, str(ROOT / "scripts" / "run_one.py"), str(kid), str(n), "3"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(kid, "ok" if r.returncode == 0 else r.stderr[-300:], flush=True)
    # the encode pre-pass of the bench kernel
    rep = OUT / f"r02_ncu_encode_{n}"
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "-k", "regex:encode_b_kernel", "-s", "2", "-c", "1", "-f", "-o",
                    str(rep), sys.executable, str(ROOT / "scripts" / "run_one.py"), "31", str(n), "3", "enc_front=0"], capture_output=True, text=True)
    # gpurun brings back at most 64 MiB: summarise on the box, keep only the reports of the bench kernel and its plain twin
    summarize()
    for rep in OUT.glob("r02_ncu_*.ncu-rep"):
        if not re.search(r"id(31|21)_", rep.name):
            rep.unlink()


METRICS = {"gpu__time_duration.sum": "duration_us", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed",
           "sm__inst_executed_pipe_tensor.sum": "tensor_inst", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
           "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "sm__cycles_elapsed.avg.per_second": "sm_hz"}
UNIT = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def mode_is_ncu():
    return len(sys.argv) > 1 and sys.argv[1] == "ncu"


def summarize():
    out, traffic = {}, {}
    for rep in sorted(OUT.glob("r02_ncu_*.ncu-rep")):
        r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for i, h in enumerate(hdr):
            if h in METRICS:
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                rec[METRICS[h]] = v * UNIT.get(units[i], 1.0) if units[i] in UNIT else v
        key = rep.stem.replace("r02_ncu_", "")
        m0 = re.match(r"id(\d+)_", key)
        if m0 and int(m0.group(1)) in IDS:
            rec["variant"] = IDS[int(m0.group(1))]
        if "dram_read_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
        out[key] = rec
        m = re.match(r"id(\d+)_(\d+)", key)
        if m and "dram_bytes" in rec:
            traffic.setdefault(m.group(1), {})[m.group(2)] = int(rec["dram_bytes"])
    dest = OUT if mode_is_ncu() else PROF
    (dest / "r02_ncu_summary.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    tp = dest / "traffic.json"
    old = {}
    if tp.exists():
        try:
            old = json.loads(tp.read_text())
        except Exception:
            old = {}
    for k, v in old.items():  # keep earlier single-value entries as the 4096 column
        if not isinstance(v, dict):
            old[k] = {"4096": v}
    for k, v in traffic.items():
        old.setdefault(k, {}).update(v)
    tp.write_text(json.dumps(old, indent=1, sort_keys=True))
    for k, v in sorted(out.items()):
        print(k, {x: (round(y, 2) if isinstance(y, float) else y) for x, y in v.items() if x != "kernel"})


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "sass"
    if mode == "sass":
        sass()
    elif mode == "ncu":
        ncu(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
    else:
        summarize()
