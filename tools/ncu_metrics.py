"""Extract the counters bench.py reports as `from_profile` out of `ncu --set full` captures into
profiles/r02_ncu_metrics.json (kernel name -> counters, with the capture file and the git revision).

    python tools/ncu_metrics.py gpurun_out/r02_*.ncu-rep
"""
import csv, io, json, os, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        rec = {}
        for k, short in KEYS.items():
            if k in d and d[k] not in ("", "n/a"):
                v = float(d[k].replace(",", ""))
                u = units[hdr.index(k)]
                rec[short] = v * UNIT_SCALE.get(u, 1)
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes"] = rec.pop("dram_read") + rec.pop("dram_write")
        if "duration" in rec:
            rec["duration_us"] = rec.pop("duration")
        rec["grid"], rec["block"] = d.get("Grid Size"), d.get("Block Size")
        res.setdefault(name, rec)            # first launch of each kernel in the capture
    return res


def main():
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=root).stdout.strip()
    kernels, caps = {}, []
    for rep in sys.argv[1:]:
        for name, rec in parse(rep).items():
            rec["capture"] = os.path.basename(rep)
            kernels[name] = rec
        caps.append(os.path.basename(rep))
    path = os.path.join(root, "profiles", "r02_ncu_metrics.json")
    with open(path, "w") as f:
        json.dump({"git": git, "capture": caps, "how": "ncu --set full --clock-control none, one B200, see profiles/README.md",
                   "kernels": kernels}, f, indent=1)
    print("wrote", path, list(kernels))


if __name__ == "__main__":
    main()
