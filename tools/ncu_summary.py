"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/:  python tools/ncu_summary.py in.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        e = {}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                e[k + (f" [{units[i]}]" if units[i] else "")] = r[i]
        out.append(e)
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(out)} kernels -> {dst}")
    if len(sys.argv) > 3:  # third argument: traffic table (bench.py's roofline.traffic), mean DRAM bytes per launch
        acc = {}
        ir, iw, iu = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), None
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in data:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            b = float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
            acc.setdefault(name, []).append(b)
        tr = {"source": src.split("/")[-1], "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items()}}
        json.dump(tr, open(sys.argv[3], "w"), indent=1)
        print(f"traffic of {len(acc)} kernels -> {sys.argv[3]}")


if __name__ == "__main__":
    main()
