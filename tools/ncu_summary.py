"""Pull the per-kernel numbers quoted in profiles/README.md and profiles/traffic.json out of
`ncu --page raw --csv` files.  usage: ncu_summary.py name=file.csv ..."""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.per_cycle_active": "warps_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "sm__sass_inst_executed_op_shared_ld.sum": "lds",
    "smsp__inst_executed_pipe_fma.sum": "fma_inst",
    "smsp__inst_executed_pipe_fmaheavy.sum": "fmaheavy_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}


def to_float(v, unit):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
             "byte": 1.0}.get(unit)
    return v * scale if scale else v


def main():
    out = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=")
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
        d = {"kernel": vals[names.index("Kernel Name")]}
        for k, short in WANT.items():
            if k in names:
                i = names.index(k)
                try:
                    d[short] = to_float(vals[i], units[i])
                except ValueError:
                    d[short] = vals[i]
        out[name] = d
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
