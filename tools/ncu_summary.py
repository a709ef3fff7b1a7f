"""Summarise an `ncu --set full` capture of one solve into profiles/:

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/rNN_ncu_full_cfg2.csv [--traffic cfg2]

One line per kernel launch (grid, registers, dynamic smem, duration, DRAM bytes,
instructions, active warps); --traffic WORKLOAD also records the summed DRAM
bytes of the solve kernels as profiles/traffic.json[WORKLOAD] (bench.py's
roofline.traffic)."""
import csv
import json
import os
import re
import sys

COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem_kb"), ("gpu__time_duration.sum", "duration_us"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("smsp__inst_executed.sum", "inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def unit(u):
    u = u.split("/")[0]
    return UNIT.get(u, 1.0)


def num(x):
    return float(x.replace(",", "")) if x not in ("", "n/a") else float("nan")


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(k) for k, _ in COLS}
    total = 0.0
    with open(out, "w") as f:
        f.write("kernel,grid,regs,dyn_smem_kb,duration_us,dram_bytes,inst,warps_active_pct\n")
        for r in body:
            name = re.sub(r"\(.*$", "", r[idx["Kernel Name"]]).replace("lfr::", "").replace("(int)", "")
            name = re.sub(r"^void ", "", name)
            grid = r[idx["Grid Size"]].replace(",", "").split()[0].strip("()")
            dur = num(r[idx["gpu__time_duration.sum"]]) * unit(units[idx["gpu__time_duration.sum"]])
            rd = num(r[idx["dram__bytes_read.sum"]]) * unit(units[idx["dram__bytes_read.sum"]])
            wr = num(r[idx["dram__bytes_write.sum"]]) * unit(units[idx["dram__bytes_write.sum"]])
            smem = num(r[idx["launch__shared_mem_per_block_dynamic"]]) * unit(
                units[idx["launch__shared_mem_per_block_dynamic"]]) / 1e3
            f.write('"%s",%s,%s,%.3f,%.3f,%d,%d,%.3f\n' % (
                name, grid, r[idx["launch__registers_per_thread"]], smem, dur, (int(rd + wr) if rd == rd and wr == wr else 0),
                int(num(r[idx["smsp__inst_executed.sum"]])),
                num(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]])))
            if "solve_" in name:
                total += rd + wr
    if "--traffic" in sys.argv:
        w = sys.argv[sys.argv.index("--traffic") + 1]
        tp = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {"dram_bytes_per_step": {}}
        t["dram_bytes_per_step"][w] = int(total)
        t.pop("launches", None)
        t.setdefault("sources", {})[w] = os.path.basename(out) + " (ncu --set full --clock-control none; " \
            "dram__bytes_read.sum + dram__bytes_write.sum summed over the solve launches of one step)"
        t.pop("source", None)
        json.dump(t, open(tp, "w"), indent=1)
        print("traffic", w, int(total))


if __name__ == "__main__":
    main()
