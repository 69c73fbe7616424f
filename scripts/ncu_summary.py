#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: per launch the duration, DRAM
traffic (dram__bytes_read.sum + dram__bytes_write.sum), L2->SM bytes, tensor-pipe and SM
utilisation.  usage: ncu_summary.py report.ncu-rep [label ...]   (labels name the launches in order)"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tmem.sum", "tmem_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster"),
]


def scale(v, u):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    return v * mult.get(u, 1.0)


def main():
    rep, labels = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [(m, n) for m, n in COLS if m in col]
    print("# %s" % rep)
    print("# time us | DRAM read / write MB (traffic = sum) | L2->SM MB | percentages of peak")
    print("%-3s %-44s " % ("#", "kernel [label]") + " ".join("%9s" % n for _, n in have))
    for k, r in enumerate(body):
        name = r[col["Kernel Name"]]
        name = name[name.find("::") + 2:] if "unnamed" in name else name
        name = name.split("(")[0][:30] + (" [%s]" % labels[k] if k < len(labels) else "")
        vals = []
        for m, n in have:
            v = scale(r[col[m]], units[col[m]])
            if "byte" in units[col[m]]:
                v /= 1e6
            vals.append("%9.2f" % v)
        print("%-3d %-44s " % (k, name) + " ".join(vals))


if __name__ == "__main__":
    main()
