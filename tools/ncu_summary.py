"""Condense an `ncu --page raw --csv` export into a small per-launch table (kept under profiles/)."""
import csv, sys

COLS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
    return v * scale.get(unit, 1.0)


def main(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as fh:
        fh.write("# kernel | " + " | ".join(n for _, n in COLS) + "\n")
        for d in data:
            name = d[ki].split("(")[0].replace("void ", "")
            vals = []
            for col, _ in COLS:
                if col not in hdr:
                    vals.append("-")
                    continue
                i = hdr.index(col)
                try:
                    vals.append(f"{to_unit(d[i], units[i], None):.1f}")
                except ValueError:
                    vals.append(d[i])
            fh.write(name + " | " + " | ".join(vals) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
