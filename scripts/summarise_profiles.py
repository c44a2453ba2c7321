"""Condenses the ncu raw-page CSVs of scripts/profile_r2.sh into profiles/<tag>_ncu_counters.csv (one row per captured
launch, the counters the roofline discussion uses) and prints a markdown table.

    python scripts/summarise_profiles.py <tag> gpurun_out/<tag>_cfg2.raw.csv [more.raw.csv ...]
"""
import csv
import os
import sys

WANT = [
    ("gpu__time_duration.sum", "duration_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu_shared_wavefront_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared_bank_conflicts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def conv(v, unit, name):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    if name.endswith("_MB"):
        scale = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
        return round(x * scale, 3)
    if name == "duration_us":
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        return round(x * scale, 2)
    return round(x, 3)


def main():
    tag, files = sys.argv[1], sys.argv[2:]
    out_rows = []
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").replace("kccot::", "").replace("<unnamed>::", "")
            rec = {"capture": os.path.basename(f).replace(".raw.csv", ""), "kernel": short}
            for metric, key in WANT:
                if metric in col:
                    rec[key] = conv(r[col[metric]], units[col[metric]], key)
            out_rows.append(rec)
    keys = ["capture", "kernel"] + [k for _, k in WANT]
    os.makedirs("profiles", exist_ok=True)
    path = f"profiles/{tag}_ncu_counters.csv"
    with open(path, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=keys)
        w.writeheader()
        for rec in out_rows:
            w.writerow(rec)
    print(f"wrote {path} ({len(out_rows)} launches)")
    show = ["kernel", "duration_us", "grid", "dram_read_MB", "dram_write_MB", "dram_pct", "tensor_pipe_pct",
            "lsu_shared_wavefront_pct", "l2_hit_pct", "sm_pct", "regs"]
    print("| " + " | ".join(show) + " |")
    print("|" + "---|" * len(show))
    for rec in out_rows:
        print("| " + " | ".join(str(rec.get(k, "")) for k in show) + " |")


if __name__ == "__main__":
    main()
