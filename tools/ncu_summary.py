"""Summarise ncu --set full reports (gpurun_out/prof_*.ncu-rep) into one table: per launch duration, DRAM bytes,
tensor-pipe / issue utilisation, occupancy limits.  Run in the build container: python tools/ncu_summary.py <reps...>"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem_KB"), ("launch__grid_size", "grid"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs")]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "no data")
            continue
        h, units = rows[0], rows[1]
        kn = h.index("Kernel Name")
        print(f"== {path}")
        for r in rows[2:]:
            name = r[kn].split("(")[0][:60]
            vals = []
            for k, label in KEYS:
                if k in h:
                    i = h.index(k)
                    v = r[i].replace(",", "")
                    try:
                        f = float(v)
                        u = units[i]
                        if label.endswith("_MB") and u == "Kbyte":
                            f /= 1e3
                        if label.endswith("_MB") and u == "byte":
                            f /= 1e6
                        if label == "us" and u == "ns":
                            f /= 1e3
                        if label == "us" and u == "ms":
                            f *= 1e3
                        if label.endswith("_KB") and u == "byte/block":
                            f /= 1e3
                        v = f"{f:.1f}"
                    except ValueError:
                        pass
                    vals.append(f"{label}={v}")
            print(f"  {name:60s} " + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1:])
