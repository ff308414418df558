"""Summarise an .ncu-rep (read here, no GPU): per kernel the metrics the roofline uses."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[name_col][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-70s %s %s" % (k, r[i], units[i]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
