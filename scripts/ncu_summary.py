"""Summarise an .ncu-rep (raw page) into one line per launch: the metrics the roofline discussion needs.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("lts__t_bytes.sum", "l2MB"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smemKB"), ("sm__cycles_elapsed.max", "cycles")]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"# {path}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]][:60]
            parts = []
            for key, label in KEYS:
                if key not in col:
                    continue
                v, u = to_float(r[col[key]]), units[col[key]]
                if v is None:
                    continue
                if label.endswith("MB"):
                    v = v / 1e6 if u == "byte" else (v / 1e3 if u == "Kbyte" else (v * 1e3 if u == "Gbyte" else v))
                if label == "us" and u == "ns":
                    v /= 1e3
                if label == "us" and u == "ms":
                    v *= 1e3
                if label == "smemKB" and u == "byte/block":
                    v /= 1e3
                parts.append(f"{label}={v:.1f}" if isinstance(v, float) and not v.is_integer() else f"{label}={int(v)}")
            print(f"{name:60s} " + " ".join(parts))


if __name__ == "__main__":
    main()
