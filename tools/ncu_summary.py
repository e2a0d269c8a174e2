"""Key metrics of `ncu --set full` reports as CSV rows (one per captured launch).
Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/summary.csv"""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]

def main(paths):
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + KEYS)
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            w.writerow([path.split("/")[-1], d.get("Kernel Name", "")[:70]] +
                       [f"{d.get(k, '')} {u.get(k, '')}".strip() for k in KEYS])

if __name__ == "__main__":
    main(sys.argv[1:])
