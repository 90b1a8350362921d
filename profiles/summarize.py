"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/r01_launches_X.txt
    python profiles/summarize.py full     gpurun_out/prof_X.ncu-rep  profiles/r01_full_X.txt
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = row["Kernel Name"].split("(")[0][:90]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES\n")
        f.write("# source: %s   total %.1f us over %d launches\n" % (src, tot, sum(v[0] for v in agg.values())))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s n=%5d total_us=%10.1f avg_us=%9.2f share=%.4f\n" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))


WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, source: %s\n" % src)
        for r in rows[2:]:
            f.write("---\n")
            for w, i in idx:
                f.write("%-70s %s %s\n" % (w, r[i][:100], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
