"""Prints the kernels of the last full learn step found in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_list.py gpurun_out/launches.csv"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].isdigit()]
    idx = [i for i, r in enumerate(data) if "sample_gather" in r[kn]]
    seg = data[idx[-2]:idx[-1]] if len(idx) >= 2 else data
    tot = 0.0
    for r in seg:
        t = float(r[mv].replace(",", "")) / 1000.0
        tot += t
        print("%8.1f us  %-16s %s" % (t, r[gs], r[kn][:100]))
    print("total %.1f us over %d launches" % (tot, len(seg)))


if __name__ == "__main__":
    main(sys.argv[1])
