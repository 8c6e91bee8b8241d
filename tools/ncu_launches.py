"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
one line per launch of our kernels with its duration, DRAM bytes and DRAM rate.   python tools/ncu_launches.py FILE.csv [min_ms]"""
import collections
import csv
import io
import sys


def main(path, min_ms=0.02):
    txt = open(path).read()
    rd = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    by = collections.OrderedDict()
    for r in rd:
        k = (int(r["ID"]), r["Kernel Name"])
        by.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    scale_t = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for (i, name), m in by.items():
        t, tu = m.get("gpu__time_duration.sum", (0.0, "ns"))
        ms = t * scale_t.get(tu, 1e-6)
        tot = sum(v * scale_b.get(u, 1.0) for key, (v, u) in m.items() if key.startswith("dram__bytes"))
        if ms >= min_ms:
            print(f"{i:>4} {name.split('(')[0][:64]:64s} {ms:8.3f} ms {tot / 1e9:7.3f} GB {tot / 1e9 / ms:6.2f} TB/s")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.02)
