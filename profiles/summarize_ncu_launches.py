"""Summarise an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch; tests/gpu_batch.sh) per kernel: launches, summed time, share of the transform, DRAM bytes.

    python profiles/summarize_ncu_launches.py gpurun_out/r2g/launches_C4.csv > profiles/r02_launches_C4_summary.md
"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3}[unit]
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            d["rd" if "read" in r["Metric Name"] else "wr"] = v * scale
    return list(launches.values())


def short(name):
    m = re.match(r"(?:void )?([A-Za-z_0-9]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:40]
    t = m.group(2) or ""
    t = t.replace("unsigned long", "u64").replace("unsigned int", "u32").replace("(bool)", "").replace("(int)", "")
    return m.group(1) + t


def main():
    L = load(sys.argv[1])
    tot = sum(x.get("ms", 0) for x in L)
    agg = collections.OrderedDict()
    for x in L:
        a = agg.setdefault(short(x["name"]), {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1; a["ms"] += x.get("ms", 0); a["rd"] += x.get("rd", 0); a["wr"] += x.get("wr", 0)
    print(f"# ncu launch list: {sys.argv[1].split('/')[-1]} -- {len(L)} launches, {tot:.2f} ms of kernel time (one forward + one inverse,")
    print("# cold caches, serialised by the profiler: compare SHARES with the CUDA-event classes of bench.py, not absolutes)\n")
    print("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | DRAM GB/s |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = (a["rd"] + a["wr"]) / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0
        print(f"| `{k}` | {a['n']} | {a['ms']:.3f} | {100 * a['ms'] / tot:.1f} % | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} | {gbs:.0f} |")


if __name__ == "__main__":
    main()
