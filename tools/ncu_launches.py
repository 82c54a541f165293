"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv)
by kernel: launches, summed / average duration, share of the listed kernel time, DRAM bytes per launch.  With --traffic
OUT.json also writes the average DRAM bytes per launch of the conv engine (dfgan_conv_run's kernels: conv_umma,
conv_halo, conv_halo2), the figure bench.py reports as roofline.traffic.

    python tools/ncu_launches.py gpurun_out/launches.csv [--traffic profiles/rNN_traffic.json] [--source "text"]
"""
import collections
import csv
import io
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name).replace("dfgan::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)
    if name.startswith("at::"):
        return "at:: (torch glue: fills, copies, cat, mean)"
    return name[:56]


def main(argv):
    path = argv[1]
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO(txt)):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["dram"] = d.get("dram", 0.0) + v * scale
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        a = agg[short(d["name"])]
        a[0] += 1
        a[1] += d.get("us", 0.0)
        a[2] += d.get("dram", 0.0)
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if not k.startswith("at::"))
    print(f"{len(per)} launches, {total / 1e3:.2f} ms of kernel time, {100 * ours / total:.1f} % in dfgan:: kernels")
    print(f"{'kernel':58s} {'n':>5s} {'sum us':>9s} {'avg us':>8s} {'share':>7s} {'DRAM MB/launch':>15s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:58s} {a[0]:5d} {a[1]:9.1f} {a[1] / a[0]:8.1f} {100 * a[1] / total:6.1f}% {a[2] / a[0] / 1e6:15.2f}")
    if "--traffic" in argv:
        out = argv[argv.index("--traffic") + 1]
        src = argv[argv.index("--source") + 1] if "--source" in argv else path
        conv = [d for d in per.values() if re.search(r"conv_umma_kernel|conv_halo_kernel|conv_halo2_kernel", d["name"])]
        n = len(conv)
        json.dump({"engine": "dfgan_conv_run (conv_umma_kernel, conv_halo_kernel, conv_halo2_kernel)", "launches": n,
                   "dram_bytes_per_launch": sum(d.get("dram", 0.0) for d in conv) / max(n, 1),
                   "kernel_us_per_launch": sum(d.get("us", 0.0) for d in conv) / max(n, 1),
                   "source": src}, open(out, "w"), indent=1)
        print(f"conv engine: {n} launches, {sum(d.get('dram', 0.0) for d in conv) / max(n, 1) / 1e6:.1f} MB DRAM per launch "
              f"-> {out}")


if __name__ == "__main__":
    main(sys.argv)
