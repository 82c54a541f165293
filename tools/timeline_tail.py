"""Kernels of one CUDA-graph replay (tools/prof_timeline.py CSV) whose EXCLUSIVE time -- wall time during which nothing
else runs -- exceeds a threshold, in time order: where a faster / better-placed kernel gives time back directly.

    python tools/timeline_tail.py gpurun_out/timeline.csv [min_exclusive_us] [name filter regex]
"""
import csv
import io
import re
import sys


def main(path, thr=40.0, pat=None):
    txt = open(path).read()
    txt = txt[txt.index("start_us,dur_us,name"):]
    rows = [(float(r["start_us"]), float(r["dur_us"]), r["name"]) for r in csv.DictReader(io.StringIO(txt))]
    starts = [i for i, r in enumerate(rows) if r[2].startswith("Memcpy DtoD") and r[1] > 5.0]
    a, b = starts[1], starts[2] if len(starts) > 2 else len(rows)
    it = rows[a:b]
    t0 = it[0][0]
    ev = []
    for i, (s, d, _) in enumerate(it):
        ev += [(s, 1, i), (s + d, -1, i)]
    ev.sort()
    active, last, excl = set(), None, [0.0] * len(it)
    for t, k, i in ev:
        if last is not None and len(active) == 1:
            excl[next(iter(active))] += t - last
        (active.add if k == 1 else active.discard)(i)
        last = t

    def cls(n):
        n = re.sub(r"^void ", "", n).replace("dfgan::", "").replace("(anonymous namespace)::", "")
        return re.sub(r"\(.*", "", n)[:44]
    print(f"iteration span {max(s + d for s, d, _ in it) - t0:.1f} us, {len(it)} kernels")
    tot = 0.0
    for (s, d, n), e in zip(it, excl):
        if (pat and re.search(pat, n)) or (not pat and e > thr):
            print(f"{s - t0:9.1f} dur {d:7.1f} excl {e:7.1f}  {cls(n)}")
            tot += e
    print(f"sum of listed exclusive times {tot:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 40.0, sys.argv[3] if len(sys.argv) > 3 else None)
