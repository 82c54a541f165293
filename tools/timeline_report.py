"""Report on a kernel timeline written by tools/prof_timeline.py: per-iteration span, busy / idle time, concurrency,
and for every kernel class its total duration and its EXCLUSIVE time (wall time during which nothing else runs --
the part a faster kernel would give back to the iteration directly).

    python tools/timeline_report.py gpurun_out/timeline.csv
"""
import collections
import csv
import re
import sys


def cls(name):
    name = re.sub(r"^void ", "", name).replace("dfgan::", "")
    name = re.sub(r"\(.*", "", name)
    return name[:60]


def main(path):
    rows = [(float(r["start_us"]), float(r["dur_us"]), r["name"]) for r in csv.DictReader(open(path))]
    # split into iterations at the input copies (three DtoD memcpys open each DFGANTrainer.step)
    starts = [i for i, r in enumerate(rows) if r[2].startswith("Memcpy DtoD") and r[1] > 5.0]
    iters = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]
    print(f"{len(rows)} kernels, {len(iters)} iterations")
    for it_i, it in enumerate(iters):
        t0, t1 = it[0][0], max(s + d for s, d, _ in it)
        ev = []
        for s, d, n in it:
            ev.append((s, 1, n))
            ev.append((s + d, -1, n))
        ev.sort(key=lambda e: (e[0], e[1]))
        active, last, busy, conc_time = collections.Counter(), t0, 0.0, collections.Counter()
        excl = collections.Counter()
        for t, k, n in ev:
            na = sum(active.values())
            if na > 0:
                busy += t - last
                conc_time[min(na, 4)] += t - last
                if na == 1:
                    excl[cls(next(x for x, c in active.items() if c > 0))] += t - last
            last = t
            active[n] += k
        tot = collections.Counter()
        cnt = collections.Counter()
        for s, d, n in it:
            tot[cls(n)] += d
            cnt[cls(n)] += 1
        span = t1 - t0
        print(f"\n== iteration {it_i}: span {span / 1e3:.3f} ms, {len(it)} kernels, busy {busy / 1e3:.3f} ms, idle "
              f"{(span - busy) / 1e3:.3f} ms, sum of durations {sum(d for _, d, _ in it) / 1e3:.3f} ms")
        print("   wall time by number of concurrently running kernels: " +
              ", ".join(f"{k}{'+' if k == 4 else ''}: {v / 1e3:.2f} ms" for k, v in sorted(conc_time.items())))
        if it_i != len(iters) - 1:
            continue
        print(f"   {'kernel':62s} {'n':>4s} {'sum us':>9s} {'avg us':>8s} {'exclusive us':>13s}")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:32]:
            print(f"   {k:62s} {cnt[k]:4d} {v:9.1f} {v / cnt[k]:8.1f} {excl[k]:13.1f}")
        print(f"   exclusive total {sum(excl.values()) / 1e3:.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1])
