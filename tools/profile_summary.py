"""Turn the ncu outputs of a bench.py run into the tracked files under profiles/.

usage: python tools/profile_summary.py <launches.csv from `ncu --metrics gpu__time_duration.sum --csv --log-file`> <out prefix>
  writes <prefix>_launches_bench.csv (id, kernel, block, grid, gpu_time_ns) and <prefix>_launches_bench_summary.md
"""
import csv
import sys


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else "python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = [(int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Block Size"]], r[ix["Grid Size"]], int(float(r[ix["Metric Value"]])))
                for r in rows if r[ix["Metric Name"]] == "gpu__time_duration.sum"]
    with open(prefix + "_launches_bench.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "gpu_time_ns"])
        w.writerows(launches)
    tot = {}
    for _, k, _, _, ns in launches:
        n, t = tot.get(k, (0, 0))
        tot[k] = (n + 1, t + ns)
    total = sum(t for _, t in tot.values())
    with open(prefix + "_launches_bench_summary.md", "w") as f:
        f.write("# Launch list of `%s` (ncu --metrics gpu__time_duration.sum --clock-control none)\n" % cmd)
        f.write("Per-kernel totals over the whole process (warm-ups, the timed steps, the e2e loop, the pipelined extra); cold-cache, serialised.\n\n")
        f.write("| kernel | launches | total us | share | mean us |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% | %.1f |\n" % (k[:90], n, t / 1e3, 100.0 * t / total, t / 1e3 / n))


if __name__ == "__main__":
    main()
