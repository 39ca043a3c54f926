"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file = ""
    hdr = None
    recs = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr and len(r) > 8 and r[0].isdigit():
            try:
                recs.append((int(r[si] or 0), int(r[ii] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    ts, ti = sum(x[0] for x in recs), sum(x[1] for x in recs)
    print("total samples %d, warp instructions %d" % (ts, ti))
    print("-- by samples")
    for s, i, f, ln, src in sorted(recs, reverse=True)[:top]:
        print("%6d %5.1f%% %10d %5.1f%% %s:%d  %s" % (s, 100.0 * s / max(ts, 1), i, 100.0 * i / max(ti, 1), f, ln, src))
    print("-- by instructions")
    for s, i, f, ln, src in sorted(recs, key=lambda x: -x[1])[:top // 2]:
        print("%6d %5.1f%% %10d %5.1f%% %s:%d  %s" % (s, 100.0 * s / max(ts, 1), i, 100.0 * i / max(ti, 1), f, ln, src))


if __name__ == "__main__":
    main()
