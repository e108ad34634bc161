#!/usr/bin/env python
"""Per-source-line instruction and stall shares of one kernel in an .ncu-rep
(captured with --set full --import-source on, code built with -lineinfo).

    python tools/ncu_lines.py report.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys
import collections


def page(rep, view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view],
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    sass = page(rep, "sass")
    hi = [i for i, r in enumerate(sass) if "Instructions Executed" in r][0]
    h = sass[hi]
    ci, wi = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    addr_stat = collections.OrderedDict()
    for r in sass[hi + 1:]:
        if len(r) > ci and r[0].startswith("0x"):
            addr_stat[int(r[0], 16)] = (int(r[ci] or 0), int(r[wi] or 0), r[1].strip())
    addrs = list(addr_stat)
    mixed = page(rep, "cuda,sass")
    hi = [i for i, r in enumerate(mixed) if "Instructions Executed" in r][0]
    line_of = {}
    cur, last, gap = None, None, False
    for r in mixed[hi + 1:]:
        if r[0].strip().isdigit():
            cur, last, gap = (int(r[0]), r[1].strip()), None, False
            continue
        if len(r) > 2 and r[2] == "...":
            gap = True
            continue
        if len(r) > 2 and r[2].startswith("0x") and cur:
            a = int(r[2], 16)
            if gap and last is not None:
                for x in addrs:
                    if last < x < a:
                        line_of[x] = cur
            line_of[a] = cur
            last, gap = a, False
    per = collections.OrderedDict()
    tot_i = sum(v[0] for v in addr_stat.values()) or 1
    tot_w = sum(v[1] for v in addr_stat.values()) or 1
    for a, (n, w, txt) in addr_stat.items():
        key = line_of.get(a, (-1, "?"))
        d = per.setdefault(key, [0, 0, 0])
        d[0] += n
        d[1] += w
        d[2] += 1
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_w))
    for (ln, src), (n, w, k) in sorted(per.items(), key=lambda kv: kv[0][0]):
        if 100.0 * n / tot_i >= min_pct or 100.0 * w / tot_w >= min_pct:
            print("%5d  inst %5.1f%%  stall %5.1f%%  sass %4d  %s" % (ln, 100.0 * n / tot_i,
                                                                     100.0 * w / tot_w, k, src[:100]))


if __name__ == "__main__":
    main()
