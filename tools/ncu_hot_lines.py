#!/usr/bin/env python
"""Per-source-line attribution of an `ncu --set full --import-source on` capture.

ncu's CLI only prints the SASS view of the source page; this joins it, instruction by instruction, with the line table that
`nvdisasm -g` prints for the same kernel in librepaq_b200.so (the library must be the build the capture was taken from).

usage: ncu_hot_lines.py capture.ncu-rep kernel_name [top_n]
prints: warp instructions, share, stall samples per source line (innermost inlined location), sorted by instructions.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "repaq_b200", "librepaq_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, inside = None, False
        for line in txt.splitlines():
            if line.startswith(".text."):
                inside = re.search(r"\d+%s[A-Z]" % re.escape(kernel), line) is not None
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                if "inlined at" not in line or cur is None:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip(), cur))
        if out:
            break
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    # a report may hold many kernels: take the first launch of the one asked for
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:^%s$" % kernel, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[hi + 1:]:
        if r and r[0] == "Address":
            break                                            # the next launch of the same kernel
        if len(r) == len(hdr):
            body.append(r)
    sass = sass_lines(kernel)
    if len(sass) != len(body):
        print("warning: %d instructions in the capture, %d in the library (stale build?)" % (len(body), len(sass)))
    agg = {}
    tot_i = tot_s = 0
    for k, r in enumerate(body):
        loc = sass[k][2] if k < len(sass) else ("?", 0)
        ins = int(r[col["Instructions Executed"]])
        smp = int(r[col["# Samples"]])
        thr = int(r[col["Thread Instructions Executed"]])
        a = agg.setdefault(loc, [0, 0, 0])
        a[0] += ins
        a[1] += smp
        a[2] += thr
        tot_i += ins
        tot_s += smp
    print("kernel %s: %d warp instructions, %d stall samples" % (kernel, tot_i, tot_s))
    src_cache = {}

    def src(loc):
        f, n = loc
        if f not in src_cache:
            p = os.path.join(ROOT, "repaq_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        L = src_cache[f]
        return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
    print("%-24s %10s %6s %6s %5s  %s" % ("file:line", "warp inst", "inst%", "smpl%", "thr", "source"))
    for loc, (ins, smp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-24s %10d %6.2f %6.2f %5.1f  %s" % ("%s:%d" % loc, ins, 100.0 * ins / max(1, tot_i), 100.0 * smp / max(1, tot_s), thr / max(1, ins), src(loc)))
    print("---- by stall samples")
    for loc, (ins, smp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-24s %10d %6.2f %6.2f %5.1f  %s" % ("%s:%d" % loc, ins, 100.0 * ins / max(1, tot_i), 100.0 * smp / max(1, tot_s), thr / max(1, ins), src(loc)))


if __name__ == "__main__":
    main()
