#!/usr/bin/env python
"""Per-source-line stall-sample summary of one kernel from an .ncu-rep (no GPU needed).

    python profiles/ncu_hotspots.py gpurun_out/prof.ncu-rep k_match [top_n]

ncu's CSV source page is per SASS instruction; the line table of the cubin inside
deepdish_b200/libdeepdish_b200.so (nvdisasm -g) maps instruction offsets back to file:line.
The .so must be the build that was profiled."""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "deepdish_b200", "libdeepdish_b200.so")],
                   cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin") or f.count("-"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        in_k, cur = False, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                in_k = kernel in m.group(1)
                continue
            if not in_k:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")] or []
    table = line_table(kernel)
    base = None
    per_line = collections.Counter()
    per_line_inst = collections.Counter()
    stalls = collections.Counter()
    total = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= si or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break           # first launch only
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        n = int(r[si] or 0)
        total += n
        loc = table.get(addr - base, (None, r[1].strip()))[0]
        per_line[loc] += n
        per_line_inst[loc] += int(r[ii] or 0)
        for c in stall_cols:
            try:
                stalls[hdr[c]] += int(r[c] or 0)
            except ValueError:
                pass
    print("kernel %s: %d stall samples, %d warp-instructions" % (kernel, total, sum(per_line_inst.values())))
    for k, v in stalls.most_common(8):
        print("  %-28s %6.1f%%" % (k, 100.0 * v / max(1, sum(stalls.values()))))
    for loc, n in per_line.most_common(top):
        print("%6d %5.1f%%  inst=%9d  %s" % (n, 100.0 * n / max(1, total), per_line_inst[loc], loc))


if __name__ == "__main__":
    main()
