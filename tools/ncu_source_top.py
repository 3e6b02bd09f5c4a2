#!/usr/bin/env python
"""Top stall-sample instructions per kernel from an ncu report (source page, SASS view) + opcode mix.

    python tools/ncu_source_top.py gpurun_out/x.ncu-rep [top=25] [kernel-substring]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for ln in out.splitlines():
        if ln.startswith('"Kernel Name"'):
            cur = [ln]
            blocks.append(cur)
        elif cur is not None:
            cur.append(ln)
    seen = set()
    for b in blocks:
        name = next(csv.reader([b[0]]))[1]
        if filt not in name or name in seen:
            continue
        seen.add(name)
        rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
        hdr = rows[0]
        i_src, i_smp, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
        data = rows[1:]
        tot = sum(int(r[i_smp] or 0) for r in data)
        tex = sum(int(r[i_ex] or 0) for r in data)
        print("=" * 100)
        print(name[:120], " samples", tot, " warp-instructions", tex)
        mix = collections.Counter()
        for r in data:
            op = r[i_src].split()[0] if r[i_src].split() else "?"
            if op.startswith("@"):
                op = r[i_src].split()[1]
            mix[op.split(".")[0]] += int(r[i_ex] or 0)
        print(" opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tex, 1)) for k, v in mix.most_common(14)))
        st_tot = collections.Counter()
        for r in data:
            for i, h in stall_cols:
                st_tot[h] += int(r[i] or 0)
        print(" stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in st_tot.most_common(8)))
        order = sorted(range(len(data)), key=lambda k: -int(data[k][i_smp] or 0))[:top]
        for k in sorted(order):
            r = data[k]
            why = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
            print("  %5d  %5.1f%%  [%d] %-60s %s" % (int(r[i_smp] or 0), 100.0 * int(r[i_smp] or 0) / max(tot, 1), k,
                                                  r[i_src].strip()[:60], " ".join("%s:%d" % (h, v) for v, h in why if v)))


if __name__ == "__main__":
    main()
