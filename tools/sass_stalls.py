#!/usr/bin/env python
"""Static look at a kernel's SASS without a GPU: find its loops (backward branches), and for each print the opcode mix
and the sum of the compiler's stall counts (= cycles one warp alone needs per trip, scoreboard waits excluded).

    python tools/sass_stalls.py <cubin-or-so> <mangled-function-substring> [--dump LO HI]

Used for the line-search kernel: FP64 instructions occupy the pipe 2 cycles each, so  2 * (#FP64 ops) / (sum of
stalls)  is the FP64-pipe share one warp can reach alone; profiles/r1_linesearch_schedule.md records the numbers.
"""
import collections
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def load(path, fun):
    names = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = names.split("Function : ")
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if fun in name:
            return name, b
    raise SystemExit("function not found: " + fun)


def parse(text):
    lines = text.split("\n")
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            ins.append((int(m.group(1), 16), m.group(2).strip(), hi))
            i += 2
        else:
            i += 1
    return ins


def opcode(t):
    t = re.sub(r"^@!?U?P\d\s+", "", t)
    return t.split()[0].split(".")[0]


def main():
    path, fun = sys.argv[1], sys.argv[2]
    name, text = load(path, fun)
    ins = parse(text)
    print(name, len(ins), "instructions")
    if "--dump" in sys.argv:
        k = sys.argv.index("--dump")
        lo, hi_ = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        for a, t, h in ins:
            if lo <= a <= hi_:
                print("%05x st=%2d wm=%02x %s" % (a, (h >> 41) & 0xF, (h >> 52) & 0x3F, t))
        return
    skips = []
    if "--skip" in sys.argv:                      # --skip lo-hi,lo-hi : address ranges (hex) left out of the sums
        for r in sys.argv[sys.argv.index("--skip") + 1].split(","):
            lo, hi_ = r.split("-")
            skips.append((int(lo, 16), int(hi_, 16)))
    loops = []
    for a, t, h in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi_ in loops:
        c = collections.Counter()
        st = 0
        n = 0
        for a, t, h in ins:
            if lo <= a <= hi_ and not any(x <= a <= y for x, y in skips):
                c[opcode(t)] += 1
                st += (h >> 41) & 0xF
                n += 1
        f = sum(c[o] for o in FP64)
        if f < 20:
            continue
        fwd = [(a, t) for a, t, h in ins if lo <= a <= hi_ and "BRA" in t and a != hi_]
        print("loop %05x..%05x: %d instr, stall sum %d, FP64 %d (%.0f%% of the lone-warp cycles)  %s" % (
            lo, hi_, n, st, f, 200.0 * f / max(st, 1), dict(c.most_common(12))))
        if "--branches" in sys.argv:
            for a, t in fwd:
                print("      %05x %s" % (a, t))


if __name__ == "__main__":
    main()
