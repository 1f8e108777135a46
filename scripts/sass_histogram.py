"""Per-kernel SASS instruction histogram of idiaptts_b200/libb200world.so (the evidence that the tcgen05 / bulk-copy / packed-f32x2
paths are real and that nothing spills): python scripts/sass_histogram.py [out.txt]
Counts, per kernel (template instantiations merged by base name unless --all): total instructions, UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), SYNCS (mbarrier), UTMALDG (tensor-map TMA), FFMA2 / FMUL2 / FADD2
(packed f32x2), DFMA + DADD + DMUL (fp64), HMMA / IMMA (legacy mma.sync), STL / LDL (local-memory spills), registers."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "idiaptts_b200", "libb200world.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "UTMALDG", "FFMA2", "FMUL2", "FADD2", "FP64", "HMMA", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            base = op.split(".")[0]
            c = per[cur]
            c["total"] += 1
            if base in ("DFMA", "DADD", "DMUL"):
                c["FP64"] += 1
            elif base in ("HMMA", "IMMA"):
                c["HMMA"] += 1
            elif base in KEYS:
                c[base] += 1
    names = demangle(list(per))
    show_all = "--all" in sys.argv
    rows = collections.OrderedDict()
    for mangled, c in per.items():
        full = names.get(mangled, mangled)
        key = full if show_all else re.sub(r"<.*", "", re.sub(r"^void ", "", full)).replace("b2w::(anonymous namespace)::", "b2w::")
        r = rows.setdefault(key, {"n": 0, "c": collections.Counter(), "regs": 0, "stack": 0})
        r["n"] += 1
        r["c"] += c
        rg = regs.get(mangled, (0, 0))
        r["regs"] = max(r["regs"], rg[0])
        r["stack"] = max(r["stack"], rg[1])
    lines = ["SASS instruction histogram of idiaptts_b200/libb200world.so (cuobjdump -sass, sm_100a); template instantiations of a kernel are "
             "summed, registers / stack = maximum over instantiations",
             "%-44s %5s %8s %s  %5s %5s" % ("kernel", "inst.", "total", " ".join("%7s" % k for k in KEYS), "regs", "stack")]
    for k, r in rows.items():
        lines.append("%-44s %5d %8d %s  %5d %5d" % (k[:44], r["n"], r["c"]["total"], " ".join("%7d" % r["c"][x] for x in KEYS), r["regs"], r["stack"]))
    tot = collections.Counter()
    for r in rows.values():
        tot += r["c"]
    lines.append("%-44s %5s %8d %s" % ("ALL", "", tot["total"], " ".join("%7d" % tot[x] for x in KEYS)))
    text = "\n".join(lines) + "\n"
    outs = [a for a in sys.argv[1:] if not a.startswith("--")]
    if outs:
        open(outs[0], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
