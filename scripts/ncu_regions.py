"""Aggregate an ncu source-page export per region of a source file.
usage: python scripts/ncu_regions.py report.ncu-rep kernel-regex file.cu name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys, io, collections
rep, kre, fsel = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None
samp = collections.Counter(); inst = collections.Counter(); wav = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "" and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        line = int(r[0]); key = fname
        if fname == fsel:
            key = fsel + ":other"
            for n, lo, hi in regions:
                if lo <= line <= hi: key = n; break
        samp[key] += num("# Samples"); inst[key] += num("Instructions Executed"); wav[key] += num("L1 Wavefronts Shared")
ts, ti, tw = sum(samp.values()), sum(inst.values()), sum(wav.values())
print("total samples %d, warp instructions %.3e, shared wavefronts %.3e" % (ts, ti, tw))
for k, v in samp.most_common():
    print("%-34s samples %5.1f%%  inst %5.1f%%  shared wavefronts %5.1f%%" % (k, 100 * v / ts, 100 * inst[k] / ti, 100 * wav[k] / max(tw, 1)))
