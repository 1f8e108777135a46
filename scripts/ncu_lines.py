"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.
usage: python scripts/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; agg = []
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
        agg.append((fname, int(r[0]), r[1].strip()[:90], num("# Samples"), num("Instructions Executed"),
                    num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Ideal")))
ts = sum(a[3] for a in agg); ti = sum(a[4] for a in agg)
print("total samples %d, warp instructions %.3e" % (ts, ti))
print("%-16s %5s %7s %7s %9s %9s  %s" % ("file", "line", "samp%", "inst%", "shWave", "shIdeal", "source"))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-16s %5d %6.2f%% %6.2f%% %9.2e %9.2e  %s" % (a[0], a[1], 100 * a[3] / ts, 100 * a[4] / ti, a[5], a[6], a[2]))
byf = collections.Counter(); byi = collections.Counter()
for a in agg: byf[a[0]] += a[3]; byi[a[0]] += a[4]
print("per file: " + ", ".join("%s %.1f%% samples / %.1f%% inst" % (f, 100 * s / ts, 100 * byi[f] / ti) for f, s in byf.most_common()))
