"""Turns the outputs of scripts/gpu_profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:
python scripts/make_profile_summary.py r01h <frames per launch of the --set full capture>"""
import collections, csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
frames_full = int(sys.argv[2]) if len(sys.argv) > 2 else 128 * 1301
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# 1. bench lines
for src, dst in (("bench_full_%s.log", "%s_bench_full.json"), ("bench_ref_%s.log", "%s_bench_reference.json"), ("bench_extra_%s.log", "%s_bench_extra.txt")):
    f = os.path.join(G, src % tag)
    if os.path.exists(f):
        lines = [l for l in open(f).read().splitlines() if l.strip()]
        with open(os.path.join(P, dst % tag), "w") as o:
            o.write(("\n".join(lines) if dst.endswith(".txt") else lines[-1]) + "\n")
# 2. launch list
lf = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lf):
    shutil.copy(lf, os.path.join(P, "%s_launches.csv" % tag))
    rows = list(csv.reader(l for l in open(lf) if l.startswith('"')))
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        k = r[ki].split("(")[0][:60]
        agg[k][0] += 1; agg[k][1] += v
    mine = {k: v for k, v in agg.items() if any(s in k for s in ("cheaptrick", "d4c", "mcep", "lf0_vuv", "bap_from", "stats_kernel", "dio_", "stonemask"))}
    tot = sum(v[1] for v in mine.values())
    with open(os.path.join(P, "%s_launch_summary.txt" % tag), "w") as o:
        o.write("ncu launch list (gpu__time_duration.sum, --clock-control none): bench.py --utts 1024 --steps 1 --warmup 1 (warm-up + timed + e2e passes)\n")
        o.write("kernel, launches, total ms, share of my kernels\n")
        for k, v in sorted(mine.items(), key=lambda kv: -kv[1][1]):
            o.write("%-62s n=%4d %11.3f ms %6.1f%%\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
        bf = os.path.join(P, "%s_bench_full.json" % tag)
        if os.path.exists(bf):
            o.write("\nbench.py (full corpus) CUDA-event shares of the same kernels: %s\n" % json.dumps(json.load(open(bf))["roofline"]["shares"]))
# 3. full capture summary + DRAM traffic per frame
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_metrics.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(P, "%s_ncu_full_summary.txt" % tag), "w") as o:
        o.write("ncu --set full --clock-control none --import-source on, bench.py --utts 128 (%d frames per launch), one launch of each main kernel\n\n" % frames_full)
        o.write(out)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, "%s_ncu_full_raw.csv" % tag), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        key = "cheaptrick" if "cheaptrick" in name else "mcep" if "mcep" in name else "d4c" if "d4c" in name else None
        if not key: continue
        def nbytes(k):
            v = float(d[k]); unit = u[k].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        traffic[key] = {"dram_bytes_per_frame": (rd + wr) / frames_full, "dram_read_bytes": rd, "dram_write_bytes": wr, "frames": frames_full,
                        "source": "profiles/%s_ncu_full_raw.csv (writes include L2 write-backs of the preceding kernel's output)" % tag}
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))
