"""Turns the outputs of scripts/gpu_profile_round2.sh (gpurun_out/) into the tracked summaries under profiles/:
    python scripts/make_profile_summary2.py r02h
writes <tag>_bench_full.json, <tag>_bench_reference.json, <tag>_launches.csv + <tag>_launch_summary.txt (kernel shares under ncu next
to the bench's CUDA-event shares), <tag>_ncu_full_summary.txt (key metrics + stall reasons per captured kernel), <tag>_racecheck.txt,
and the two files bench.py reads back: ncu_pipes.json (pipe utilisation per kernel) and ncu_traffic.json (DRAM bytes per unit)."""
import collections, csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


for src, dst in (("bench_full_%s.log", "%s_bench_full.json"), ("bench_ref_%s.log", "%s_bench_reference.json"), ("bench_small_%s.log", "%s_bench_small.json")):
    f = os.path.join(G, src % tag)
    if os.path.exists(f) and last_json(f):
        json.dump(last_json(f), open(os.path.join(P, dst % tag), "w"))
small = last_json(os.path.join(G, "bench_small_%s.log" % tag)) if os.path.exists(os.path.join(G, "bench_small_%s.log" % tag)) else None
NAMES = {"cheaptrick_fast_kernel": "cheaptrick", "cheaptrick_kernel": "cheaptrick", "mcep_tc_kernel": "mcep", "d4c_fast_kernel": "d4c", "d4c_kernel": "d4c_f64_reeval", "render_fast_kernel": "render",
         "overlap_add_kernel": "overlap_add", "mc2sp_tc_kernel": "mc2sp", "mc2sp_kernel": "mc2sp", "decode_ap_kernel": "decode_ap", "lf0_vuv_kernel": "lf0_vuv",
         "bap_from_coarse_kernel": "bap_from_coarse", "stats_kernel": "stats", "phase_inc_kernel": "synth_timebase", "phase_scan_exact_kernel": "synth_timebase",
         "pulse_chunk_kernel": "synth_timebase", "allpass_tc_forward_kernel": "vtln_fwd", "allpass_tc_backward_kernel": "vtln_bwd"}


def short(name):
    for k, v in NAMES.items():
        if k + "<" in name or k + "(" in name or name.endswith(k):
            return v
    return None


# launch list
lf = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lf):
    shutil.copy(lf, os.path.join(P, "%s_launches.csv" % tag))
    rows = list(csv.reader(l for l in open(lf) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ki]) or r[ki].split("(")[0][:50]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "%s_launch_summary.txt" % tag), "w") as o:
        o.write("ncu launch list (gpu__time_duration.sum, --clock-control none): bench.py --utts 256 --steps 1 --warmup 1 --no-workloads (warm-up + timed + e2e passes);\n"
                "per-launch times under ncu are serialised and cold-cache: the SHARES are what is compared with the bench's CUDA-event shares\n")
        o.write("kernel, launches, total ms, share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("%-40s n=%4d %11.3f ms %6.1f%%\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
        if small:
            o.write("\nCUDA-event shares of the same command, not under a profiler (bench_small): %s\n" %
                    json.dumps({k: v["share_of_step"] for k, v in small["kernels"].items()}))
# full captures
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
pipes, traffic, summary = {}, {}, []
for part in ("analysis", "synthesis", "vtln"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, part))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = d["Kernel Name"]
        key = short(name)
        summary.append("== %s  [%s capture]" % (name[:110], part))
        for k in KEYS:
            if k in d:
                summary.append("   %-95s %s %s" % (k, d[k], u[k]))
        st = sorted(((float(d[h] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)
        tots = sum(v for v, _ in st) or 1.0
        summary.append("   stalls: " + " ".join("%s=%.0f%%" % (n, 100 * v / tots) for v, n in st[:8]))
        if key is None or key in pipes:
            continue

        def f(k):
            try:
                return float(d[k].replace(",", ""))
            except Exception:
                return None

        def nbytes(k):
            return f(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u[k].lower()]

        pipes[key] = {"source": "profiles/%s_ncu_full_summary.txt (ncu --set full, one launch)" % tag,
                      "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "l1_lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                      "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                      "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                      "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                      "tensor_pipe_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                      "dram_throughput_pct": f("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                      "local_mem_sectors": (f("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") or 0) + (f("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum") or 0),
                      "registers": f("launch__registers_per_thread"), "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
                      "top_stalls": [n for _, n in st[:3]]}
        rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        units_per_launch = None
        if small and key in small["kernels"]:
            units_per_launch = small["kernels"][key]["units_per_launch"]
            if part == "analysis":  # the capture skips the warm-up pass and takes the FIRST frame chunk of the next one
                units_per_launch = float(min(small["config"]["chunk_frames"], small["config"]["frames_this_rank"]))
        if units_per_launch:
            traffic[key] = {"dram_bytes_per_unit": (rd + wr) / units_per_launch, "dram_read_bytes": rd, "dram_write_bytes": wr,
                            "units_per_launch": units_per_launch, "source": "profiles/%s_ncu_full_summary.txt; units = mean per launch of the same command "
                            "(frames for analysis kernels, pulses for render / overlap-add); writes include L2 write-backs of earlier kernels" % tag}
if summary:
    open(os.path.join(P, "%s_ncu_full_summary.txt" % tag), "w").write(
        "ncu --set full --clock-control none --import-source on; bench.py --utts 256 --steps 1 --warmup 1 (333 k frames: chunks of 262144 + rest; one synthesis batch)\n\n"
        + "\n".join(summary) + "\n")
    json.dump(pipes, open(os.path.join(P, "ncu_pipes.json"), "w"), indent=1)
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(pipes, indent=1)[:3000])
# racecheck
parts = []
for n in ("synthesis", "analysis"):
    f = os.path.join(G, "racecheck_%s_%s.log" % (n, tag))
    if os.path.exists(f):
        t = open(f).read()
        parts.append("== compute-sanitizer --tool racecheck, %s tests\n%s\n" % (n, "\n".join(t.splitlines()[-12:])))
f = os.path.join(G, "memcheck_tc_%s.log" % tag)
if os.path.exists(f):
    parts.append("== compute-sanitizer --tool memcheck, tensor-core kernels (mcep_tc, allpass_tc forward / backward)\n%s\n" % "\n".join(open(f).read().splitlines()[-8:]))
if parts:
    open(os.path.join(P, "%s_racecheck.txt" % tag), "w").write("\n".join(parts))
