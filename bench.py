#!/usr/bin/env python
"""Benchmark of the WORLD feature hot path (BASELINE.json metric: audio-seconds/s, WORLD analysis + synthesis).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--utts U] [--scaling strong|weak]

Workload (config.workload = "ljspeech_roundtrip"): BASELINE.json configs[1] / configs[2], the LJSpeech-shaped synthetic corpus
(13,100 utterances, durations ~ N(6.5 s, 1.8 s), 22.05 kHz, int16 PCM + cached F0 track).  One step = one pass of the hot path
over the corpus shard of this rank:
    analysis   wav + F0 -> WorldFeatLabelGen features (mcep60, lf0, vuv, bap) + corpus normalisation statistics
               (the path's ONE exchange step: an NCCL all-reduce of the statistics, inside the timed region), then
    synthesis  the features just extracted -> waveforms (Synthesiser.run_world_synth, batches of 256 utterances).
value   : audio-seconds per second with the inputs resident in HBM (CUDA events, max over ranks); `components` gives the
          analysis and the synthesis halves separately.
e2e     : the same pass through host buffers: pinned int16 wave + F0 host->device, features + statistics device->host,
          features host->device, waveforms device->host, all inside the timed region.
scaling : N > 1 defaults to STRONG scaling (configs[2]: the same 13,100-utterance corpus dealt to the ranks by
          distributed.shard_utterances); --scaling weak gives every rank a full-size corpus of its own.
roofline: the kernel with the largest share of the step, timed live with CUDA events on the launching stream; `kernels`
          lists every kernel of the step with the bound that applies to it.
parity  : after the timed region the C oracle re-computes k utterances of the bench corpus on the host; the rows / samples the
          timed region wrote are compared with it (MCD, vuv, bap, lf0, resynthesis SNR) and the run FAILS when out of tolerance.
workloads: BASELINE.json configs[3] (batched synthesis of 256 utterances from acoustic-model-shaped features) and configs[4]
          (Neural-VTLN warp forward + backward, 109 speakers), each with value / e2e / roofline / cpu_baseline.
cpu_baseline / --impl reference: the CPU oracle (C restatement of the WORLD / SPTK algorithms; pyworld / pysptk are not
          installable here) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 22050
UTTS = 13100
DUR = 6.5
DUR_STD = 1.8
DUR_QUANTUM = 0.1   # durations are drawn from N(6.5, 1.8^2) clipped to [1.1, 10.1] s and rounded to 0.1 s
NUM_CODED_SPS = 60
SYNTH_BATCH = 256
METRIC = "audio-seconds/s (WORLD analysis+synthesis)"
TOL = {"mcd_db_max": 0.01, "vuv_mismatch": 0, "bap_abs_max": 1e-3, "lf0_abs_max": 1e-5, "resynthesis_snr_db_min": 60.0}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_pipes():
    """Pipe utilisation of the compute-bound kernels as measured by `ncu --set full` (profiles/ncu_pipes.json, written by
    scripts/make_profile_summary.py from the capture named in it): these cannot be measured from inside the process."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_pipes.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def corpus_args(fixed):
    return dict(mean_dur=DUR, std_dur=0.0 if fixed else DUR_STD, dur_quantum=DUR_QUANTUM)


# ----------------------------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline of the default run, and the whole of --impl reference)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_oracle_roundtrip(waves, f0s, fs, alpha, cores, steps=1, warmup=0):
    """audio-seconds/s of the CPU oracle (oracle/c/world_oracle.c: C restatement of WORLD / SPTK, one utterance per call, the
    GIL is released inside) over the given sample with a pool of `cores` threads: analysis, then synthesis from the features
    just extracted -- the same step the GPU arm times.  Returns (rate, seconds per step, audio seconds, analysis share)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import world_c
    world_c.lib()

    def one(job):
        w, f = job
        t0 = time.perf_counter()
        feats = world_c.extract(w, fs, f, NUM_CODED_SPS, alpha)
        t1 = time.perf_counter()
        y = world_c.synthesize_features(feats, fs, NUM_CODED_SPS, alpha)
        t2 = time.perf_counter()
        return float(y[:8].sum()), len(w) / fs, t1 - t0, t2 - t1

    jobs = [(np.ascontiguousarray(w), f) for w, f in zip(waves, f0s)]
    with ThreadPoolExecutor(max_workers=cores) as pool:
        for _ in range(warmup):
            list(pool.map(one, jobs[:cores]))
        times, audio, ta, ts = [], 0.0, 0.0, 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            res = list(pool.map(one, jobs))
            times.append(time.perf_counter() - t0)
            audio = sum(r[1] for r in res)
            ta, ts = sum(r[2] for r in res), sum(r[3] for r in res)
    sec = sum(times) / len(times)
    return audio / sec, sec, audio, ta / max(ta + ts, 1e-9)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401  (only for the synthetic generator, CPU)
    from idiaptts_b200 import synthetic
    from oracle import sptk_np
    cores = len(os.sched_getaffinity(0))
    alpha = float(sptk_np.mcepalpha(FS))
    n_utts = max(cores, min(2 * cores, 64))  # bounded sample: about one or two ~6.5 s utterances per core and step
    waves, f0s = synthetic.make_corpus(n_utts, FS, seed=2, device="cpu", **corpus_args(args.fixed_dur))
    waves = [w.numpy() for w in waves]
    value, sec, audio, share_a = cpu_oracle_roundtrip(waves, f0s, FS, alpha, cores, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ljspeech_roundtrip", "fs": FS, "utt_seconds_mean": DUR, "num_coded_sps": NUM_CODED_SPS,
                       "sample_utts": n_utts, "sample_audio_s": audio, "analysis_share_of_cpu_time": round(share_a, 3),
                       "note": "CPU oracle = C restatement of WORLD/SPTK (analysis + synthesis), gcc -O3 -march=native, one thread per "
                               "utterance over all host cores (pyworld/pysptk unavailable offline)"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "%d utterances (%.0f audio-s) per step, analysis + synthesis" % (n_utts, audio)},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def kernel_table(events, steps):
    """events: (name, units, start, end) -> {name: {ms_per_step, launches_per_step, units_per_launch, avg_launch_ms}}"""
    per = {}
    for name, units, a, b in events:
        d = per.setdefault(name, [0.0, 0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
        d[2] += units() if callable(units) else units
    return {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps, "units_per_launch": v[2] / max(v[1], 1),
                "avg_launch_ms": v[0] / max(v[1], 1)} for k, v in per.items()}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from idiaptts_b200 import _lib, distributed, ops, pipeline, synthetic
    from idiaptts_b200.compat.pysptk import mcepalpha

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus_bound = distributed.bind_host_to_gpu(local)  # before any pinned allocation: host buffers on the GPU's own socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    alpha = float(mcepalpha(FS))
    scaling = args.scaling or "strong"
    cargs = corpus_args(args.fixed_dur)
    t0 = time.time()
    if scaling == "strong" and world > 1:
        # configs[2]: ONE corpus of args.utts utterances, dealt longest-first to the ranks (every rank knows all durations)
        durs = synthetic.corpus_durations(range(args.utts), 2, **cargs)
        frames_all = np.round(durs * 200.0).astype(np.int64) + 1
        ids = distributed.shard_utterances(frames_all, world)[rank]
        imbalance = float(max(frames_all[s].sum() for s in distributed.shard_utterances(frames_all, world)) * world / frames_all.sum())
    else:
        ids = np.arange(args.utts, dtype=np.int64) + (rank * args.utts if world > 1 else 0)
        imbalance = 1.0
    waves, f0s = synthetic.make_corpus(0, FS, seed=2, device=dev, batch=64, utt_ids=ids, **cargs)
    utts = len(waves)
    lens = np.array([w.numel() for w in waves], np.int64)
    flens = np.array([len(f) for f in f0s], np.int64)
    audio_s = float(lens.sum()) / FS
    x_dev = torch.cat(waves)
    keep_waves = [waves[u].cpu().numpy() for u in range(min(args.parity_utts, utts))] if rank == 0 else []
    del waves
    f0_np = np.concatenate(f0s)
    t_np = np.concatenate([np.arange(n) * 5.0 / 1000.0 for n in flens])
    sample_off = np.concatenate(([0], np.cumsum(lens)))
    frame_off = np.concatenate(([0], np.cumsum(flens)))
    frame_utt = np.repeat(np.arange(utts, dtype=np.int32), flens)
    gen_s = time.time() - t0
    host = {"x": x_dev.cpu().pin_memory(), "f0": torch.from_numpy(f0_np).pin_memory(), "t": torch.from_numpy(t_np).pin_memory(),
            "so": torch.from_numpy(sample_off).pin_memory(), "fo": torch.from_numpy(frame_off).pin_memory(),
            "fu": torch.from_numpy(frame_utt).pin_memory()}
    batch = ops.RaggedBatch(x_dev, host["so"].to(dev), host["f0"].to(dev), host["t"].to(dev), host["fo"].to(dev), host["fu"].to(dev), FS)
    F = batch.num_frames
    an = pipeline.WorldAnalyzer(FS, NUM_CODED_SPS, alpha, device=dev, chunk_frames=args.chunk_frames)
    syn = pipeline.WorldSynthesizer(FS, NUM_CODED_SPS, alpha, device=dev)
    an.iters = torch.zeros(F, dtype=torch.int32, device=dev)
    feats = torch.empty((F, an.dim), dtype=torch.float32, device=dev)
    stat_buf = torch.zeros(2 * an.dim + 1, dtype=torch.float64, device=dev)
    ylen = (flens * 5.0 * FS / 1000).astype(np.int64)
    y_off = np.concatenate(([0], np.cumsum(ylen)))
    y_all = torch.empty(int(y_off[-1]), dtype=torch.float32, device=dev)
    total_audio = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_audio)
    total_audio = float(total_audio.item())

    def step(events=None, marks=None):
        stat_buf.zero_()
        _, _, status = an.extract(batch, feats=feats, sums=stat_buf[:2 * an.dim], events=events)
        stat_buf[2 * an.dim] = float(F)
        if world > 1:
            dist.all_reduce(stat_buf)  # the path's only exchange step: corpus normalisation statistics
        if marks is not None:
            m = torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
        _, st2 = syn.synthesize_corpus(feats, frame_off, batch_utts=args.synth_batch, out=y_all, events=events)
        return status, st2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        status, st2 = step()
    barrier()
    ops.raise_for_status(status, "extract")
    ops.raise_for_status(st2, "synthesize")

    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    events, marks, starts = [], [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        s = torch.cuda.Event(enable_timing=True)
        s.record()
        starts.append(s)
        step(events, marks)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = total_audio / (ms_step / 1e3)
    ms_analysis = sum(s.elapsed_time(m) for s, m in zip(starts, marks)) / args.steps
    ms_synth = ms / args.steps - ms_analysis
    stats_host = stat_buf.cpu().numpy()
    peaks, peak_src = measured_peaks()

    # ---- per-kernel table (rank 0's stream) ---------------------------------------------------------------------------
    kt = kernel_table(events, args.steps)
    total_k = sum(v["ms_per_step"] for v in kt.values())
    H = FS * 0.005
    K = an.n_fft // 2 + 1
    N = an.n_fft
    mean_it = float(an.iters.float().mean().item())
    m_ord = NUM_CODED_SPS - 1
    # algorithmic bytes per unit (frames for the analysis and plane kernels, pulses for render / overlap-add, samples for the
    # time base): what the kernel must read and write given the layouts of DESIGN.md 3
    alg_bytes = {
        "lf0_vuv": 8 + 8, "cheaptrick": 2 * H + 20 + 4 * K, "mcep": 4 * K + 4 * NUM_CODED_SPS, "d4c": 2 * H + 20 + 8 * an.nap + 1,
        "bap_from_coarse": 8 * an.nap + 1 + 4 * an.nap, "stats": 4 * an.dim,
        "mc2sp": 4 * NUM_CODED_SPS + 4 * K, "decode_ap": 8 * an.nap + 4 * K, "synth_timebase": 3 * 8,   # float32 planes (fast path)
        "render": 2 * 2 * 4 * K + 4 * N, "overlap_add": 4 * N + 4 * H_pulse(FS),                      # float32 responses
    }
    bound = {"lf0_vuv": "latency", "cheaptrick": "l1_shared_pipe", "mcep": "tensor", "d4c": "issue+l1_shared_pipe",
             "bap_from_coarse": "latency", "stats": "hbm", "mc2sp": "tensor+hbm (tcgen05 3xTF32, hand-overs per 32-bin chunk)", "decode_ap": "hbm", "synth_timebase": "latency",
             "render": "l1_shared_pipe+latency", "overlap_add": "hbm"}
    pipes = ncu_pipes()
    kernels = {}
    for k, v in kt.items():
        gbs = alg_bytes.get(k, 0) * v["units_per_launch"] / (v["avg_launch_ms"] / 1e3) / 1e9 if v["avg_launch_ms"] > 0 else 0.0
        kernels[k] = {"share_of_step": round(v["ms_per_step"] / total_k, 4), "avg_launch_ms": round(v["avg_launch_ms"], 4),
                      "launches_per_step": v["launches_per_step"], "units_per_launch": round(v["units_per_launch"], 1),
                      "bound": bound.get(k, "?"), "algorithmic_GBps": round(gbs, 2), "hbm_frac": round(gbs / peaks["hbm_gbs"], 5)}
        if k in pipes:
            kernels[k]["ncu"] = pipes[k]
    # tensor-core kernel: algorithmic FLOPs of the warp contractions (initial value + mean Newton passes x (freqt+FFT and
    # IFFT+frqtr as GEMMs)); the 3xTF32 split issues three TF32 MMAs per algorithmic one
    flops_frame = 2.0 * K * (m_ord + 2) + mean_it * 2.0 * K * ((m_ord + 1) + (2 * m_ord + 1))
    mcep_tflops = flops_frame * kt["mcep"]["units_per_launch"] / (kt["mcep"]["avg_launch_ms"] / 1e3) / 1e12
    tf32_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0
    kernels["mcep"].update({"tensor_tflops_algorithmic": round(mcep_tflops, 2), "tensor_tflops_issued_tf32": round(3 * mcep_tflops, 2),
                            "tensor_frac_of_tf32_peak": round(3 * mcep_tflops / tf32_peak, 4), "mean_newton_passes": round(mean_it, 3)})
    top = max(kt, key=lambda k: kt[k]["ms_per_step"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)[top]
            traffic = float(tj.get("dram_bytes_per_unit", tj.get("dram_bytes_per_frame"))) * kt[top]["units_per_launch"]
    except Exception:
        pass
    if top == "mcep":
        roofline = {"kernel": top, "bound": "tensor", "achieved": 3 * mcep_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": 3 * mcep_tflops / tf32_peak, "traffic": traffic, "peak_source": peak_src,
                    "note": "TF32 MMAs issued (3xTF32 split = 3 per algorithmic MMA) against the dense TF32 peak = half the measured bf16 "
                            "rate; the kernel is bound by its per-frame 60x60 solves and the exp epilogue (kernels.mcep.ncu), not by the "
                            "tensor pipe"}
    else:
        a = kernels[top]["algorithmic_GBps"]
        roofline = {"kernel": top, "bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": peak_src, "binding": bound.get(top),
                    "binding_frac": (max(pipes[top].get("issue_active_pct") or 0.0, pipes[top].get("l1_lsu_wavefronts_pct") or 0.0) / 100.0
                                     if top in pipes else None),
                    "note": "the HBM fraction is what the metric asks for; this kernel is bound on chip: `binding` names the pipes, "
                            "`binding_frac` is the higher of issue-slot and L1 / shared-memory-pipe utilisation in the ncu capture under "
                            "profiles/ (kernels.%s.ncu)" % top}
    roofline["share_of_step"] = kernels[top]["share_of_step"]

    # ---- timed region 2: end to end through host buffers ---------------------------------------------------------------
    feats_host = torch.empty((F, an.dim), dtype=torch.float32).pin_memory()
    y_host = torch.empty(int(y_off[-1]), dtype=torch.float32).pin_memory()
    stats_pinned = torch.empty(2 * an.dim + 1, dtype=torch.float64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + feats_host.numel() * 4
    d2h = feats_host.numel() * 4 + stats_pinned.numel() * 8 + y_host.numel() * 4
    e2e_state = {"buf": None}

    e2e_marks = []

    def e2e_step():
        # the public end-to-end calls: pinned host corpus in, pinned host features + statistics out (WorldAnalyzer.extract_from_host),
        # then pinned host features in, pinned host waveforms out (WorldSynthesizer.synthesize_corpus); copies overlap the kernels
        m0, m1, m2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        m0.record()
        stat_buf.zero_()
        _, _, _, e2e_state["buf"] = an.extract_from_host(host, feats_host, dev_buffers=e2e_state["buf"], sums=stat_buf[:2 * an.dim])
        stat_buf[2 * an.dim] = float(F)
        if world > 1:
            dist.all_reduce(stat_buf)
        stats_pinned.copy_(stat_buf, non_blocking=True)
        m1.record()
        torch.cuda.current_stream().synchronize()   # the features must have landed in host memory before they are read back
        syn.synthesize_corpus(None, frame_off, batch_utts=args.synth_batch, feats_host=feats_host, out_host=y_host)
        m2.record()
        e2e_marks.append((m0, m1, m2))

    e2e_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        e2e_step()
    g1.record()
    barrier()
    t2 = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = total_audio / (float(t2.item()) / args.steps / 1e3)
    e2e_parts = {"analysis_ms": round(sum(a.elapsed_time(b) for a, b, _ in e2e_marks[1:]) / args.steps, 2),
                 "synthesis_ms": round(sum(b.elapsed_time(c) for _, b, c in e2e_marks[1:]) / args.steps, 2),
                 "what": "rank 0, host copies included; the resident-input passes take components.*.ms_per_step"}
    del e2e_state

    line = None
    if rank == 0:
        # ---- parity of what the timed region wrote, against the CPU oracle (bounded: args.parity_utts utterances) -------------
        parity = check_parity(keep_waves, f0s, feats, frame_off, y_all, y_off, alpha, an.nap) if keep_waves else None
        # ---- configs[3] / configs[4] as workloads of their own -----------------------------------------------------------------
        workloads = None
        if not args.no_workloads:
            workloads = {"synth256": bench_synth256(args, dev, syn, feats, frame_off, stats_host, an, peaks, alpha),
                         "vtln109": bench_vtln109(args, dev, peaks)}
            try:
                workloads["postprocess256"] = bench_postprocess256(args, dev, feats, frame_off, stats_host, an, peaks)
            except Exception as e:  # noqa: BLE001
                workloads["postprocess256"] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
            try:  # host file systems differ from box to box: a failure here must not take the headline line with it
                workloads["gen_data_files"] = bench_gen_data_files(args, host, sample_off, f0s, alpha)
            except Exception as e:  # noqa: BLE001
                workloads["gen_data_files"] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # separate process: no fork of a CUDA-initialised interpreter
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"]
            r = subprocess.run(cmd + (["--fixed-dur"] if args.fixed_dur else []), capture_output=True, text=True)
            try:
                cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception:
                cpu = {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "port", "sample": "failed: " + r.stderr[-300:]}
        n = float(stats_host[2 * an.dim])
        mean, std = pipeline.mean_std_from_sums(stats_host, n, an.dim)
        launches = an.kernel_launches(F) + syn.kernel_launches(utts, args.synth_batch)
        line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f32 FFTs + f64 scans (analysis), 3xTF32 (mel-cepstrum), f64 (synthesis)", "data": "synthetic",
                "config": {"workload": "ljspeech_roundtrip", "utts_total": int(args.utts * (world if scaling == "weak" and world > 1 else 1)),
                           "utts_this_rank": utts, "fs": FS, "utt_seconds": "N(%.1f, %.1f^2) clipped [1.1, 10.1]" % (DUR, DUR_STD)
                           if not args.fixed_dur else DUR, "audio_seconds_total": total_audio, "frames_this_rank": int(F),
                           "num_coded_sps": NUM_CODED_SPS, "mgc_alpha": alpha, "fft_size": an.n_fft, "num_bap": an.nap,
                           "chunk_frames": an.chunk_frames, "synth_batch_utts": args.synth_batch, "shard_imbalance": round(imbalance, 4), "host_cpus_bound_to_gpu_socket": None if cpus_bound is None else len(cpus_bound),
                           "l2": "inputs (%.2f GB int16), features (%.2f GB) and waveforms (%.2f GB) exceed the 126 MB L2" % (
                               host["x"].numel() * 2 / 1e9, feats.numel() * 4 / 1e9, y_all.numel() * 4 / 1e9),
                           "corpus_gen_s": round(gen_s, 1), "mean_c0": float(mean[0]), "std_c0": float(std[0])},
                "components": {"analysis": {"audio_s_per_s": audio_s / (ms_analysis / 1e3), "ms_per_step": ms_analysis,
                                            "what": "wav + cached F0 -> mcep60/lf0/vuv/bap + statistics all-reduce (rank 0)"},
                               "synthesis": {"audio_s_per_s": audio_s / (ms_synth / 1e3), "ms_per_step": ms_synth,
                                             "what": "features -> waveforms, batches of %d utterances (rank 0)" % args.synth_batch}},
                "clocks": clocks, "e2e_components": e2e_parts, "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d),
                                          "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(args.steps * launches), "roofline": roofline, "kernels": kernels, "parity": parity,
                "cpu_baseline": cpu, "workloads": workloads}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if line is not None and line["parity"] is not None and not line["parity"]["ok"]:
        raise SystemExit("parity check failed: %s" % json.dumps(line["parity"]))


def H_pulse(fs):
    """mean output samples per pulse of the synthetic corpus (~ fs / mean pulse rate, ~ 310 pulses per audio-second)"""
    return fs / 310.0


def check_parity(waves, f0s, feats, frame_off, y_all, y_off, alpha, nap):
    """C oracle on the first len(waves) utterances of rank 0's shard vs the feature rows / waveforms of the timed region."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import glue_np, world_c
    world_c.lib()
    D = NUM_CODED_SPS
    k = len(waves)
    cores = len(os.sched_getaffinity(0))
    gpu_rows = [feats[int(frame_off[u]):int(frame_off[u + 1])].cpu().numpy() for u in range(k)]
    gpu_y = [y_all[int(y_off[u]):int(y_off[u + 1])].cpu().numpy().astype(np.float64) for u in range(k)]

    def one(u):
        ref = world_c.extract(waves[u], FS, f0s[u], D, alpha)
        y_ref = world_c.synthesize_features(gpu_rows[u], FS, D, alpha).astype(np.float64)   # from the GPU's own features
        return ref, y_ref

    with ThreadPoolExecutor(max_workers=min(cores, k)) as pool:
        res = list(pool.map(one, range(k)))
    out = {"utterances": k, "mcd_db_max": 0.0, "vuv_mismatch": 0, "bap_abs_max": 0.0, "lf0_abs_max": 0.0, "resynthesis_snr_db_min": 1e9,
           "frames": int(sum(len(r) for r in gpu_rows))}
    for u, (ref, y_ref) in enumerate(res):
        g = gpu_rows[u]
        out["mcd_db_max"] = max(out["mcd_db_max"], float(glue_np.mcd_db(ref[:, :D], g[:, :D])))
        out["vuv_mismatch"] += int((ref[:, D + 1] != g[:, D + 1]).sum())
        out["bap_abs_max"] = max(out["bap_abs_max"], float(np.abs(ref[:, D + 2:] - g[:, D + 2:]).max()))
        out["lf0_abs_max"] = max(out["lf0_abs_max"], float(np.abs(ref[:, D] - g[:, D]).max()))
        err = ((gpu_y[u] - y_ref) ** 2).sum()
        out["resynthesis_snr_db_min"] = min(out["resynthesis_snr_db_min"], float(10 * np.log10((y_ref ** 2).sum() / max(err, 1e-300))))
    out["tolerances"] = TOL
    out["ok"] = bool(out["mcd_db_max"] < TOL["mcd_db_max"] and out["vuv_mismatch"] == 0 and out["bap_abs_max"] < TOL["bap_abs_max"]
                     and out["lf0_abs_max"] < TOL["lf0_abs_max"] and out["resynthesis_snr_db_min"] > TOL["resynthesis_snr_db_min"])
    out["oracle"] = "oracle/c/world_oracle.c (analysis half pinned to the reference's fixtures; synthesis half = restatement, unpinned)"
    return out


class L2Flusher:
    """Between timed iterations of the small workloads: write a buffer larger than the 126 MB L2."""

    def __init__(self, dev):
        import torch
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.zero_()


def timed_steps(fn, steps, warmup, flush):
    """W warm-up calls, then K calls timed one by one with CUDA events (an L2 flush between them, outside the timed spans)."""
    import torch
    for _ in range(max(warmup, 1)):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / steps


def bench_synth256(args, dev, syn, feats, frame_off, stats_host, an, peaks, alpha):
    """BASELINE.json configs[3]: batched WORLD synthesis of 256 utterances from acoustic-model-shaped outputs = the features of the
    first 256 corpus utterances + N(0, (0.05 sigma_d)^2) per dimension, vuv = clip(vuv + N(0, 0.1^2), 0, 1), seed 4."""
    import torch
    from idiaptts_b200 import pipeline
    D = NUM_CODED_SPS
    nu = min(256, len(frame_off) - 1)
    f_end = int(frame_off[nu])
    n = float(stats_host[2 * an.dim])
    _, std = pipeline.mean_std_from_sums(stats_host, n, an.dim)
    gen = torch.Generator(device=dev).manual_seed(4)
    sub = feats[:f_end].clone()
    sub += torch.randn(sub.shape, generator=gen, device=dev) * torch.from_numpy((0.05 * std).astype(np.float32)).to(dev)
    sub[:, D + 1] = torch.clamp(feats[:f_end, D + 1] + 0.1 * torch.randn(f_end, generator=gen, device=dev), 0.0, 1.0)
    fo = np.asarray(frame_off[:nu + 1], np.int64)
    audio = f_end * 0.005
    flush = L2Flusher(dev)
    events = []
    ms = timed_steps(lambda: syn.synthesize_corpus(sub, fo, batch_utts=256), args.steps, args.warmup, flush)
    syn.synthesize_corpus(sub, fo, batch_utts=256, events=events)
    torch.cuda.synchronize()
    kt = kernel_table(events, 1)
    sub_host = sub.cpu().pin_memory()
    ylen = (np.diff(fo) * 5.0 * FS / 1000).astype(np.int64)
    y_host = torch.empty(int(ylen.sum()), dtype=torch.float32).pin_memory()
    ms_e2e = timed_steps(lambda: syn.synthesize_corpus(None, fo, batch_utts=256, feats_host=sub_host, out_host=y_host),
                         args.steps, args.warmup, flush)
    K, N = an.n_fft // 2 + 1, an.n_fft
    oa = kt.get("overlap_add")
    pulses = oa["units_per_launch"] if oa else 0
    oa_gbs = (pulses * 4 * N + int(ylen.sum()) * 4) / (oa["avg_launch_ms"] / 1e3) / 1e9 if oa else 0.0   # float32 responses
    top = max(kt, key=lambda k: kt[k]["ms_per_step"])
    # CPU oracle on a bounded sample of the same rows
    from concurrent.futures import ThreadPoolExecutor
    from oracle import world_c
    world_c.lib()
    cores = len(os.sched_getaffinity(0))
    ns = min(nu, max(cores, 16))
    rows = [sub_host[int(fo[u]):int(fo[u + 1])].numpy() for u in range(ns)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        list(pool.map(lambda r: world_c.synthesize_features(r, FS, D, alpha)[:1], rows))
    cpu_rate = float(fo[ns]) * 0.005 / (time.perf_counter() - t0)
    return {"workload": "batched WORLD synthesis of %d utterances from acoustic-model-shaped features (mgc60+lf0+vuv+bap, seed 4)" % nu,
            "metric": "audio-seconds/s (WORLD synthesis)", "value": audio / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms,
            "steps": args.steps, "warmup": args.warmup, "l2": "flushed between timed iterations (256 MB write)",
            "e2e": {"value": audio / (ms_e2e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": int(sub_host.numel() * 4),
                    "d2h_bytes_per_step": int(y_host.numel() * 4)},
            "kernels": {k: {"ms": round(v["ms_per_step"], 4), "share": round(v["ms_per_step"] / sum(x["ms_per_step"] for x in kt.values()), 4)}
                        for k, v in kt.items()},
            "roofline": {"kernel": "overlap_add", "bound": "hbm", "achieved": oa_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": oa_gbs / peaks["hbm_gbs"], "traffic": None,
                         "note": "the HBM-bound kernel of the workload (one 4 KB float32 response read per pulse, float32 samples written); the "
                                 "dominant kernel is `%s` (on-chip bound, see kernels.render.ncu of the main line)" % top},
            "cpu_baseline": {"value": cpu_rate, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "%d utterances of the same rows" % ns}}


def bench_postprocess256(args, dev, feats, frame_off, stats_host, an, peaks):
    """The steps between the acoustic model and the vocoder on the features of the first 256 corpus
    utterances: deltas (utils.py:103-105), the padded + normalised trainer batch and its inverse (ModularModelHandlerPyTorch
    .prepare_batch, :389-499), MLPG of the 60 mel-cepstral trajectories (misc/mlpg.py:94-127) and the objective metrics
    (Metrics.py:84-164); one CUDA-event time per operator, L2 flushed between iterations."""
    import torch
    from idiaptts_b200 import ops, pipeline
    D = NUM_CODED_SPS
    nu = min(256, len(frame_off) - 1)
    f_end = int(frame_off[nu])
    fo_np = np.asarray(frame_off[:nu + 1], np.int64)
    fo = torch.from_numpy(fo_np).to(dev)
    fu = torch.from_numpy(np.repeat(np.arange(nu, dtype=np.int32), np.diff(fo_np))).to(dev)
    sub = feats[:f_end].contiguous()
    W = sub.shape[1]
    n = float(stats_host[2 * an.dim])
    mean, std = pipeline.mean_std_from_sums(stats_host, n, an.dim)
    m_dev = torch.from_numpy(mean.astype(np.float32)).to(dev)
    s_dev = torch.from_numpy(np.maximum(std, 1e-6).astype(np.float32)).to(dev)
    flush = L2Flusher(dev)
    lengths = np.diff(fo_np)
    out = {}

    def timed(name, fn, bytes_moved=None):
        ms = timed_steps(fn, args.steps, args.warmup, flush)
        out[name] = {"ms": round(ms, 4), "frames_per_s": f_end / (ms / 1e3)}
        if bytes_moved:
            gbs = bytes_moved / (ms / 1e3) / 1e9
            out[name].update({"algorithmic_GBps": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4)})

    d, dd = ops.deltas(sub, fo)
    timed("deltas", lambda: ops.deltas(sub, fo), 3 * f_end * W * 4)
    padded, mask, _ = ops.pad_normalise(sub, fo, m_dev, s_dev, lengths=lengths)
    t_max = int(lengths.max())
    timed("pad_normalise", lambda: ops.pad_normalise(sub, fo, m_dev, s_dev, lengths=lengths),
          f_end * W * 4 + nu * t_max * (W + 1) * 4)
    timed("unpad_denormalise", lambda: ops.unpad_denormalise(padded, fo, fu, m_dev, s_dev), 2 * f_end * W * 4)
    tri = torch.cat((sub[:, :D], d[:, :D], dd[:, :D]), dim=1).contiguous()
    var3 = torch.ones(3 * D, dtype=torch.float64, device=dev)
    timed("mlpg_60_trajectories", lambda: ops.mlpg(tri, var3, fo, D))
    gen = torch.Generator(device=dev).manual_seed(6)
    noisy = sub + 0.05 * torch.randn(sub.shape, generator=gen, device=dev) * s_dev
    noisy[:, D + 1] = sub[:, D + 1]
    timed("world_metrics", lambda: ops.world_metrics(sub, noisy, fu, nu, D, an.nap), 2 * f_end * W * 4)
    return {"workload": "post-network steps on %d utterances (%d frames x %d features)" % (nu, f_end, W), "operators": out,
            "l2": "flushed between timed iterations (256 MB write)", "steps": args.steps, "warmup": args.warmup}


def bench_gen_data_files(args, host, sample_off, f0s, alpha):
    """WorldFeatLabelGen.gen_data FILES TO FILES -- wav files + cached F0 in, per-feature .npz archives and the
    normalisation files out (WorldFeatLabelGen.py:947-1071) -- on the first --io-utts corpus utterances, wall clock (host IO is
    the subject).  Beside it the reference's per-utterance IO loop alone (wave-module read + four numpy.savez per utterance, no
    feature extraction at all) on the same files and arrays."""
    import shutil
    import tempfile
    import wave
    import torch
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    nu = min(args.io_utts, len(f0s))
    if nu <= 0:
        return None
    need = 6 * int(sample_off[nu]) * 2  # wav + up to two feature sets + the numpy copies, with margin
    base = "/dev/shm" if (os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK)
                          and shutil.disk_usage("/dev/shm").free > need) else None
    if base is None and shutil.disk_usage(tempfile.gettempdir()).free < need:
        return {"skipped": "no file system with %.1f GB free for the files-to-files workload" % (need / 1e9)}
    root = tempfile.mkdtemp(prefix="b2w_io_", dir=base)
    try:
        wav_dir, ids, cache = os.path.join(root, "wav"), [], {}
        os.makedirs(wav_dir)
        x = host["x"].numpy()
        for u in range(nu):
            name = "utt%05d" % u
            with wave.open(os.path.join(wav_dir, name + ".wav"), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(FS)
                w.writeframes(x[sample_off[u]:sample_off[u + 1]].tobytes())
            ids.append(name)
            cache[name] = f0s[u]
        audio = float(sample_off[nu]) / FS
        # (only rank 0 runs the workloads: the generator must neither shard the list nor enter the statistics all-reduce)
        gen = WorldFeatLabelGen(os.path.join(root, "out2"), num_coded_sps=NUM_CODED_SPS, num_bap=2, f0_cache=cache, mgc_alpha=alpha,
                                use_distributed=False)
        times = []
        for it in range(3):  # the first pass warms the page cache and the allocator
            out = os.path.join(root, "out%d" % it)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gen.gen_data(wav_dir, out, file_id_list="train.txt", id_list=ids)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            if it < 2:
                shutil.rmtree(out)
        t_native = min(times[1:])
        # the batched reader protocol: archives -> one pinned packed matrix
        t_loads = []
        for it in range(2):  # the first call also page-locks a new host block; later calls reuse it (torch's caching host allocator)
            got = None
            t0 = time.perf_counter()
            got, goff = gen.load_batch(ids, pin=True)
            t_loads.append(time.perf_counter() - t0)
        t_load = t_loads[1]
        # the reference's IO loop alone, on the same data (what would remain if the extraction cost nothing)
        feats_np = got.numpy()
        cols = (("mcep", 0, NUM_CODED_SPS), ("lf0", NUM_CODED_SPS, 1), ("vuv", NUM_CODED_SPS + 1, 1), ("bap", NUM_CODED_SPS + 2, 2))
        py_out = os.path.join(root, "py")
        for k, _, _ in cols:
            os.makedirs(os.path.join(py_out, k))
        t0 = time.perf_counter()
        for u, name in enumerate(ids):
            with wave.open(os.path.join(wav_dir, name + ".wav"), "rb") as w:
                np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
            rows = feats_np[goff[u]:goff[u + 1]]
            for k, c0, c in cols:
                np.savez(os.path.join(py_out, k, name), **{k: rows[:, c0:c0 + c]})
        t_py_write = time.perf_counter() - t0
        t0 = time.perf_counter()
        for name in ids:
            parts = []
            for k, _, _ in cols:
                with np.load(os.path.join(py_out, k, name + ".npz")) as arc:
                    parts.append(arc[k])
            np.concatenate(parts, axis=1)
        t_py_load = time.perf_counter() - t0
        return {"metric": "gen_data wav files + cached F0 -> npz feature files + normalisation files, wall clock",
                "utts": nu, "audio_seconds": audio, "value": audio / t_native, "unit": "audio-s/s", "seconds": round(t_native, 3),
                "first_call_seconds": round(times[0], 3), "filesystem": "tmpfs (/dev/shm)" if base else "tempfile default",
                "bytes_read": int(sample_off[nu]) * 2, "bytes_written": int(feats_np.size) * 4,
                "load_batch_seconds": round(t_load, 3), "load_batch_first_call_seconds": round(t_loads[0], 3), "load_batch_frames_per_s": float(goff[-1]) / t_load,
                "python_io_loop_only": {"read_wav_savez_seconds": round(t_py_write, 3), "np_load_seconds": round(t_py_load, 3),
                                        "what": "the reference's per-utterance wave read + 4 numpy.savez / 4 numpy.load, no extraction"},
                "host_threads": os.cpu_count()}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def bench_vtln109(args, dev, peaks):
    """BASELINE.json configs[4]: Neural-VTLN all-pass warp forward + backward on VCTK-shaped mgc batches: 109 speakers x 4
    utterances x 1301 frames x 60 coefficients, one alpha ~ U[-0.2, 0.2] per speaker, grad_out ~ N(0, 1), seed 5."""
    import torch
    from idiaptts_b200 import ops
    n_v, spk, utt_v, T_v = 60, 109, 4, 1301
    rows = spk * utt_v * T_v
    gen = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((rows, n_v), generator=gen, device=dev)
    g = torch.randn((rows, n_v), generator=gen, device=dev)
    a = (torch.rand(spk, generator=gen, device=dev) * 0.4 - 0.2).repeat_interleave(utt_v * T_v).contiguous()
    flush = L2Flusher(dev)
    ms_f = timed_steps(lambda: ops.allpass_forward(x, a, n_v), args.steps, args.warmup, flush)
    ms_b = timed_steps(lambda: ops.allpass_backward(g, x, a, n_v), args.steps, args.warmup, flush)
    # end to end: pinned host x / grad_out / alpha in, y / grad_x / grad_alpha back
    xh, gh, ah = x.cpu().pin_memory(), g.cpu().pin_memory(), a.cpu().pin_memory()
    yh, gxh, gah = torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory(), torch.empty_like(ah).pin_memory()

    def e2e():
        xd, gd, ad = xh.to(dev, non_blocking=True), gh.to(dev, non_blocking=True), ah.to(dev, non_blocking=True)
        y = ops.allpass_forward(xd, ad, n_v)
        gx, ga = ops.allpass_backward(gd, xd, ad, n_v)
        yh.copy_(y, non_blocking=True)
        gxh.copy_(gx, non_blocking=True)
        gah.copy_(ga, non_blocking=True)

    ms_e = timed_steps(e2e, args.steps, args.warmup, flush)
    bytes_f, bytes_b = rows * (8 * n_v + 4), rows * (12 * n_v + 8)
    gbs = (bytes_f + bytes_b) / ((ms_f + ms_b) / 1e3) / 1e9
    # CPU: the reference layer's own formulation (AllPassWarp.py:148-205: per-frame warp matrix W = einsum(w3, alpha^k), then bmm)
    # restated with CPU torch on a bounded sample of rows, forward + autograd backward, all host cores (values do not matter for
    # the timing: at n = 60 the reference's float32 polynomial tensor overflows anyway)
    ns = 8192
    w3 = torch.randn(n_v, n_v, 2 * n_v) * 1e-3
    xs = xh[:ns].clone().requires_grad_(True)
    as_ = ah[:ns].clone().requires_grad_(True)
    t0 = time.perf_counter()
    pw = as_[:, None] ** torch.arange(2 * n_v, dtype=torch.float32)[None, :]
    Wm = torch.einsum("bk,rck->brc", pw, w3)
    ys = torch.bmm(xs[:, None, :], Wm)[:, 0, :]
    ys.backward(gh[:ns])
    cpu_rows = ns / (time.perf_counter() - t0)
    cores = torch.get_num_threads()
    return {"workload": "Neural-VTLN all-pass warp fwd + bwd, %d speakers x %d utts x %d frames, n = %d, one alpha per speaker (seed 5)" % (
                spk, utt_v, T_v, n_v),
            "metric": "frames/s (VTLN warp forward + backward)", "value": rows / ((ms_f + ms_b) / 1e3), "unit": "frames/s",
            "fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4), "steps": args.steps, "warmup": args.warmup,
            "l2": "flushed between timed iterations (256 MB write)",
            "e2e": {"value": rows / (ms_e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(rows * (8 * n_v + 4)),
                    "d2h_bytes_per_step": int(rows * (8 * n_v + 4))},
            "roofline": {"kernel": "allpass_tc_forward + allpass_tc_backward", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                         "fwd_frac": bytes_f / (ms_f / 1e3) / 1e9 / peaks["hbm_gbs"], "bwd_frac": bytes_b / (ms_b / 1e3) / 1e9 / peaks["hbm_gbs"]},
            "cpu_baseline": {"value": cpu_rows, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d rows, CPU torch restatement of AllPassWarp.forward (einsum + bmm) + autograd backward" % ns}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=UTTS, help="utterances of the corpus (strong scaling: in total; weak: per GPU)")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="N > 1: strong (default, configs[2]) or weak")
    ap.add_argument("--fixed-dur", action="store_true", help="the fixed 6.5 s variant of the corpus")
    ap.add_argument("--chunk-frames", type=int, default=1 << 18)
    ap.add_argument("--parity-utts", type=int, default=16)
    ap.add_argument("--synth-batch", type=int, default=SYNTH_BATCH, help="utterances per batched synthesis call of the corpus pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the synth256 / vtln109 / gen_data_files workloads")
    ap.add_argument("--io-utts", type=int, default=2048, help="utterances of the files-to-files gen_data workload (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
