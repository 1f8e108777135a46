#!/usr/bin/env python
"""Benchmark of the WORLD feature hot path (BASELINE.json metric: audio-seconds/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--utts U]

Workload (config.workload = "ljspeech_extract"): BASELINE.json configs[1], the LJSpeech-shaped synthetic corpus
(13,100 utterances x 6.5 s at 22.05 kHz, int16 PCM + cached F0 track) -> WorldFeatLabelGen-style features
(mcep60, lf0, vuv, bap) + corpus normalisation statistics.  One step = one pass of the extraction over the corpus shard
of every rank (weak scaling: every rank owns a full-size shard with its own seeds); ranks only exchange the
statistics (one NCCL all-reduce per step, inside the timed region).

value   : audio-seconds per second with the inputs resident in HBM (CUDA events, max over ranks).
e2e     : the same pass through host buffers: pinned int16 wave + F0 host->device, extraction, features + statistics
          device->host, all inside the timed region.
roofline: the kernel with the largest share of the step, timed live with CUDA events on the launching stream.
cpu_baseline / --impl reference: the CPU oracle (C restatement of the WORLD/SPTK algorithms; pyworld/pysptk are not installable here)
          on all host cores over a bounded sample of the same workload.
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
NUM_CODED_SPS = 60
METRIC = "audio-seconds/s (WORLD analysis: CheapTrick + D4C + mcep60 + lf0/vuv/bap + stats, cached F0)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline of the default run, and the whole of --impl reference)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_oracle_throughput(waves, f0s, fs, alpha, cores, steps=1, warmup=0):
    """audio-seconds/s of the CPU oracle (oracle/c/world_oracle.c: C restatement of WORLD / SPTK, one utterance per call,
    the GIL is released inside) over the given sample with a pool of `cores` threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import world_c
    world_c.lib()

    def one(job):
        w, f = job
        feats = world_c.extract(w, fs, f, NUM_CODED_SPS, alpha)
        return feats.sum(0, dtype=np.float64), len(w) / fs

    jobs = [(np.ascontiguousarray(w), f) for w, f in zip(waves, f0s)]
    with ThreadPoolExecutor(max_workers=cores) as pool:
        for _ in range(warmup):
            list(pool.map(one, jobs[:cores]))
        times = []
        audio = 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            res = list(pool.map(one, jobs))
            times.append(time.perf_counter() - t0)
            audio = sum(r[1] for r in res)
    return audio / (sum(times) / len(times)), sum(times) / len(times), audio


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch  # noqa: F401  (only for the synthetic generator, CPU)
    from idiaptts_b200 import synthetic
    from oracle import sptk_np
    cores = len(os.sched_getaffinity(0))
    alpha = float(sptk_np.mcepalpha(FS))
    n_utts = max(cores, min(2 * cores, 64))  # bounded sample: about one or two 6.5 s utterances per core and step
    waves, f0s = synthetic.make_corpus(n_utts, FS, seed=2, mean_dur=DUR, device="cpu")
    waves = [w.numpy() for w in waves]
    value, sec, audio = cpu_oracle_throughput(waves, f0s, FS, alpha, cores, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ljspeech_extract", "fs": FS, "utt_seconds": DUR, "num_coded_sps": NUM_CODED_SPS,
                       "sample_utts": n_utts, "note": "CPU oracle = C restatement of WORLD/SPTK, gcc -O3 -march=native, one thread per utterance (pyworld/pysptk unavailable offline)"},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "%d utterances x %.1f s per step" % (n_utts, DUR)},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from idiaptts_b200 import _lib, ops, pipeline, synthetic
    from idiaptts_b200.compat.pysptk import mcepalpha

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    alpha = float(mcepalpha(FS))
    utts = args.utts
    # every rank synthesises its own shard (weak scaling), distinct seeds
    t0 = time.time()
    waves, f0s = synthetic.make_corpus(utts, FS, seed=2, mean_dur=DUR, device=dev, batch=64, first_utt=rank * utts)
    lens = np.array([w.numel() for w in waves], np.int64)
    flens = np.array([len(f) for f in f0s], np.int64)
    audio_s = float(lens.sum()) / FS
    x_dev = torch.cat(waves)
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

    def to_dev():
        return ops.RaggedBatch(host["x"].to(dev, non_blocking=True), host["so"].to(dev, non_blocking=True),
                               host["f0"].to(dev, non_blocking=True), host["t"].to(dev, non_blocking=True),
                               host["fo"].to(dev, non_blocking=True), host["fu"].to(dev, non_blocking=True), FS)

    batch = to_dev()
    F = batch.num_frames
    an = pipeline.WorldAnalyzer(FS, NUM_CODED_SPS, alpha, device=dev, chunk_frames=args.chunk_frames)
    an.iters = torch.zeros(F, dtype=torch.int32, device=dev)
    feats = torch.empty((F, an.dim), dtype=torch.float32, device=dev)
    stat_buf = torch.zeros(2 * an.dim + 1, dtype=torch.float64, device=dev)

    def step(b, events=None):
        stat_buf.zero_()
        _, _, status = an.extract(b, feats=feats, sums=stat_buf[:2 * an.dim], events=events)
        stat_buf[2 * an.dim] = float(F)
        if world > 1:
            dist.all_reduce(stat_buf)  # the path's only exchange step: corpus normalisation statistics
        return status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        status = step(batch)
    barrier()
    ops.raise_for_status(status, "extract")

    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(batch, events)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = audio_s * world / (ms_step / 1e3)
    stats_host = stat_buf.cpu().numpy()

    # per-kernel shares (rank 0's stream)
    per = {}
    for name, frames, a, b in events:
        d = per.setdefault(name, [0.0, 0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
        d[2] += frames
    total_k = sum(v[0] for v in per.values())
    top = max(per, key=lambda k: per[k][0])
    H = FS * 0.005
    K = an.n_fft // 2 + 1
    alg_bytes_per_frame = {
        "cheaptrick": 2 * H + 20 + 4 * K,              # int16 samples of one hop + f0/t/frame_utt in, float32 envelope row out
        "mcep": 4 * K + 4 * NUM_CODED_SPS,             # float32 envelope row in, 60 float32 coefficients out
        "d4c": 2 * H + 20 + 8 * an.nap + 1,            # samples + f0/t in, coarse aperiodicity + voiced flag out
        "bap_from_coarse": 8 * an.nap + 1 + 4 * an.nap,
        "lf0_vuv": 8 + 8,
        "stats": 4 * an.dim,
    }
    peaks, peak_src = measured_peaks()
    top_ms = per[top][0] / per[top][1]
    top_frames = per[top][2] / per[top][1]
    achieved = alg_bytes_per_frame[top] * top_frames / (top_ms / 1e3) / 1e9
    # the tensor-core kernel of the step, for the record: algorithmic FLOPs of the warp contractions (initial value +
    # mean Newton passes x (freqt+FFT and IFFT+frqtr as GEMMs)), not counting the 3x of the TF32 split nor the solves
    mean_it = float(an.iters.float().mean().item())
    m = NUM_CODED_SPS - 1
    flops_frame = 2.0 * K * (m + 2) + mean_it * 2.0 * K * ((m + 1) + (2 * m + 1))
    mcep_ms = per["mcep"][0] / per["mcep"][1]
    mcep_frames = per["mcep"][2] / per["mcep"][1]
    mcep_tflops = flops_frame * mcep_frames / (mcep_ms / 1e3) / 1e12
    # measured DRAM traffic of the dominant kernel (ncu --set full capture of this command, bytes per frame -> per launch)
    traffic = None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            traffic = float(json.load(f)[top]["dram_bytes_per_frame"]) * top_frames
    except Exception:
        pass
    # compute side of the dominant kernels: algorithmic fp64 FFT flops (5 N log2 N per complex FFT, nothing else counted)
    # against the fp64 FMA rate measured right here with a pure-FMA probe kernel
    probe = torch.zeros(8, dtype=torch.float64, device=dev)
    lib = _lib.load()
    lib.b2w_probe_fp64_fma(1 << 12, probe.data_ptr(), torch.cuda.current_stream().cuda_stream)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    nfma = lib.b2w_probe_fp64_fma(1 << 15, probe.data_ptr(), torch.cuda.current_stream().cuda_stream)
    p1.record()
    torch.cuda.synchronize()
    fp64_peak = 2.0 * nfma / (p0.elapsed_time(p1) / 1e3) / 1e12
    n4 = 2048 if FS <= 24000 else 4096
    lo_f, hi_f = 0, min(F, an.chunk_frames)
    _, vflag, _ = ops.d4c_coarse(batch, frame_lo=lo_f, frame_hi=hi_f)
    f0_pos = float((batch.f0[lo_f:hi_f] > 0).float().mean().item())     # frames that enter D4C (1 FFT: LoveTrain)
    full = float(vflag.float().mean().item())                            # frames that run the whole of D4C
    fft_flop = lambda n: 5.0 * n * math.log2(n)
    d4c_flop_frame = f0_pos * fft_flop(n4) + full * (2 + (an.nap + 1) // 2) * fft_flop(n4)
    ct_flop_frame = 3 * fft_flop(an.n_fft // 2)
    def tfl(name, flop_frame):
        return flop_frame * (per[name][2] / per[name][1]) / (per[name][0] / per[name][1] / 1e3) / 1e12
    compute = {"fp64_fma_peak_tflops_measured": round(fp64_peak, 2),
               "d4c": {"fft_tflops": round(tfl("d4c", d4c_flop_frame), 3), "frac_of_fp64_peak": round(tfl("d4c", d4c_flop_frame) / fp64_peak, 4),
                       "frames_with_f0": round(f0_pos, 4), "frames_full_d4c": round(full, 4)},
               "cheaptrick": {"fft_tflops": round(tfl("cheaptrick", ct_flop_frame), 3),
                              "frac_of_fp64_peak": round(tfl("cheaptrick", ct_flop_frame) / fp64_peak, 4)},
               "note": "FFT flops only (5 N log2 N per complex FFT); windows, smoothing, order statistics, log/exp are not counted. "
                       "ncu (profiles/): these kernels are bound by the L1/shared-memory data pipe and the fp64 pipe together"}
    roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src, "compute": compute,
                "share_of_step": per[top][0] / total_k,
                "shares": {k: round(v[0] / total_k, 4) for k, v in per.items()},
                "avg_launch_ms": {k: round(v[0] / v[1], 3) for k, v in per.items()},
                "mcep_tensor": {"bound": "tensor", "achieved": mcep_tflops, "unit": "TFLOP/s (algorithmic, fp32-equivalent)",
                                "peak": peaks.get("bf16_tflops_sustained"), "peak_unit": "TFLOP/s bf16 dense (measured)",
                                "mean_newton_passes": mean_it,
                                "note": "tcgen05 kind::tf32 with the 3xTF32 split; the kernel is bound by issue slots / latency of the exp "
                                        "epilogue and the per-frame 60x60 register-resident solves, not by the tensor pipe"},
                "note": "compute-bound kernel (fp64 FFTs / fp32 contractions): the HBM fraction is reported because the "
                        "metric asks for it, see DESIGN.md for the per-kernel bounds"}

    # ---- timed region 2: end to end through host buffers ---------------------------------------------------------------
    feats_host = torch.empty((F, an.dim), dtype=torch.float32).pin_memory()
    stats_pinned = torch.empty(2 * an.dim + 1, dtype=torch.float64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = feats_host.numel() * 4 + stats_pinned.numel() * 8

    e2e_state = {"buf": None}

    def e2e_step():
        # the public end-to-end call: pinned host corpus in, pinned host features + statistics out, copies overlapped with
        # the analysis of neighbouring frame chunks (pipeline.WorldAnalyzer.extract_from_host)
        stat_buf.zero_()
        _, _, _, e2e_state["buf"] = an.extract_from_host(host, feats_host, dev_buffers=e2e_state["buf"], sums=stat_buf[:2 * an.dim])
        stat_buf[2 * an.dim] = float(F)
        if world > 1:
            dist.all_reduce(stat_buf)
        stats_pinned.copy_(stat_buf, non_blocking=True)

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
    e2e_value = audio_s * world / (float(t2.item()) / args.steps / 1e3)

    # ---- secondary measurements (rank 0): BASELINE.json configs[3] (batched synthesis of 256 utterances from the features just
    # extracted) and configs[4] (Neural-VTLN warp forward + backward, VCTK-shaped) -- reported next to the headline, not part of it
    secondary = None
    if rank == 0 and not args.no_secondary:
        def t_ms(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            return best
        nu = min(256, utts)
        f_end = int(frame_off[nu])
        syn = pipeline.WorldSynthesizer(FS, NUM_CODED_SPS, alpha, device=dev)
        sub_feats, sub_off = feats[:f_end].contiguous(), batch.frame_off[:nu + 1].contiguous()
        ms_syn = t_ms(lambda: syn.synthesize(sub_feats, sub_off))
        syn_audio = f_end * 0.005
        n_v, spk, utt_v, T_v = 60, 109, 4, 1301
        rows_v = spk * utt_v * T_v
        gen_v = torch.Generator(device=dev).manual_seed(5)
        xv = torch.randn((rows_v, n_v), generator=gen_v, device=dev)
        gv = torch.randn((rows_v, n_v), generator=gen_v, device=dev)
        av = (torch.rand(spk, generator=gen_v, device=dev) * 0.4 - 0.2).repeat_interleave(utt_v * T_v).contiguous()
        ms_vf = t_ms(lambda: ops.allpass_forward(xv, av, n_v), 5)
        ms_vb = t_ms(lambda: ops.allpass_backward(gv, xv, av, n_v), 5)
        secondary = {"synthesis": {"workload": "batched WORLD synthesis of %d utterances from mgc60+lf0+vuv+bap" % nu, "ms": round(ms_syn, 3),
                                   "audio_s_per_s": syn_audio / (ms_syn / 1e3)},
                     "analysis_plus_synthesis_audio_s_per_s": 1.0 / (1.0 / (audio_s / (ms_step / 1e3)) + 1.0 / (syn_audio / (ms_syn / 1e3))),
                     "vtln": {"workload": "all-pass warp, 109 speakers x 4 utts x 1301 frames, n = 60, one alpha per speaker",
                              "fwd_ms": round(ms_vf, 4), "bwd_ms": round(ms_vb, 4),
                              "fwd_frac_of_hbm_peak": rows_v * (8 * n_v + 4) / (ms_vf / 1e3) / 1e9 / peaks["hbm_gbs"],
                              "bwd_frac_of_hbm_peak": rows_v * (12 * n_v + 8) / (ms_vb / 1e3) / 1e9 / peaks["hbm_gbs"]}}
        # F0 stage (not part of the headline, which is quoted with cached F0): DIO + StoneMask over the first 512
        # utterances, and what the un-cached pyworld.wav2world path would run at
        nf0 = min(512, utts)
        sub = ops.RaggedBatch(batch.x[:int(sample_off[nf0])], batch.sample_off[:nf0 + 1].contiguous(), batch.f0[:int(frame_off[nf0])],
                              batch.t[:int(frame_off[nf0])], batch.frame_off[:nf0 + 1].contiguous(),
                              batch.frame_utt[:int(frame_off[nf0])], FS)
        ms_dio = t_ms(lambda: ops.dio(sub), 2)
        f0_dio = ops.dio(sub)
        ms_sm = t_ms(lambda: ops.stonemask(sub, f0_dio), 2)
        f0_audio = int(sample_off[nf0]) / FS
        f0_rate = f0_audio / ((ms_dio + ms_sm) / 1e3)
        dio_dfma = 2.0 * int(sample_off[nf0]) * ops.dio_fir_taps(FS)   # flop of the two FIR stages
        secondary["f0_stage"] = {"workload": "DIO + StoneMask (pyworld.wav2world's F0 half), %d utterances" % nf0,
                                 "dio_ms": round(ms_dio, 3), "stonemask_ms": round(ms_sm, 3), "audio_s_per_s": f0_rate,
                                 "dio_fir_tflops_fp64": dio_dfma / (ms_dio / 1e3) / 1e12,
                                 "analysis_with_f0_estimation_audio_s_per_s": 1.0 / (1.0 / (audio_s / (ms_step / 1e3)) + 1.0 / f0_rate)}
        del xv, gv, av, sub_feats, sub, f0_dio

    if rank == 0:
        # ---- CPU baseline on a bounded sample of the same workload (N = 1 only) -------------------------------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # separate process: no fork of a CUDA-initialised interpreter
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                               capture_output=True, text=True)
            try:
                cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception:
                cpu = {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "port", "sample": "failed: " + r.stderr[-300:]}
        n = float(stats_host[2 * an.dim])
        mean, std = pipeline.mean_std_from_sums(stats_host, n, an.dim)
        line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32",
                "data": "synthetic",
                "config": {"workload": "ljspeech_extract", "utts_per_gpu": utts, "fs": FS, "utt_seconds": DUR,
                           "frames_per_gpu": int(F), "audio_seconds_per_gpu": audio_s, "num_coded_sps": NUM_CODED_SPS,
                           "mgc_alpha": alpha, "fft_size": an.n_fft, "num_bap": an.nap, "chunk_frames": an.chunk_frames,
                           "l2": "inputs (%.2f GB int16) and outputs (%.2f GB) exceed the 126 MB L2" % (
                               host["x"].numel() * 2 / 1e9, feats.numel() * 4 / 1e9),
                           "corpus_gen_s": round(gen_s, 1), "mean_c0": float(mean[0]), "std_c0": float(std[0])},
                "clocks": clocks, "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d),
                                          "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(args.steps * an.kernel_launches(F)), "roofline": roofline, "cpu_baseline": cpu,
                "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=UTTS, help="utterances per GPU (default: the full LJSpeech-shaped corpus)")
    ap.add_argument("--chunk-frames", type=int, default=1 << 18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the synthesis / VTLN side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
