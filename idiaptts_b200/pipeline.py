"""Fused WORLD feature extraction / synthesis over ragged batches on one GPU (the engine behind WorldFeatLabelGen.gen_data
and Synthesiser.run_world_synth).

extract:   packed waveforms + cached F0  ->  [F, D] float32 rows  [coded_sp(num_coded_sps) | lf0 | vuv | bap(nap)]
           (WorldFeatLabelGen.extract_features + convert_from_world_features, world/WorldFeatLabelGen.py:810-889, :765-776)
           plus per-column sum / sum-of-squares in fp64 (MeanStdDevExtractor.add_sample, misc/normalisation/...:43-47).
           Frames are processed in chunks so the only large intermediate (the float32 spectral-envelope plane of one
           chunk) stays bounded; every kernel launch goes to the caller's current stream, nothing synchronises.
synthesise: [F, D] rows -> packed waveforms (Synthesiser.run_world_synth, src/Synthesiser.py:39-80: convert_to_world_features,
           decode_sp, world_features_to_raw)."""
import math

import numpy as np
import torch

from . import ops


class WorldAnalyzer:
    def __init__(self, fs, num_coded_sps=60, mgc_alpha=None, hop_size_ms=5.0, n_fft=None, f0_silence_threshold=30,
                 lf0_zero=0, chunk_frames=1 << 18, device="cuda", sp_type="mcep", mgc_gamma=-1.0 / 3.0):
        """sp_type: "mcep" (pysptk.mcep, the default recipe) or "mgc" (pysptk.mgcep with mgc_gamma)."""
        from .compat.pysptk import mcepalpha
        assert sp_type in ("mcep", "mgc")
        self.sp_type = sp_type
        self.gamma = float(mgc_gamma) if sp_type == "mgc" else 0.0
        self.fs = int(fs)
        self.num_coded_sps = int(num_coded_sps)
        self.alpha = float(mcepalpha(fs) if mgc_alpha is None else mgc_alpha)
        self.hop_size_ms = float(hop_size_ms)
        self.n_fft = int(n_fft) if n_fft is not None else ops.get_cheaptrick_fft_size(fs)
        self.nap = ops.get_num_aperiodicities(fs)
        self.f0_silence_threshold = f0_silence_threshold
        self.lf0_zero = lf0_zero
        self.chunk_frames = int(chunk_frames)
        self.device = torch.device(device)
        self.dim = self.num_coded_sps + 2 + self.nap
        # upload the warping matrices once
        ops.McepTables.get(self.num_coded_sps - 1, self.alpha, self.n_fft, self.device)
        self._sp = None
        self.iters = None  # optional int32 [F] buffer: Newton iterations per frame (bench.py's FLOP accounting)

    def _sp_buffer(self, frames):
        K = self.n_fft // 2 + 1
        if self._sp is None or self._sp.shape[0] < frames:
            # rows padded to a multiple of 8 floats: the mel-cepstrum kernel reads them with aligned 16-byte loads
            self._sp = torch.empty((frames, (K + 7) // 8 * 8), dtype=torch.float32, device=self.device)
        return self._sp[:, :K]

    def extract(self, batch, feats=None, sums=None, status=None, events=None, pre_chunk=None, post_chunk=None):
        """batch: ops.RaggedBatch.  Returns (feats [F, dim] float32, sums [2*dim] float64, status int32[1]); `sums` is
        accumulated into when given (corpus statistics over several calls).  events: optional list that receives
        (kernel name, frames, start event, end event) per launch (bench.py's per-kernel timing).  pre_chunk / post_chunk:
        optional callables (lo, hi) run before / after the launches of every frame chunk (extract_from_host uses them to
        order the chunk behind its host-to-device copy and to start the device-to-host copy of its feature rows)."""
        def timed(name, frames, fn):
            if events is None:
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            events.append((name, frames, e0, e1))
            return r

        F = batch.num_frames
        D = self.num_coded_sps
        if feats is None:
            feats = torch.empty((F, self.dim), dtype=torch.float32, device=self.device)
        if sums is None:
            sums = torch.zeros(2 * self.dim, dtype=torch.float64, device=self.device)
        if status is None:
            status = ops.new_status(self.device)
        if F == 0:  # nothing to launch
            return feats, sums, status
        flat = feats.view(-1)
        # lf0 / vuv straight into their columns
        timed("lf0_vuv", F, lambda: ops.lf0_vuv(batch.f0, batch.frame_off, self.f0_silence_threshold, self.lf0_zero,
                                                lf0_out=flat[D:], vuv_out=flat[D + 1:], out_stride=self.dim))
        chunk = min(self.chunk_frames, max(F, 1))
        sp = self._sp_buffer(chunk)
        for lo in range(0, F, chunk):
            hi = min(F, lo + chunk)
            spc = sp[:hi - lo]
            nf = hi - lo
            if pre_chunk is not None:
                pre_chunk(lo, hi)
            timed("cheaptrick", nf, lambda: ops.cheaptrick(batch, fft_size=self.n_fft, status=status, frame_lo=lo, frame_hi=hi,
                                                           out=spc))
            it = None if self.iters is None else self.iters[lo:hi]
            if self.gamma == 0.0:
                timed("mcep", nf, lambda: ops.mcep(spc, D - 1, self.alpha, is_power=True, out=flat[lo * self.dim:],
                                                   out_stride=self.dim, status=status, iters=it))
            else:
                timed("mgcep", nf, lambda: ops.mgcep(spc, D - 1, self.alpha, self.gamma, is_power=True, out=flat[lo * self.dim:],
                                                     out_stride=self.dim, status=status, iters=it))
            coarse, voiced, _ = timed("d4c", nf, lambda: ops.d4c_coarse(batch, status=status, frame_lo=lo, frame_hi=hi))
            timed("bap_from_coarse", nf, lambda: ops.bap_from_coarse(coarse, voiced, self.fs, self.n_fft,
                                                                     out=flat[lo * self.dim + D + 2:], out_stride=self.dim))
            if post_chunk is not None:
                post_chunk(lo, hi)
        timed("stats", F, lambda: ops.stats_accumulate(feats, sums))
        return feats, sums, status

    def extract_from_host(self, host, feats_host, dev_buffers=None, sums=None):
        """End-to-end extraction of a corpus shard that lives in pinned host memory: the packed waveform is copied in pieces
        on a copy stream while earlier frame chunks are analysed, and every chunk's feature rows travel back to the pinned
        `feats_host` [F, dim] on a second copy stream while the next chunk is analysed (gen_data's wav -> npz path without
        the file IO).  host: dict of pinned CPU tensors x (int16 / float), so (sample offsets), f0, t, fo (frame offsets),
        fu (frame -> utterance).  Returns (feats (device), sums, status, dev_buffers); pass dev_buffers back in to reuse
        the device allocations.  Everything is asynchronous; synchronise the current stream before reading feats_host."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if dev_buffers is None:
            dev_buffers = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()}
            dev_buffers["_in"] = torch.cuda.Stream(dev)
            dev_buffers["_out"] = torch.cuda.Stream(dev)
            dev_buffers["feats"] = torch.empty((host["f0"].numel(), self.dim), dtype=torch.float32, device=dev)
        s_in, s_out = dev_buffers["_in"], dev_buffers["_out"]
        d = dev_buffers
        s_in.wait_stream(cur)   # earlier work on the current stream may still read the device buffers
        s_out.wait_stream(cur)
        so_np, fu_np = host["so"].numpy(), host["fu"].numpy()
        F = host["f0"].numel()
        with torch.cuda.stream(s_in):
            for k in ("so", "fo", "f0", "t", "fu"):
                d[k].copy_(host[k], non_blocking=True)
            small_done = torch.cuda.Event()
            small_done.record(s_in)
        cur.wait_event(small_done)
        batch = ops.RaggedBatch(d["x"], d["so"], d["f0"], d["t"], d["fo"], d["fu"], self.fs)
        state = {"copied": 0, "out_events": []}
        chunk = min(self.chunk_frames, max(F, 1))
        # issue all waveform pieces up front on the copy stream, one event per frame chunk
        piece_events = []
        with torch.cuda.stream(s_in):
            for lo in range(0, F, chunk):
                hi = min(F, lo + chunk)
                end = int(so_np[int(fu_np[hi - 1]) + 1])   # last sample the chunk's utterances can touch
                if end > state["copied"]:
                    d["x"][state["copied"]:end].copy_(host["x"][state["copied"]:end], non_blocking=True)
                    state["copied"] = end
                ev = torch.cuda.Event()
                ev.record(s_in)
                piece_events.append(ev)

        def pre_chunk(lo, hi):
            cur.wait_event(piece_events[lo // chunk])

        def post_chunk(lo, hi):
            ev = torch.cuda.Event()
            ev.record(cur)
            s_out.wait_event(ev)
            with torch.cuda.stream(s_out):
                feats_host[lo:hi].copy_(d["feats"][lo:hi], non_blocking=True)

        feats, sums, status = self.extract(batch, feats=d["feats"], sums=sums, pre_chunk=pre_chunk, post_chunk=post_chunk)
        cur.wait_stream(s_out)
        return feats, sums, status, dev_buffers

    def kernel_launches(self, num_frames):
        """Number of libb200world kernels one extract() call launches (for bench.py's gpu_launches)."""
        chunks = max(1, math.ceil(num_frames / self.chunk_frames))
        return 1 + 5 * chunks + 1   # lf0_vuv; per chunk cheaptrick, mcep, d4c (single-precision pass + fp64 re-evaluation), bap; stats


def mean_std_from_sums(sums, n, dim):
    """MeanStdDevExtractor.get_params / combine_mean_std (misc/normalisation/MeanStdDevExtractor.py:49-53, :230-241)."""
    s = np.asarray(sums[:dim], np.float64)
    q = np.asarray(sums[dim:2 * dim], np.float64)
    mean = s / n
    var = q / n - mean ** 2
    var = np.where(var < 0, 0.0, var)
    return mean, np.sqrt(var)


class WorldSynthesizer:
    def __init__(self, fs, num_coded_sps=60, mgc_alpha=None, hop_size_ms=5.0, n_fft=None, f0_silence_threshold=30, lf0_zero=0,
                 device="cuda", precision="fast", sp_type="mcep", mgc_gamma=-1.0 / 3.0, post_filtering=False):
        """precision: "fast" (default) = single-precision per-pulse transforms (ops.synth_render), "f64" = double precision.
        sp_type "mgc": the coded spectrum is a generalised mel-cepstrum with mgc_gamma (AudioProcessing.mgc_to_amp_sp);
        post_filtering: nnmnkwii's merlin_post_filter on the coded spectrum before decoding (AudioProcessing.decode_sp)."""
        from .compat.pysptk import mcepalpha
        assert sp_type in ("mcep", "mgc")
        self.gamma = float(mgc_gamma) if sp_type == "mgc" else 0.0
        self.post_filtering = bool(post_filtering)
        self.fs = int(fs)
        self.num_coded_sps = int(num_coded_sps)
        self.alpha = float(mcepalpha(fs) if mgc_alpha is None else mgc_alpha)
        self.hop_size_ms = float(hop_size_ms)
        self.n_fft = int(n_fft) if n_fft is not None else ops.get_cheaptrick_fft_size(fs)
        self.nap = ops.get_num_aperiodicities(fs)
        self.f0_silence_threshold = f0_silence_threshold
        self.lf0_zero = lf0_zero
        self.device = torch.device(device)
        self.precision = precision
        ops.McepTables.get(self.num_coded_sps - 1, self.alpha, self.n_fft, self.device)

    def synthesize(self, feats, frame_off, preemphasis=0.0, out_dtype=torch.float32, events=None, frame_off_host=None):
        """feats [F, D + 2 + nap] float32 rows [coded_sp | lf0 | vuv | bap] on the device; frame_off int64 [U+1] on the device.
        Returns (y packed, out_off numpy int64 [U+1], status).  events: optional list that receives (kernel name, units, start
        event, end event) per launch (bench.py's per-kernel timing)."""
        D = self.num_coded_sps
        F = feats.shape[0]
        assert feats.shape[1] == D + 2 + self.nap, "WORLD requires all features to be present."
        lf0 = feats[:, D].double()
        vuv = (feats[:, D + 1] >= 0.5)
        f0 = torch.exp(lf0)
        vuv = vuv & ~(f0 < self.f0_silence_threshold)
        f0 = torch.where(vuv, f0, torch.full_like(f0, float(self.lf0_zero)))
        # decode_sp: amp = exp(Re mgc2sp) as float32 (AudioProcessing.py:256), then pow_sp = amp^2 in float64 (W:924)
        coded, cstride = feats, feats.shape[1]
        if self.post_filtering:  # decode_sp filters with fs_to_mgc_alpha(fs), whatever alpha the features were coded with (A:310)
            from .compat.pysptk import mcepalpha
            coded = ops.merlin_post_filter(feats[:, :D], float(mcepalpha(self.fs)), self.n_fft).float().contiguous()
            cstride = D
        # the fast path keeps both spectral planes in float32 (half the HBM traffic of the three kernels that touch them; the
        # per-pulse kernel interpolates and clamps them in double either way)
        plane_dtype = torch.float32 if (self.precision == "fast" and self.n_fft == 1024) else torch.float64
        if self.gamma == 0.0:
            pow_sp = ops._timed(events, "mc2sp", F, lambda: ops.mc2sp(coded, self.alpha, self.n_fft, scale=1.0, do_exp=True,
                                                                     out_dtype=plane_dtype, order=D - 1, mc_stride=cstride,
                                                                     square=True))
        else:
            amp = ops._timed(events, "mgc2sp", F, lambda: ops.mgc2sp(coded, self.alpha, self.gamma, self.n_fft, out_dtype=torch.float32,
                                                                    order=D - 1, mgc_stride=cstride))
            pow_sp = (amp.double() ** 2).to(plane_dtype)
        bap = feats[:, D + 2:].double().contiguous()
        ap = ops._timed(events, "decode_ap", F, lambda: ops.decode_aperiodicity(bap, self.fs, self.n_fft, out_dtype=plane_dtype))
        # (Measured: decoding the two planes on a side stream while the sequential pulse placement runs on this one is SLOWER,
        # 23.2 ms against 19.8 ms for 256 utterances -- scripts/gpu_synth_phases.py -- so the stages stay in one stream.)
        # (frame_off_host given: nothing below reads back from the device, the host can queue the next batch right away)
        plan = ops.synth_timebase(f0.contiguous(), frame_off, self.fs, self.n_fft, self.hop_size_ms, events=events,
                                  frame_off_host=frame_off_host)
        de = float(preemphasis)
        y, out_off, status = ops.synth_render(plan, pow_sp, ap, deemphasis=de, out_dtype=torch.float64 if de != 0.0 else out_dtype,
                                              events=events, precision=self.precision, sync_counts=frame_off_host is None)
        return y, out_off, status

    def kernel_launches(self, num_utts, batch_utts=256):
        """Number of libb200world kernels synthesize_corpus launches (for bench.py's gpu_launches): per batch mc2sp, decode_ap,
        the four time-base kernels, render, overlap-add."""
        return 8 * max(1, math.ceil(num_utts / batch_utts))

    def synthesize_corpus(self, feats, frame_off_host, batch_utts=256, out=None, feats_host=None, out_host=None, events=None):
        """Synthesis of many utterances in batches of `batch_utts` (Synthesiser.run_world_synth over a corpus; the batch bounds
        the per-pulse response buffer).  feats [F, dim] float32 on the device, or -- end to end -- `feats_host` [F, dim] in
        pinned host memory, copied batch by batch on a copy stream one batch ahead of the synthesis; frame_off_host: numpy int64
        [U+1].  Output: `out` (device, float32, packed [sum y_len]) and / or `out_host` (pinned; every batch's samples travel
        back on a second copy stream while the next batch is synthesised).  Returns (sample offsets numpy int64 [U+1], status).
        Synchronise the current stream before reading out_host."""
        dev = self.device
        fo = np.asarray(frame_off_host, np.int64)
        U = len(fo) - 1
        ylen = (np.diff(fo) * self.hop_size_ms * self.fs / 1000).astype(np.int64)
        out_off = np.concatenate(([0], np.cumsum(ylen)))
        cur = torch.cuda.current_stream(dev)
        status = ops.new_status(dev)
        if not hasattr(self, "_streams"):
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._streams
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        bounds = [(u0, min(U, u0 + batch_utts)) for u0 in range(0, U, batch_utts)]
        fo_dev = torch.from_numpy(fo).to(dev)   # one upload; the per-batch offsets are slices of it

        def fetch(b):  # device rows of batch b (staged from the host one batch ahead)
            u0, u1 = bounds[b]
            if feats_host is None:
                return feats[fo[u0]:fo[u1]], None
            with torch.cuda.stream(s_in):
                buf = torch.empty((int(fo[u1] - fo[u0]), feats_host.shape[1]), dtype=torch.float32, device=dev)
                buf.copy_(feats_host[fo[u0]:fo[u1]], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            return buf, ev

        nxt = fetch(0) if bounds else None
        for b, (u0, u1) in enumerate(bounds):
            rows, ev = nxt
            nxt = fetch(b + 1) if b + 1 < len(bounds) else None
            if ev is not None:
                cur.wait_event(ev)
                rows.record_stream(cur)
            off_b = fo_dev[u0:u1 + 1] - fo_dev[u0]
            y, _, st = self.synthesize(rows, off_b, events=events, frame_off_host=fo[u0:u1 + 1] - fo[u0])
            status |= st
            if out is not None:
                out[out_off[u0]:out_off[u1]].copy_(y)
            if out_host is not None:
                done = torch.cuda.Event()
                done.record(cur)
                s_out.wait_event(done)
                with torch.cuda.stream(s_out):
                    out_host[out_off[u0]:out_off[u1]].copy_(y, non_blocking=True)
                y.record_stream(s_out)
        cur.wait_stream(s_out)
        return out_off, status
