"""Deterministic synthetic speech-like corpora for the benchmark and the parity tests.

There is no network and no dataset: an "LJSpeech-shaped" corpus is 13,100 utterances of ~6.5 s at 22.05 kHz.  Every
utterance has an F0 track (the "cached F0" the path takes as input: alternating voiced / unvoiced segments, slow
vibrato plus a random walk) and a waveform that is consistent with it (harmonics under a 3-formant envelope in voiced
segments, tilted noise elsewhere), quantised to int16 like a PCM16 wav file.

The generator is written with torch ops only so the same code runs on the CPU (tests, small cases) and on the GPU (the
full-size corpus in bench.py).  Parameters come from numpy's default_rng(seed) on the host; noise comes from a seeded
torch generator on the target device."""
import math

import numpy as np
import torch


def f0_track(rng, num_frames, frame_period_ms=5.0):
    """Frame-rate F0 in Hz (0 = unvoiced): voiced U[0.15, 0.8] s / unvoiced U[0.05, 0.3] s, edges unvoiced >= 0.1 s."""
    fp = frame_period_ms / 1000.0
    t = np.arange(num_frames) * fp
    f0 = np.zeros(num_frames)
    fb = rng.uniform(90.0, 250.0)
    tau = rng.uniform(0.8, 2.0)
    walk = np.cumsum(rng.standard_normal(num_frames)) * 0.03 / math.sqrt(max(num_frames, 1))
    contour = fb * 2.0 ** (0.2 * np.sin(2 * np.pi * t / tau) + walk)
    contour = np.clip(contour, 71.0, 400.0)
    pos = 0.1
    end = num_frames * fp - 0.1
    voiced = True
    while pos < end:
        dur = rng.uniform(0.15, 0.8) if voiced else rng.uniform(0.05, 0.3)
        if voiced:
            a, b = int(round(pos / fp)), int(round(min(pos + dur, end) / fp))
            f0[a:b] = contour[a:b]
        pos += dur
        voiced = not voiced
    return f0


def waveforms_from_f0(f0, fs, rngs, frame_period_ms=5.0, device="cpu", num_harmonics=48):
    """A batch of equal-length utterances: f0 [B, T] numpy (Hz) -> int16 [B, n] on `device`."""
    device = torch.device(device)
    f0 = np.atleast_2d(f0)
    B, T = f0.shape
    n = int((T - 1) * frame_period_ms * fs / 1000.0)
    t = torch.arange(n, dtype=torch.float64, device=device) / fs
    pos = t / (frame_period_ms / 1000.0)
    i0 = torch.clamp(pos.floor().long(), max=T - 2)
    w = (pos - i0)[None, :]
    f0t = torch.from_numpy(np.ascontiguousarray(f0)).to(device)
    a, b = f0t[:, i0], f0t[:, i0 + 1]
    v0, v1 = a > 0, b > 0
    # inside voiced runs interpolate, at the edges hold the voiced neighbour
    f0s = (1 - w) * torch.where(v0, a, b) + w * torch.where(v1, b, a)
    voiced = v0 | v1
    zero = torch.zeros((), dtype=torch.float64, device=device)
    gate = (torch.where(v0, 1.0 - w, zero) + torch.where(v1, w, zero)).float()  # 5 ms cross-fade at voicing edges
    phase = 2 * math.pi * torch.cumsum(torch.where(voiced, f0s, zero) / fs, 1)  # float64 keeps high harmonics coherent
    fc = torch.tensor([[500.0 * r.uniform(0.85, 1.15), 1500.0 * r.uniform(0.85, 1.15), 2500.0 * r.uniform(0.85, 1.15)]
                       for r in rngs], dtype=torch.float32, device=device)
    bw = (90.0, 130.0, 180.0)
    x = torch.zeros((B, n), dtype=torch.float32, device=device)
    f0f = f0s.float()
    for k in range(1, num_harmonics + 1):
        fk = f0f * k
        g = torch.zeros_like(fk)
        for q in range(3):
            g += 1.0 / (1.0 + ((fk - fc[:, q:q + 1]) / bw[q]) ** 2)
        amp = g / torch.clamp(fk / 500.0, min=1.0) * (fk < 0.45 * fs).float()
        x += amp * torch.sin((phase * k).remainder(2 * math.pi)).float()
    x = x * gate
    noise = torch.empty((B, n), device=device, dtype=torch.float32)
    gen = torch.Generator(device=device)
    for row, r in enumerate(rngs):  # one noise stream per utterance: an utterance does not depend on its batch
        gen.manual_seed(int(r.integers(0, 2 ** 31 - 1)))
        noise[row] = torch.randn(n, generator=gen, device=device, dtype=torch.float32)
    # one-pole tilt of the unvoiced noise, y[n] = x[n] + 0.7 y[n-1], as a short FIR (keeps it parallel)
    tilt = noise.clone()
    coef = 0.7
    for d in range(1, 8):
        tilt[:, d:] += coef * noise[:, :n - d]
        coef *= 0.7
    x = x / torch.clamp(x.abs().amax(1, keepdim=True), min=1e-3)
    y = x + (10 ** (-30 / 20.0)) * noise * gate + (10 ** (-20 / 20.0)) * tilt * (1.0 - gate) / 2.0
    y = 0.5 * y / torch.clamp(y.abs().amax(1, keepdim=True), min=1e-6)
    return torch.round(y * 32767.0).to(torch.int16)


def corpus_durations(utt_ids, seed, mean_dur=6.5, std_dur=0.0, min_dur=1.1, max_dur=10.1, dur_quantum=0.0):
    """Durations (s) of the utterances with the given global indices, exactly as make_corpus draws them (utterance u's first
    draw from default_rng(7919 * seed + u)) -- lets every rank of a sharded run know all lengths without synthesising anything."""
    out = []
    for u in utt_ids:
        d = mean_dur
        if std_dur != 0:
            d = float(np.clip(np.random.default_rng(7919 * seed + int(u)).normal(mean_dur, std_dur), min_dur, max_dur))
            if dur_quantum > 0:
                d = round(d / dur_quantum) * dur_quantum
        out.append(d)
    return np.array(out)


def make_corpus(num_utts, fs, seed, mean_dur=6.5, std_dur=0.0, min_dur=1.1, max_dur=10.1, frame_period_ms=5.0,
                device="cpu", batch=64, first_utt=0, utt_ids=None, dur_quantum=0.0):
    """Returns (list of int16 torch tensors on `device`, list of numpy F0 tracks).  Utterance u (global index first_utt + u, or
    utt_ids[u]) uses default_rng(7919 * seed + index); durations ~ N(mean_dur, std_dur^2) clipped to [min_dur, max_dur] and
    rounded to multiples of dur_quantum (equal-length utterances are synthesised `batch` at a time, so a quantum keeps the
    generator fast); std_dur = 0 gives the fixed-length variant."""
    if utt_ids is None:
        utt_ids = [first_utt + u for u in range(num_utts)]
    num_utts = len(utt_ids)
    rngs = [np.random.default_rng(7919 * seed + int(u)) for u in utt_ids]
    durs = [mean_dur if std_dur == 0 else float(np.clip(r.normal(mean_dur, std_dur), min_dur, max_dur)) for r in rngs]
    if std_dur != 0 and dur_quantum > 0:
        durs = [round(d / dur_quantum) * dur_quantum for d in durs]
    Ts = [int(round(d * 1000.0 / frame_period_ms)) + 1 for d in durs]
    f0s = [f0_track(r, T, frame_period_ms) for r, T in zip(rngs, Ts)]
    waves = [None] * num_utts
    by_len = {}
    for u, T in enumerate(Ts):
        by_len.setdefault(T, []).append(u)
    for T, idx in by_len.items():
        for s0 in range(0, len(idx), batch):
            sel = idx[s0:s0 + batch]
            w = waveforms_from_f0(np.stack([f0s[u] for u in sel]), fs, [rngs[u] for u in sel], frame_period_ms, device)
            for row, u in enumerate(sel):
                waves[u] = w[row]
    for u in range(num_utts):
        # the F0 track must have exactly the frame count WORLD derives from the sample count
        t_world = int(1000.0 * waves[u].numel() / fs / frame_period_ms) + 1
        if t_world < Ts[u]:
            f0s[u] = f0s[u][:t_world]
        elif t_world > Ts[u]:
            f0s[u] = np.concatenate((f0s[u], np.zeros(t_world - Ts[u])))
    return waves, f0s
