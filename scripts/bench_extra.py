"""Secondary measurements (not the driver's bench contract): BASELINE.json configs[3] (batched WORLD synthesis of 256
utterances from acoustic-model-shaped features) and configs[4] (Neural-VTLN all-pass warp fwd + bwd on VCTK-shaped mgc
batches).  Prints one JSON line per config; CUDA events, inputs resident in HBM."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops, pipeline, synthetic  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def synthesis():
    fs, U = 22050, 256
    waves, f0s = synthetic.make_corpus(U, fs, seed=4, mean_dur=6.5, device=dev)
    lens = np.array([w.numel() for w in waves])
    fl = np.array([len(f) for f in f0s])
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    batch = ops.RaggedBatch(torch.cat(waves), up(np.concatenate(([0], np.cumsum(lens))).astype(np.int64)), up(np.concatenate(f0s)),
                            up(np.concatenate([np.arange(n) * 5.0 / 1000.0 for n in fl])), up(np.concatenate(([0], np.cumsum(fl))).astype(np.int64)),
                            up(np.repeat(np.arange(U, dtype=np.int32), fl)), fs)
    an = pipeline.WorldAnalyzer(fs, 60, device=dev)
    feats, _, _ = an.extract(batch)
    # acoustic-model-shaped outputs: analysis features + small noise, soft vuv
    g = torch.Generator(device=dev).manual_seed(4)
    sd = feats.std(0, keepdim=True)
    noisy = feats + 0.05 * sd * torch.randn(feats.shape, generator=g, device=dev)
    noisy[:, 61] = torch.clamp(feats[:, 61] + 0.1 * torch.randn(feats.shape[0], generator=g, device=dev), 0, 1)
    noisy[:, 62:] = torch.clamp(noisy[:, 62:], max=0.0)
    syn = pipeline.WorldSynthesizer(fs, 60, device=dev)
    dbg = {}
    y, out_off, st = syn.synthesize(noisy.contiguous(), batch.frame_off)
    ops.raise_for_status(st, "synth")
    ms = timed(lambda: syn.synthesize(noisy.contiguous(), batch.frame_off), steps=3, warmup=2)
    audio_s = float(out_off[-1]) / fs
    print(json.dumps({"config": "batched WORLD synthesis, 256 utterances x 6.5 s @ 22.05 kHz from mgc60+lf0+vuv+bap", "ms": ms,
                      "audio_seconds": audio_s, "audio_s_per_s": audio_s / (ms / 1e3), "samples": int(out_off[-1]),
                      "finite": bool(torch.isfinite(y).all())}), flush=True)


def vtln():
    n, speakers, utts, T = 60, 109, 4, 1301
    rows = speakers * utts * T
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((rows, n), generator=g, device=dev)
    alpha_s = (torch.rand(speakers, generator=g, device=dev) * 0.4 - 0.2)
    alpha = alpha_s.repeat_interleave(utts * T).contiguous()
    gy = torch.randn((rows, n), generator=g, device=dev)
    ms_f = timed(lambda: ops.allpass_forward(x, alpha, n))
    ms_b = timed(lambda: ops.allpass_backward(gy, x, alpha, n))
    bytes_f, bytes_b = rows * (8 * n + 4), rows * (12 * n + 8)
    print(json.dumps({"config": "Neural-VTLN all-pass warp, 109 speakers x 4 utts x 1301 frames, n = 60, one alpha per speaker",
                      "rows": rows, "fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_GBs": bytes_f / ms_f / 1e6, "bwd_GBs": bytes_b / ms_b / 1e6,
                      "fwd_frac_of_hbm_peak": bytes_f / ms_f / 1e6 / PEAK, "bwd_frac_of_hbm_peak": bytes_b / ms_b / 1e6 / PEAK,
                      "hbm_peak_GBs": PEAK}), flush=True)


if __name__ == "__main__" and "batch_only" not in sys.argv:
    synthesis()
    vtln()

# ---- trainer-facing batch: ragged rows -> padded normalised [T_max, B, 64] and back, HBM-bound --------------
try:
    rng = np.random.default_rng(9)
    lens = rng.integers(600, 1302, size=2048)
    F = int(lens.sum())
    feats_b = torch.randn((F, 64), device=dev)
    off_b = torch.from_numpy(np.concatenate(([0], np.cumsum(lens))).astype(np.int64)).to(dev)
    fu_b = torch.from_numpy(np.repeat(np.arange(len(lens), dtype=np.int32), lens)).to(dev)
    mean_b, std_b = torch.randn(64, device=dev), torch.rand(64, device=dev) + 0.5
    ms_p = timed(lambda: ops.pad_normalise(feats_b, off_b, mean_b, std_b, lengths=lens), steps=20)
    padded_b, _, _ = ops.pad_normalise(feats_b, off_b, mean_b, std_b)
    ms_u = timed(lambda: ops.unpad_denormalise(padded_b, off_b, fu_b, mean_b, std_b), steps=20)
    bytes_p = F * 64 * 4 + padded_b.numel() * 4 + padded_b.shape[0] * padded_b.shape[1] * 4
    bytes_u = 2 * F * 64 * 4
    print(json.dumps({"config": "trainer batch: 2048 utterances (600-1301 frames) x 64 features, normalise + pad / un-pad + de-normalise",
                      "pad_ms": ms_p, "unpad_ms": ms_u, "pad_GBs": bytes_p / ms_p / 1e6, "unpad_GBs": bytes_u / ms_u / 1e6,
                      "pad_frac_of_hbm_peak": bytes_p / ms_p / 1e6 / PEAK, "unpad_frac_of_hbm_peak": bytes_u / ms_u / 1e6 / PEAK}))
except Exception as e:  # noqa: BLE001
    print("trainer batch bench failed:", e)
