"""Development diagnostic: runs every kernel against the CPU oracle on one golden utterance and prints error stats.
Not part of the product; `pytest -m gpu` holds the asserting versions."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops  # noqa: E402
from idiaptts_b200.compat import pyworld as pw, pysptk as ps  # noqa: E402
from oracle import glue_np, sptk_np, world_np  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ljspeech_world_golden.npz"))
ID = sys.argv[1] if len(sys.argv) > 1 else "LJ001-0008"
x = g[ID + "/wav"].astype(np.float64) / 32768.0
x = np.append(x[0], x[1:] - 0.97 * x[:-1])
c = g[ID + "/cmp"]
T = c.shape[0]
fs = 16000
f0 = np.where(c[:, 63] > 0, np.exp(c[:, 60].astype(np.float64)), 0.0)
t = world_np.temporal_positions(T)
dev = torch.device("cuda")
print("device", torch.cuda.get_device_name(0), "T", T, flush=True)


def stage(name):
    def deco(fn):
        t0 = time.time()
        try:
            fn()
            print("[ok  ] %-14s %.2fs" % (name, time.time() - t0), flush=True)
        except Exception:
            print("[FAIL] %-14s" % name, flush=True)
            traceback.print_exc()
        return fn
    return deco


state = {}


@stage("cheaptrick")
def _():
    ref = world_np.cheaptrick(x, f0, t, fs)
    out = pw.cheaptrick(x, f0, t, fs)
    rel = np.abs(out - ref) / ref
    print("   sp rel err max %.3e  median %.3e  nan %d" % (np.nanmax(rel), np.nanmedian(rel), np.isnan(out).sum()))
    worst = np.unravel_index(np.nanargmax(rel), rel.shape)
    print("   worst frame/bin", worst, "f0", f0[worst[0]], "ref", ref[worst], "out", out[worst])
    state["sp_ref"], state["sp"] = ref, out
    # int16 + in-kernel pre-emphasis path, float32 output
    raw16 = g[ID + "/wav"]
    b = ops.RaggedBatch.from_host([raw16], [f0], fs, preemphasis=0.97)
    sp32, st = ops.cheaptrick(b, out_dtype=torch.float32)
    rel32 = np.abs(sp32.cpu().numpy() - ref) / ref
    print("   int16+preemph f32-out rel err max %.3e status %d" % (rel32.max(), int(st.item())))


@stage("d4c")
def _():
    voiced_ref, coarse_ref = world_np.d4c_coarse(x, f0, t, fs)
    b = ops.RaggedBatch.from_host([x], [f0], fs)
    coarse, voiced, st = ops.d4c_coarse(b)
    coarse, voiced = coarse.cpu().numpy(), voiced.cpu().numpy().astype(bool)
    print("   status", int(st.item()), "voiced ref/out", voiced_ref.sum(), voiced.sum(), "mismatch", (voiced != voiced_ref).sum())
    both = voiced & voiced_ref
    d = np.abs(coarse[both] - coarse_ref[both])
    print("   coarse dB abs err max %.3e median %.3e nan %d" % (np.nanmax(d), np.nanmedian(d), np.isnan(coarse[both]).sum()))
    bap = ops.bap_from_coarse(torch.from_numpy(coarse).to(dev), torch.from_numpy(voiced.astype(np.uint8)).to(dev), fs, 1024).cpu().numpy()
    print("   bap vs golden cmp max abs %.3e" % np.abs(bap[:, 0] - c[:, 64]).max())
    ap = pw.d4c(x, f0, t, fs)
    ap_ref = world_np.d4c(x, f0, t, fs)
    print("   ap plane max abs err %.3e" % np.abs(ap - ap_ref).max())
    bap2 = pw.code_aperiodicity(ap, fs)
    print("   code_aperiodicity vs oracle %.3e" % np.abs(bap2 - world_np.code_aperiodicity(ap_ref, fs)).max())
    dec = pw.decode_aperiodicity(bap2, fs, 1024)
    print("   decode_aperiodicity vs oracle %.3e" % np.abs(dec - world_np.decode_aperiodicity(bap2, fs, 1024)).max())


@stage("mcep")
def _():
    sp = state.get("sp_ref")
    if sp is None:
        sp = world_np.cheaptrick(x, f0, t, fs)
    amp = np.sqrt(sp)
    for order, alpha in ((19, 0.58), (59, 0.58), (59, 0.41)):
        ref = [sptk_np.mcep_frame(a, order, alpha, eps=1e-8) for a in amp]
        rm = np.stack([r[0] for r in ref])
        rit = np.array([r[1] for r in ref])
        plane = torch.from_numpy(amp).to(dev)
        iters = torch.zeros(T, dtype=torch.int32, device=dev)
        mc, st = ops.mcep(plane, order, alpha, out_dtype=torch.float64, iters=iters)
        mc = mc.cpu().numpy()
        print("   order %d alpha %.2f: max|d| %.3e MCD %.3e dB iters equal %.3f (gpu %d..%d) status %d nan %d" % (
            order, alpha, np.nanmax(np.abs(mc - rm)), glue_np.mcd_db(rm, mc), np.mean(iters.cpu().numpy() == rit),
            iters.min().item(), iters.max().item(), int(st.item()), np.isnan(mc).sum()))
        if order == 19:
            print("   vs golden cmp: max|d| %.3e" % np.abs(mc.astype(np.float32) - c[:, :20]).max())
    mc60 = torch.from_numpy(rm).to(dev)
    rec = ops.mc2sp(mc60, 0.41, 1024, scale=1.0, do_exp=True, out_dtype=torch.float64).cpu().numpy()
    rec_ref = np.exp(sptk_np.mgc2sp(rm, 0.41, 0.0, 1024).real)
    print("   mc2sp rel err max %.3e" % (np.abs(rec - rec_ref) / rec_ref).max())


@stage("labels")
def _():
    f0d = torch.from_numpy(f0).to(dev)
    off = torch.tensor([0, T], dtype=torch.int64, device=dev)
    lf0, vuv = ops.lf0_vuv(f0d, off)
    lf0_ref, vuv_ref = glue_np.interpolate_lin(glue_np.lf0_from_f0(f0))
    print("   vuv exact:", np.array_equal(vuv.cpu().numpy(), vuv_ref.astype(np.float32)), " lf0 max abs %.3e" % np.abs(lf0.cpu().numpy() - lf0_ref).max())
    feats = torch.from_numpy(np.ascontiguousarray(c[:, :20])).to(dev)
    d, dd = ops.deltas(feats, off)
    print("   deltas exact:", np.array_equal(d.cpu().numpy(), c[:, 20:40]), np.array_equal(dd.cpu().numpy(), c[:, 40:60]))
    sums = torch.zeros(40, dtype=torch.float64, device=dev)
    gram = torch.zeros(400, dtype=torch.float64, device=dev)
    ops.stats_accumulate(feats, sums, gram)
    c64 = c[:, :20].astype(np.float64)
    print("   stats err %.3e %.3e gram %.3e" % (np.abs(sums.cpu().numpy()[:20] - c64.sum(0)).max(), np.abs(sums.cpu().numpy()[20:] - (c64 ** 2).sum(0)).max(),
                                               np.abs(gram.cpu().numpy().reshape(20, 20) - c64.T @ c64).max()))


@stage("synthesis")
def _():
    sp = state.get("sp_ref")
    if sp is None:
        sp = world_np.cheaptrick(x, f0, t, fs)
    ap = world_np.d4c(x, f0, t, fs)
    tab = ops.randn_table(5000, dev).cpu().numpy()
    print("   randn table err %.3e" % np.abs(tab[:5000] - world_np.xorshift_randn_sequence(5000)).max())
    y_ref = world_np.synthesize(f0, sp, ap, fs)
    y = pw.synthesize(f0, sp, ap, fs)
    err = y - y_ref
    snr = 10 * np.log10((y_ref ** 2).sum() / max((err ** 2).sum(), 1e-300))
    print("   len", len(y), len(y_ref), "SNR %.1f dB max abs err %.3e peak %.3f nan %d" % (snr, np.abs(err).max(), np.abs(y_ref).max(), np.isnan(y).sum()))
    idx, shift, ivuv = world_np.synthesis_time_base(f0, fs, 0.005, len(y_ref), 1024)
    print("   oracle pulses", len(idx))


@stage("vtln")
def _():
    rng = np.random.default_rng(0)
    for n in (30, 60):
        rows, blocks = 300, 3
        xx = rng.standard_normal((rows, blocks * n)).astype(np.float32)
        al = rng.uniform(-0.2, 0.2, rows).astype(np.float32)
        ref = glue_np.allpass_warp_forward(xx.astype(np.float64), al.astype(np.float64), n)
        xd, ad = torch.from_numpy(xx).to(dev), torch.from_numpy(al).to(dev)
        y = ops.allpass_forward(xd, ad, n).cpu().numpy()
        print("   n=%d fwd max abs err %.3e (ref max %.2f)" % (n, np.abs(y - ref).max(), np.abs(ref).max()))
        gy = rng.standard_normal((rows, blocks * n)).astype(np.float32)
        gx, ga = ops.allpass_backward(torch.from_numpy(gy).to(dev), xd, ad, n)
        # fp64 reference gradients: gx = W-applied, galpha by central differences of the oracle
        gx_ref = np.zeros_like(ref)
        ga_ref = np.zeros(rows)
        eps = 1e-6
        for r in range(0, rows, 37):
            A = sptk_np.freqt_matrix(n - 1, n - 1, float(al[r]))
            S1 = np.eye(n); S1[0, 0] = 0.5
            S2 = np.eye(n); S2[0, 0] = 2.0
            J = S2 @ A @ S1
            for b in range(blocks):
                gx_ref[r, b * n:(b + 1) * n] = J.T @ gy[r, b * n:(b + 1) * n].astype(np.float64)
            yp = glue_np.allpass_warp_forward(xx[r:r + 1].astype(np.float64), np.array([al[r] + eps]), n)
            ym = glue_np.allpass_warp_forward(xx[r:r + 1].astype(np.float64), np.array([al[r] - eps]), n)
            ga_ref[r] = ((yp - ym) / (2 * eps) * gy[r]).sum()
        sel = np.arange(0, rows, 37)
        print("   n=%d bwd gx max abs err %.3e  galpha max rel err %.3e" % (
            n, np.abs(gx.cpu().numpy()[sel] - gx_ref[sel]).max(), (np.abs(ga.cpu().numpy()[sel] - ga_ref[sel]) / np.abs(ga_ref[sel]).max()).max()))


torch.cuda.synchronize()
print("done", flush=True)
