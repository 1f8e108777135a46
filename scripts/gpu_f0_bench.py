"""F0-stage timing on one B200 (scripts/, not the product): DIO + StoneMask over an LJSpeech-shaped slice."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from idiaptts_b200 import ops, synthetic

fs, U = 22050, int(sys.argv[1]) if len(sys.argv) > 1 else 512
waves, f0s = synthetic.make_corpus(U, fs, seed=3, mean_dur=6.5, device="cpu")
waves = [w.numpy() for w in waves]
batch = ops.RaggedBatch.from_host(waves, f0s, fs, device="cuda")
audio_s = sum(len(w) for w in waves) / fs
res = {"utts": U, "audio_s": audio_s, "frames": batch.num_frames}
for name, fn in (("dio", lambda: ops.dio(batch)), ("stonemask", None)):
    if name == "stonemask":
        f0d = ops.dio(batch)
        fn = lambda: ops.stonemask(batch, f0d)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[name + "_ms"] = ms
    res[name + "_audio_s_per_s"] = audio_s / ms * 1e3
f0 = ops.stonemask(batch, ops.dio(batch)).cpu().numpy()
ref = np.concatenate(f0s)
res["voiced_frac_estimated"] = float((f0 > 0).mean())
res["voiced_frac_truth"] = float((ref > 0).mean())
both = (f0 > 0) & (ref > 0)
res["median_rel_err_vs_generator_f0"] = float(np.median(np.abs(f0[both] / ref[both] - 1)))
res["f0_stage_audio_s_per_s"] = audio_s / (res["dio_ms"] + res["stonemask_ms"]) * 1e3
print(json.dumps(res))
