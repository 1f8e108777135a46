"""Where does the end-to-end overhead of the corpus-level synthesis come from?  Times WorldSynthesizer.synthesize_corpus over the same
features in four modes: resident, features from pinned host memory, waveforms to pinned host memory, both."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from idiaptts_b200 import ops, pipeline, synthetic
dev = torch.device("cuda", 0)
FS, U = 22050, int(os.environ.get("UTTS", "2048"))
waves, f0s = synthetic.make_corpus(U, FS, seed=2, mean_dur=6.5, std_dur=1.8, dur_quantum=0.1, device=dev)
batch = ops.RaggedBatch.from_host([w.cpu().numpy() for w in waves], f0s, FS, device=dev)
an = pipeline.WorldAnalyzer(FS, 60, device=dev)
feats, _, _ = an.extract(batch)
fo = batch.frame_off.cpu().numpy()
syn = pipeline.WorldSynthesizer(FS, 60, device=dev)
ylen = int((np.diff(fo) * 5.0 * FS / 1000).astype(np.int64).sum())
feats_host = feats.cpu().pin_memory()
y_host = torch.empty(ylen, dtype=torch.float32).pin_memory()
y_dev = torch.empty(ylen, dtype=torch.float32, device=dev)
audio = ylen / FS
def run(mode):
    kw = {}
    f = feats
    if mode in ("h2d", "both"): kw["feats_host"] = feats_host; f = None
    if mode in ("d2h", "both"): kw["out_host"] = y_host
    else: kw["out"] = y_dev
    for _ in range(2): syn.synthesize_corpus(f, fo, batch_utts=256, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): syn.synthesize_corpus(f, fo, batch_utts=256, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-8s %8.2f ms  %9.0f audio-s/s" % (mode, ms, audio / ms * 1e3), flush=True)
for m in ("resident", "h2d", "d2h", "both", "resident"):
    run(m)
