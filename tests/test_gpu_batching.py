"""GPU parity of the trainer-facing batch kernels: b2w_pad_normalise / b2w_unpad_denormalise against the numpy
restatement of WorldFeatLabelGen.preprocess_sample + ModularModelHandlerPyTorch.prepare_batch (oracle/glue_np.py); float32
arithmetic, so the comparison is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import glue_np

pytestmark = pytest.mark.gpu


def _ragged(rng, lengths, W):
    samples = [rng.standard_normal((n, W)).astype(np.float32) * 3 + 1 for n in lengths]
    feats = torch.from_numpy(np.concatenate(samples)).cuda()
    off = torch.from_numpy(np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)).cuda()
    fu = torch.from_numpy(np.repeat(np.arange(len(lengths), dtype=np.int32), lengths)).cuda()
    return samples, feats, off, fu


@pytest.mark.parametrize("batch_first", [False, True])
@pytest.mark.parametrize("W", [64, 190, 5])
def test_pad_normalise_matches_reference_protocol(batch_first, W):
    from idiaptts_b200 import ops
    rng = np.random.default_rng(W + batch_first)
    lengths = [37, 1301, 1, 640, 0, 12]
    samples, feats, off, fu = _ragged(rng, lengths, W)
    mean = rng.standard_normal(W).astype(np.float32)
    std = (rng.random(W).astype(np.float32) + 0.5)
    out, mask, lens = ops.pad_normalise(feats, off, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), batch_first=batch_first)
    ref, ref_mask, ref_lens = glue_np.prepare_batch(samples, mean, std, batch_first=batch_first)
    assert np.array_equal(lens, ref_lens)
    assert np.array_equal(out.cpu().numpy(), ref)          # bit-exact float32
    assert np.array_equal(mask.cpu().numpy(), ref_mask)
    # without normalisation parameters and with min_frames
    out2, mask2, _ = ops.pad_normalise(feats, off, batch_first=batch_first, min_frames=1400)
    ref2, ref_mask2, _ = glue_np.prepare_batch(samples, batch_first=batch_first, min_frames=1400)
    assert out2.shape == ref2.shape and np.array_equal(out2.cpu().numpy(), ref2) and np.array_equal(mask2.cpu().numpy(), ref_mask2)
    # inverse: network output -> ragged de-normalised rows
    back = ops.unpad_denormalise(out, off, fu, torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda(), batch_first=batch_first)
    ref_back = glue_np.unprepare_batch(ref, ref_lens, mean, std, batch_first=batch_first)
    assert np.array_equal(back.cpu().numpy(), np.concatenate(ref_back))
    np.testing.assert_allclose(back.cpu().numpy(), np.concatenate(samples), rtol=1e-5, atol=1e-5)


def test_pad_normalise_strided_rows_and_corpus_scale():
    """Feature rows with a row stride (a column block of the packed [F, 64] plane) and a round trip at the synthesis benchmark's
    size (256 utterances x 1301 frames): un-pad(pad(x)) == x exactly when no normalisation is applied."""
    from idiaptts_b200 import ops
    rng = np.random.default_rng(3)
    lengths = rng.integers(200, 1302, size=256)
    samples, feats, off, fu = _ragged(rng, lengths, 64)
    out, mask, _ = ops.pad_normalise(feats, off)
    assert out.shape == (int(lengths.max()), 256, 64) and float(mask.sum()) == float(lengths.sum())
    assert torch.equal(ops.unpad_denormalise(out, off, fu), feats)
    sub = feats[:, :60]   # mcep columns only: row stride 64
    out60, _, _ = ops.pad_normalise(sub, off, batch_first=True)
    ref60, _, _ = glue_np.prepare_batch([s[:, :60] for s in samples], batch_first=True)
    assert np.array_equal(out60.cpu().numpy(), ref60)


@pytest.mark.parametrize("add_deltas", [False, True])
def test_prefetching_batch_loader_equals_the_per_sample_reader(tmp_path, add_deltas):
    """WorldFeatLabelGen.batches (background load_batch into pinned memory + pad / normalise on the device) yields exactly what the
    reference protocol gives sample by sample: reader[id] (load + preprocess_sample) padded by prepare_batch -- bit for bit."""
    from test_host_logic import _write_world_label_dir
    from idiaptts_b200.WorldFeatLabelGen import WorldFeatLabelGen
    rng = np.random.default_rng(5)
    ids = ["u%s" % ("x" * k) for k in range(11)]              # 11 utterances of different lengths
    D, nap = 6, 2
    root = str(tmp_path / "labels")
    _write_world_label_dir(root, rng, ids, D, nap, add_deltas)
    reader = WorldFeatLabelGen(root, add_deltas=add_deltas, num_coded_sps=D, num_bap=nap)
    with pytest.raises(RuntimeError, match="get_normalisation_params"):
        next(reader.batches(ids, 4))
    reader.get_normalisation_params(root, "train")
    for batch_first, shuffle in ((False, False), (True, True)):
        seen = []
        for b in reader.batches(ids, 4, shuffle=shuffle, seed=3, batch_first=batch_first, prefetch=2):
            samples = [reader[i] for i in b["ids"]]           # the reference protocol, one sample at a time
            ref, ref_mask, ref_lens = glue_np.prepare_batch(samples, batch_first=batch_first)
            assert np.array_equal(b["lengths"], ref_lens)
            assert np.array_equal(b["padded"].cpu().numpy(), ref) and np.array_equal(b["mask"].cpu().numpy(), ref_mask)
            assert b["frame_off"].cpu().tolist() == np.concatenate(([0], np.cumsum(ref_lens))).tolist()
            seen += b["ids"]
        assert sorted(seen) == sorted(ids) and (seen != ids) == shuffle
    assert [len(b["ids"]) for b in reader.batches(ids, 4, drop_last=True)] == [4, 4]
    # abandoning the iterator stops the background thread; a missing archive surfaces in the consumer
    it = reader.batches(ids, 2, prefetch=1)
    next(it)
    it.close()
    with pytest.raises(FileNotFoundError):
        list(reader.batches(ids[:3] + ["missing"], 2))
