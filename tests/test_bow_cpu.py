"""K9's oracle (oracle/bow_oracle.py) against the fixtures generated from the unmodified reference
(tests/golden/bow_golden.npz <- make_bow_golden.py: persistent_map.compute_bow_histogram,
BoWDatabase._compute_hist / rank_candidates, sklearn cosine_similarity), and the bridge's host
form used for non-ORB descriptors."""
import numpy as np

from oracle import bow_oracle as bo


def _cases(golden_dir):
    g = np.load(golden_dir / "bow_golden.npz")
    for name in g["names"]:
        name = str(name)
        frames = [g[f"{name}.desc{i}"] for i in range(int(g[f"{name}.n_frames"]))]
        yield name, g, g[f"{name}.vocab"], frames


def test_oracle_histograms_equal_the_reference(golden_dir):
    n = 0
    for name, g, vocab, frames in _cases(golden_dir):
        hists = np.vstack([bo.compute_bow_histogram(f, vocab) for f in frames])
        assert hists.dtype == np.float32
        assert np.array_equal(hists, g[f"{name}.hists"]), name          # bit-exact, ties and empty frames included
        n += len(frames)
    assert n >= 20


def test_oracle_scores_and_ranking_equal_the_reference(golden_dir):
    for name, g, vocab, frames in _cases(golden_dir):
        hists = g[f"{name}.hists"]
        s = bo.cosine_scores(hists[0], hists)
        np.testing.assert_allclose(s, g[f"{name}.scores0"], atol=2e-7, rtol=0)   # sklearn accumulates in float32
        ids = np.arange(100, 100 + len(frames))
        order = bo.rank(g[f"{name}.scores0"], ids)
        assert [int(ids[i]) for i in order] == [int(x) for x in g[f"{name}.rank_ids"]], name


def test_bridge_host_form_equals_the_reference(golden_dir):
    from integration.relocalization_bridge import host_bow_histogram, host_bow_scores
    for name, g, vocab, frames in _cases(golden_dir):
        for i, f in enumerate(frames):
            assert np.array_equal(host_bow_histogram(f, vocab), g[f"{name}.hists"][i]), name
        np.testing.assert_allclose(host_bow_scores(frames[0], vocab, g[f"{name}.hists"]), g[f"{name}.scores0"], atol=1e-7)
    # float32 / L2 descriptors keep the reference's host arithmetic and its errors
    rng = np.random.default_rng(0)
    import pytest
    with pytest.raises(ValueError):
        host_bow_histogram(rng.normal(size=(5, 16)).astype(np.float32), rng.normal(size=(4, 32)).astype(np.float32))
    assert host_bow_histogram(np.zeros((0, 32), np.uint8), np.ones((4, 32), np.float32)).tolist() == [0, 0, 0, 0]
