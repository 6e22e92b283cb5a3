#!/usr/bin/env python
"""Generate tests/golden/bow_golden.npz by running the UNMODIFIED reference
(/root/reference/persistent_map.py: compute_bow_histogram; /root/reference/loop_closure.py:
BoWDatabase._compute_hist / rank_candidates; sklearn cosine_similarity as MapRelocalizer calls
it, persistent_map.py:235).  Build container only.

Cases: random byte descriptors, real cv2.ORB descriptors, a vocabulary with duplicated
centroids (exact ties), descriptors equal to centroids, an empty frame; vocabulary sizes 8..500."""
import importlib.machinery
import importlib.util
import sys
from pathlib import Path

import cv2
import numpy as np
from sklearn.metrics.pairwise import cosine_similarity

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))                      # persistent_map imports homography
sys.path.insert(1, str(OUT.parents[1]))           # ... and, through it, nothing of this repo


def _load(name):
    loader = importlib.machinery.SourceFileLoader("_ref_" + name, str(REF / (name + ".py")))
    spec = importlib.util.spec_from_loader("_ref_" + name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_" + name] = mod
    loader.exec_module(mod)
    return mod


ref_pm = _load("persistent_map")
ref_lc = _load("loop_closure")


def orb_frames(n_frames, seed):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (376, 1241), dtype=np.uint8), (5, 5), 0)
    orb = cv2.ORB_create(600)
    out = []
    for f in range(n_frames):
        img = np.roll(base, (f, 3 * f), axis=(0, 1))
        cv2.setRNGSeed(1337)
        _, d = orb.detectAndCompute(img, None)
        out.append(d)
    return out


def main():
    out, names = {}, []
    rng = np.random.default_rng(7)
    cases = []
    for name, k, n_frames, n in (("rand_k8", 8, 3, 50), ("rand_k64", 64, 6, 400), ("rand_k500", 500, 4, 2000)):
        vocab = rng.uniform(0, 255, (k, 32)).astype(np.float32)
        frames = [rng.integers(0, 256, (n - 7 * f, 32), dtype=np.uint8) for f in range(n_frames)]
        cases.append((name, vocab, frames))
    orb = orb_frames(5, 3)
    stacked = np.vstack(orb).astype(np.float32)
    vocab = stacked[rng.permutation(len(stacked))[:100]] + rng.normal(0, 4.0, (100, 32)).astype(np.float32)
    cases.append(("orb_k100", vocab.astype(np.float32), orb))
    # exact ties: duplicated centroids, descriptors equal to centroids, integer-valued centroids
    vocab = rng.integers(0, 256, (16, 32)).astype(np.float32)
    vocab[9] = vocab[2]
    vocab[13] = vocab[5]
    frames = [np.vstack([vocab[[2, 5, 9, 13, 0]].astype(np.uint8), rng.integers(0, 256, (40, 32), dtype=np.uint8)]),
              rng.integers(0, 4, (30, 32), dtype=np.uint8)]
    cases.append(("ties_k16", vocab, frames))
    cases.append(("with_empty", rng.uniform(0, 255, (12, 32)).astype(np.float32),
                  [rng.integers(0, 256, (20, 32), dtype=np.uint8), np.zeros((0, 32), np.uint8), rng.integers(0, 256, (1, 32), dtype=np.uint8)]))
    for name, vocab, frames in cases:
        names.append(name)
        hists = np.vstack([ref_pm.compute_bow_histogram(f, vocab) for f in frames])
        db = ref_lc.BoWDatabase(vocab_size=vocab.shape[0])
        db.vocab, db.vocab_trained = vocab, True
        h2 = np.vstack([db._compute_hist(f) if len(f) else np.zeros(vocab.shape[0], np.float32) for f in frames])
        assert np.array_equal(hists, h2), name          # the two reference call sites agree
        out[f"{name}.vocab"] = vocab
        out[f"{name}.n_frames"] = np.int64(len(frames))
        for i, f in enumerate(frames):
            out[f"{name}.desc{i}"] = f
        out[f"{name}.hists"] = hists
        # ranking of frame 0 against all frames (persistent_map.py:235)
        out[f"{name}.scores0"] = cosine_similarity([hists[0]], hists)[0].astype(np.float32)
        db.hists = [h for h in hists]
        db.frame_ids = list(range(100, 100 + len(frames)))
        ranked = db.rank_candidates(frames[0])
        out[f"{name}.rank_ids"] = np.array([r[0] for r in ranked], np.int64)
        out[f"{name}.rank_scores"] = np.array([r[1] for r in ranked], np.float64)
    out["names"] = np.array(names)
    np.savez_compressed(OUT / "bow_golden.npz", **out)
    print("wrote", OUT / "bow_golden.npz", "cases", names)


if __name__ == "__main__":
    main()
