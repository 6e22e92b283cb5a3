"""CPU oracle — bag-of-words histogram and cosine ranking.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of /root/reference/persistent_map.py:82-96 (compute_bow_histogram; the same
arithmetic as BoWDatabase._compute_hist, /root/reference/loop_closure.py:36-48) and of the
ranking in MapRelocalizer.relocalize (persistent_map.py:234-243) / BoWDatabase.rank_candidates
(loop_closure.py:56-74).  The nearest-centroid search of the reference lives in scikit-learn
(pairwise_distances_argmin_min; un-pinned dependency, requirements.txt:7, 1.9.0 in this image):
squared Euclidean distances through ||x||^2 - 2 x.y + ||y||^2 accumulated in float64, first index
on ties.  Pinned by tests/golden/bow_golden.npz (make_bow_golden.py runs the unmodified
reference).  Nothing in the product imports this module.
"""
from __future__ import annotations

import numpy as np


def bow_words(descriptors, vocab) -> np.ndarray:
    """Word (nearest centroid, lowest index on ties) of every descriptor; float64."""
    x = np.asarray(descriptors).astype(np.float64)
    v = np.asarray(vocab, dtype=np.float32).astype(np.float64)
    d2 = (v * v).sum(1)[None, :] - 2.0 * (x @ v.T)
    return np.argmin(d2, axis=1).astype(np.int32)


def compute_bow_histogram(descriptors, vocab) -> np.ndarray:
    """persistent_map.py:82-96."""
    vocab = np.asarray(vocab)
    if descriptors is None or len(descriptors) == 0:
        return np.zeros(vocab.shape[0], dtype=np.float32)
    words = bow_words(descriptors, vocab)
    hist = np.bincount(words, minlength=vocab.shape[0]).astype(np.float32)
    if hist.sum() > 0:
        hist /= hist.sum()
    return hist


def cosine_scores(hist, hists) -> np.ndarray:
    """cosine_similarity([hist], hists)[0] in float64, rounded to float32; zero rows score 0."""
    a = np.asarray(hist, dtype=np.float64)
    B = np.asarray(hists, dtype=np.float64)
    na, nb = np.sqrt((a * a).sum()), np.sqrt((B * B).sum(1))
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (B @ a) / (na * nb)
    return np.where((na > 0) & (nb > 0), s, 0.0).astype(np.float32)


def rank(scores, frame_ids, top_k=None):
    """sorted by (-score, frame_id) (persistent_map.py:236-242, loop_closure.py:68)."""
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), int(frame_ids[i])))
    return order if top_k is None else order[:top_k]
