"""CPU oracle — Hamming kNN-2 / cross-check / ratio matching.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of the matching half of the hot path.  Every function cites
the reference lines (relative to /root/reference) or the third-party behaviour
(OpenCV 4.13 ``cv::BFMatcher``) it follows.  Nothing in the product imports this
module (see oracle/__init__.py).

Packed keys
-----------
``key = (distance << 22) | index`` (uint32, distance <= 256, index < 2**22; any key >= 0x80000000 is "none").
``min`` over keys == the lexicographic ``(distance, index)`` minimum, which is
OpenCV's tie rule (lowest train index wins a tie).
``NONE_KEY = 0xFFFFFFFF`` marks "no such neighbour" (e.g. second-best when the
train set has one row).
"""

from __future__ import annotations

import math

import numpy as np

IDX_BITS = 22
IDX_MASK = (1 << IDX_BITS) - 1
NONE_KEY = np.uint32(0xFFFFFFFF)


# --------------------------------------------------------------------------- #
# distance matrix
# --------------------------------------------------------------------------- #

def _as_desc(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("descriptors must be (N, W) uint8")
    return np.ascontiguousarray(a)


def hamming_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """(Nq, Nt) int32 Hamming distances between byte rows.

    Follows ``popcount_uint8(np.bitwise_xor(d1, desc2)).sum(axis=1)``
    (homography.py:14, :105-108) == ``cv::hal::normHamming`` used by
    ``cv::batchDistance`` behind ``BFMatcher(NORM_HAMMING)``
    (feature_pipeline.py.bak:68).  Chunked over queries to bound memory.
    """
    q = _as_desc(q)
    t = _as_desc(t)
    if q.shape[1] != t.shape[1]:
        raise ValueError("descriptor widths differ")
    nq, w = q.shape
    nt = t.shape[0]
    out = np.empty((nq, nt), dtype=np.int32)
    if w % 8 == 0:
        qv, tv = q.view(np.uint64), t.view(np.uint64)
    else:
        qv, tv = q, t
    lanes = qv.shape[1]
    chunk = max(1, (1 << 24) // max(1, nt * lanes))
    for s in range(0, nq, chunk):
        x = np.bitwise_xor(qv[s:s + chunk, None, :], tv[None, :, :])
        out[s:s + chunk] = np.bitwise_count(x).sum(axis=2, dtype=np.int32)
    return out


# --------------------------------------------------------------------------- #
# kNN-2, column minimum, packed keys
# --------------------------------------------------------------------------- #

def packed_keys(q: np.ndarray, t: np.ndarray):
    """(fwd_best, fwd_second, bwd_best) packed uint32 keys for one pair.

    fwd_* index the train set per query row; bwd_best indexes the query set per
    train row.  This is the exact contract of the CUDA kernel's outputs
    (include/b2s.h: b2s_hamming_knn2_batched).
    """
    D = hamming_matrix(q, t)
    nq, nt = D.shape
    if nq >= (1 << IDX_BITS) or nt >= (1 << IDX_BITS):
        raise ValueError("too many rows for 22-bit indices")
    kf = (D.astype(np.uint32) << IDX_BITS) | np.arange(nt, dtype=np.uint32)[None, :]
    kb = (D.astype(np.uint32) << IDX_BITS) | np.arange(nq, dtype=np.uint32)[:, None]
    if nt == 0:
        fwd_best = np.full(nq, NONE_KEY, np.uint32)
        fwd_second = np.full(nq, NONE_KEY, np.uint32)
    elif nt == 1:
        fwd_best = kf[:, 0].copy()
        fwd_second = np.full(nq, NONE_KEY, np.uint32)
    else:
        part = np.partition(kf, 1, axis=1)[:, :2]
        fwd_best = part.min(axis=1)
        fwd_second = part.max(axis=1)
    bwd_best = kb.min(axis=0) if nq > 0 else np.full(nt, NONE_KEY, np.uint32)
    return fwd_best, fwd_second, bwd_best.astype(np.uint32)


def key_index(k: np.ndarray) -> np.ndarray:
    return (np.asarray(k, dtype=np.uint32) & np.uint32(IDX_MASK)).astype(np.int64)


def key_distance(k: np.ndarray) -> np.ndarray:
    return (np.asarray(k, dtype=np.uint32) >> np.uint32(IDX_BITS)).astype(np.int64)


def knn2(q: np.ndarray, t: np.ndarray):
    """``cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2)`` semantics.

    Returns ``(idx (Nq,2) int64, dist (Nq,2) int64, k_found int)``; columns past
    ``k_found = min(2, Nt)`` are -1.  Per query: the two smallest
    ``(distance, trainIdx)`` in lexicographic order (call site
    feature_pipeline.py.bak:84).
    """
    b, s, _ = packed_keys(q, t)
    nq = b.shape[0]
    nt = np.asarray(t).shape[0]
    idx = np.full((nq, 2), -1, np.int64)
    dist = np.full((nq, 2), -1, np.int64)
    kf = min(2, nt)
    if kf >= 1:
        idx[:, 0], dist[:, 0] = key_index(b), key_distance(b)
    if kf >= 2:
        idx[:, 1], dist[:, 1] = key_index(s), key_distance(s)
    return idx, dist, kf


def cross_check_match(q: np.ndarray, t: np.ndarray):
    """``cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(q, t)`` semantics.

    ``f[i] = argmin_j D[i,j]`` (lowest j on ties), ``b[j] = argmin_i D[i,j]``
    (lowest i on ties); keep ``(i, f[i], D[i,f[i]])`` iff ``b[f[i]] == i``;
    ascending i (call sites feature_pipeline.py.bak:82,
    persistent_map.py:266, keyframe_manager.py:126,141).
    """
    b, _, bw = packed_keys(q, t)
    if b.size == 0 or bw.size == 0:
        z = np.zeros(0, np.int64)
        return z, z.copy(), z.copy()
    f = key_index(b)
    keep = key_index(bw)[f] == np.arange(b.shape[0])
    qi = np.flatnonzero(keep)
    return qi, f[qi], key_distance(b)[qi]


def ratio_lut(ratio: float) -> np.ndarray:
    """257-entry table: keep iff ``d1 < lut[d2]``.

    The reference compares Python doubles ``m.distance < ratio * n.distance``
    (feature_pipeline.py.bak:90; homography.py:16) with integer-valued
    distances, so the strict test equals ``d1 < ceil(ratio * d2)`` with the
    product evaluated in float64.
    """
    return np.array([math.ceil(float(ratio) * float(d2)) for d2 in range(257)], dtype=np.int32)


def pipeline_match(q, t, cross_check: bool = True, ratio_test: float = 0.8,
                   max_matches: int | None = 500):
    """``ORBFeaturePipeline.match`` (feature_pipeline.py.bak:78-95).

    Returns ``(queryIdx, trainIdx, distance)`` int64 arrays in the reference's
    output order: stable sort by distance (ties keep ascending queryIdx), then
    truncation to ``max_matches``.
    """
    z = np.zeros(0, np.int64)
    if q is None or t is None or len(q) == 0 or len(t) == 0:       # .bak:79-80
        return z, z.copy(), z.copy()
    if cross_check:                                                 # .bak:81-82
        qi, ti, d = cross_check_match(q, t)
    else:                                                           # .bak:83-91
        idx, dist, kf = knn2(q, t)
        if kf < 2:                                                  # len(pair) < 2 -> skipped
            return z, z.copy(), z.copy()
        lut = ratio_lut(ratio_test)
        keep = dist[:, 0] < lut[dist[:, 1]]
        qi = np.flatnonzero(keep)
        ti, d = idx[qi, 0], dist[qi, 0]
    order = np.argsort(d, kind="stable")                            # .bak:92
    qi, ti, d = qi[order], ti[order], d[order]
    if max_matches is not None:                                     # .bak:93-94
        qi, ti, d = qi[:max_matches], ti[:max_matches], d[:max_matches]
    return qi, ti, d


def match_orb_descriptors(q, t, ratio: float = 0.8):
    """``homography.match_orb_descriptors`` (homography.py:9-26).

    Keep ``(i, j)`` iff j is the row minimum, ``D[i,j] < ratio * second_min_i``
    (float64) and ``argmin_i' D[i',j] == i`` (first lowest).  A tie at the row
    minimum always fails the ratio test for ratio <= 1, so argsort's tie order
    (homography.py:15) never shows.  Needs Nt >= 2 like the reference (which
    fails to unpack ``[:2]`` otherwise).
    """
    if np.asarray(t).shape[0] < 2:
        raise ValueError("match_orb_descriptors needs at least two train descriptors")
    b, s, bw = packed_keys(q, t)
    lut = ratio_lut(ratio)
    d1, d2 = key_distance(b), key_distance(s)
    j = key_index(b)
    keep = (d1 < lut[d2]) & (key_index(bw)[j] == np.arange(b.shape[0]))
    qi = np.flatnonzero(keep)
    return [(int(i), int(j[i])) for i in qi]


def select_matches(fwd_best, fwd_second, bwd_best, *, use_ratio: bool, use_cross: bool,
                   ratio: float = 0.8, sort_by_distance: bool = True,
                   max_matches: int | None = None):
    """Selection stage on packed keys — contract of ``b2s_select_matches``.

    Generalises pipeline_match / cross_check_match / match_orb_descriptors:
    a query survives iff it has a best neighbour, (use_ratio) has a second
    neighbour and ``d1 < lut[d2]``, (use_cross) is the column minimum of its
    best train row.
    """
    fwd_best = np.asarray(fwd_best, np.uint32)
    nq = fwd_best.shape[0]
    keep = fwd_best != NONE_KEY
    j = key_index(fwd_best)
    d1 = key_distance(fwd_best)
    if use_ratio:
        fs = np.asarray(fwd_second, np.uint32)
        has2 = fs != NONE_KEY
        lut = ratio_lut(ratio)
        d2 = np.where(has2, key_distance(fs), 0)
        keep &= has2 & (d1 < lut[np.minimum(d2, 256)])
    if use_cross:
        bw = np.asarray(bwd_best, np.uint32)
        jj = np.where(keep, j, 0)
        keep &= key_index(bw)[jj] == np.arange(nq)
    qi = np.flatnonzero(keep)
    ti, d = j[qi], d1[qi]
    if sort_by_distance:
        o = np.argsort(d, kind="stable")
        qi, ti, d = qi[o], ti[o], d[o]
    if max_matches is not None and max_matches > 0:
        qi, ti, d = qi[:max_matches], ti[:max_matches], d[:max_matches]
    return qi, ti, d


# --------------------------------------------------------------------------- #
# small host-side helpers of the pipeline
# --------------------------------------------------------------------------- #

def match_stats(distances):
    """``FeaturePipeline.match_stats`` (feature_pipeline.py.bak:53-61)."""
    if len(distances) == 0:
        return 0, 0.0, 0.0
    d = np.asarray(distances, dtype=np.float32)
    return len(d), float(d.mean()), float(np.median(d))


def adaptive_ransac_threshold(pts1, pts2, base_threshold, min_threshold, max_threshold):
    """``adaptive_ransac_threshold`` (feature_pipeline.py.bak:114-129)."""
    pts1, pts2 = np.asarray(pts1), np.asarray(pts2)
    if pts1.size == 0 or pts2.size == 0:
        return float(np.clip(base_threshold, min_threshold, max_threshold))
    disp = np.linalg.norm(pts2 - pts1, axis=1)
    if disp.size == 0:
        return float(np.clip(base_threshold, min_threshold, max_threshold))
    scale = float(np.clip(float(np.median(disp)) / 25.0, 0.5, 2.0))
    return float(np.clip(base_threshold * scale, min_threshold, max_threshold))
