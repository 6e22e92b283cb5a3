"""GPU: result records (csrc/records.cu), the device-side candidate ranking, the whole-map sweep (BASELINE config #5)
and the pair-sharded library path (configs #2 / #3) — against the separate result arrays, NumPy and the oracle."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]


def _seq_batch(F, N, seed, counts=None):
    import torch
    from b200slam.frontend import sequence_batch
    from b200slam.synthetic import tracking_sequence
    desc, kp = tracking_sequence(F, N, seed=seed)
    counts = np.full(F, N, np.int32) if counts is None else counts
    b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), counts, 0, F - 1, N)
    return b, desc, kp


@pytest.mark.parametrize("with_pose,winner_only", [(True, False), (False, False), (False, True)],
                         ids=["pack_kernel_with_pose", "fused_into_winner_kernel", "fused_winner_only"])
def test_records_equal_the_separate_result_arrays(with_pose, winner_only):
    """Both ways a record is written: by the winner kernel itself (no pose recovery: no extra launch) and by the
    stand-alone pack kernel after the pose kernels."""
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, record_bytes, unpack_records
    S = 300
    counts = np.array([700, 650, 700, 1, 700, 699, 12], np.int32)
    b, _, _ = _seq_batch(7, 700, 21, counts)
    fe = Frontend(FrontendConfig(hypotheses=256, max_matches=S, with_pose=with_pose, winner_only=winner_only))
    rec = torch.full((b.n_pairs, record_bytes(S)), 0xAB, dtype=torch.uint8, device="cuda")
    res = fe.run(b, records=rec, pair_id0=40)
    torch.cuda.synchronize()
    u = unpack_records(rec.cpu().numpy(), S)
    cnt = res.sel.count.cpu().numpy()
    np.testing.assert_array_equal(u["n_matches"], cnt)
    np.testing.assert_array_equal(u["best_h"], res.best_h.cpu().numpy())
    np.testing.assert_array_equal(u["inliers"], res.best_count.cpu().numpy())
    np.testing.assert_array_equal(u["pair_id"], 40 + np.arange(b.n_pairs))
    if with_pose:
        np.testing.assert_array_equal(u["R"].reshape(-1, 9), res.R.cpu().numpy().astype(np.float32))
        np.testing.assert_array_equal(u["t"], res.t.cpu().numpy().astype(np.float32))
    else:
        assert not u["R"].any() and not u["t"].any()
    assert not res.inlier_mask.cpu().numpy().reshape(b.n_pairs, S)[np.arange(S)[None] >= cnt[:, None]].any()   # unused tails cleared
    oq, ot, od, mk = (x.cpu().numpy().reshape(b.n_pairs, S) for x in (res.sel.out_q, res.sel.out_t, res.sel.out_d, res.inlier_mask))
    for p in range(b.n_pairs):
        c = int(cnt[p])
        np.testing.assert_array_equal(u["q"][p, :c], oq[p, :c])
        np.testing.assert_array_equal(u["t_idx"][p, :c], ot[p, :c])
        np.testing.assert_array_equal(u["d"][p, :c], od[p, :c])
        np.testing.assert_array_equal(u["inlier"][p, :c], mk[p, :c])
        assert not u["q"][p, c:].any() and not u["inlier"][p, c:].any()          # the tail of a record is zeroed, not stale
        assert int(u["inlier"][p].sum()) == int(u["inliers"][p])


def test_rank_pairs_equals_numpy_lexsort():
    import torch
    from b200slam.frontend import rank_pairs
    rng = np.random.default_rng(5)
    for n, k in ((1, 1), (7, 5), (4541, 5), (4541, 32), (3, 8)):
        score = rng.integers(0, 40, n).astype(np.int32)              # tie-heavy
        score[rng.random(n) < 0.1] = -1                               # excluded pairs
        ids = rng.permutation(n * 3)[:n].astype(np.int32)
        sel = np.minimum(score, 25).astype(np.int32)
        top, c_off, c_cnt = (x.cpu().numpy() for x in rank_pairs(torch.from_numpy(score).cuda(), k, ids=torch.from_numpy(ids).cuda(),
                                                                 sel_count=torch.from_numpy(sel).cuda(), stride=500))
        ok = np.flatnonzero(score >= 0)
        want = ok[np.lexsort((ids[ok], -score[ok]))][:k]
        want = np.concatenate([want, np.full(k - len(want), -1)])
        np.testing.assert_array_equal(top, want, err_msg=str((n, k)))
        live = want >= 0
        np.testing.assert_array_equal(c_off[live], want[live] * 500)
        np.testing.assert_array_equal(c_cnt[live], np.maximum(sel[want[live]], 0))
        assert not c_cnt[~live].any()


def _map(rng, n_kf, n_desc, true_kf):
    kf_desc = [rng.integers(0, 256, (int(rng.integers(n_desc - 60, n_desc + 1)), 32), dtype=np.uint8) for _ in range(n_kf)]
    kf_kp = [rng.uniform(-0.5, 0.5, (len(d), 2)).astype(np.float32) for d in kf_desc]
    n = len(kf_desc[true_kf])
    P = np.stack([rng.uniform(-6, 6, n), rng.uniform(-2, 2, n), rng.uniform(6, 30, n)], axis=1)
    kf_kp[true_kf] = (P[:, :2] / P[:, 2:]).astype(np.float32)
    yaw = 0.03
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    P2 = P @ R.T + np.array([0.2, 0.0, -0.8])
    perm = rng.permutation(n)
    bits = np.unpackbits(kf_desc[true_kf], axis=1)
    bits ^= (rng.random(bits.shape) < 0.06).astype(np.uint8)
    q_desc = np.packbits(bits, axis=1)[perm]
    q_kp = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 0.0007, (n, 2)))[perm].astype(np.float32)
    # a second, weaker candidate: half of another keyframe's descriptors planted into the query's tail
    return kf_desc, kf_kp, q_desc, q_kp


def test_map_sweep_equals_oracle_per_keyframe():
    """Config #5 engine: per-keyframe cross-check counts, the ranking, and the verified candidates' matches equal
    the oracle's BFMatcher(crossCheck=True).match(kf, query) sorted by distance and cut to max_matches."""
    import torch
    from b200slam.frontend import FrontendConfig, MapSweep, record_bytes, unpack_records
    from oracle import hamming_oracle as ho
    rng = np.random.default_rng(44)
    n_kf, true_kf, S = 23, 9, 200
    kf_desc, kf_kp, q_desc, q_kp = _map(rng, n_kf, 400, true_kf)
    ids = (np.arange(n_kf) * 3 + 5).astype(np.int32)
    sw = MapSweep(kf_desc, kf_kp, ids, FrontendConfig(hypotheses=256, max_matches=S, threshold=0.01), top=5, max_query_rows=512)
    sw.set_query(torch.from_numpy(q_desc).cuda(), torch.from_numpy(q_kp).cuda())
    rec = torch.zeros((5, record_bytes(S)), dtype=torch.uint8, device="cuda")
    cnt_dev = torch.zeros(n_kf, dtype=torch.int32, device="cuda")
    for _ in range(2):                                                # second query reuses the resident map
        out = sw.run(records=rec, counts_out=cnt_dev)
    torch.cuda.synchronize()
    counts = cnt_dev.cpu().numpy()
    want_cnt, want_sel = [], []
    for d in kf_desc:
        b, s, bw = ho.packed_keys(d, q_desc)
        qi, ti, dd = ho.select_matches(b, s, bw, use_ratio=False, use_cross=True, ratio=1.0, max_matches=None, sort_by_distance=True)
        want_cnt.append(len(qi))
        want_sel.append((qi[:S], ti[:S], dd[:S]))
    np.testing.assert_array_equal(counts, want_cnt)
    order = np.lexsort((ids, -np.asarray(want_cnt)))[:5]
    np.testing.assert_array_equal(out["top_idx"].cpu().numpy(), order)
    assert order[0] == true_kf
    u = unpack_records(rec.cpu().numpy(), S)
    np.testing.assert_array_equal(u["pair_id"], ids[order])
    np.testing.assert_array_equal(u["n_matches"], np.asarray(want_cnt)[order])
    for r, kf in enumerate(order):
        qi, ti, dd = want_sel[kf]
        np.testing.assert_array_equal(u["q"][r, :len(qi)], qi)
        np.testing.assert_array_equal(u["t_idx"][r, :len(qi)], ti)
        np.testing.assert_array_equal(u["d"][r, :len(qi)], dd)
    assert u["inliers"][0] > 0.8 * min(S, want_cnt[true_kf])          # the planted keyframe verifies
    t_true = np.array([0.2, 0.0, -0.8]) / np.linalg.norm([0.2, 0.0, -0.8])
    assert abs(float(u["t"][0] @ t_true)) > 0.98 and abs(np.linalg.det(u["R"][0].astype(np.float64)) - 1.0) < 1e-4
    assert (u["inliers"][1:] < 0.5 * u["inliers"][0]).all()           # random keyframes do not


def test_sharded_frontend_world1_graph_equals_eager_frontend():
    """ShardedFrontend on one rank (no process group): the captured step's records equal a plain Frontend.run."""
    import dataclasses
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, record_bytes, unpack_records
    from b200slam.sharding import ShardedFrontend
    S = 300
    b, _, _ = _seq_batch(9, 700, 33)
    cfg = FrontendConfig(hypotheses=256, max_matches=S, with_pose=True)
    sf = ShardedFrontend(cfg, b.n_pairs)
    sf.capture(b)
    sf.replay()
    sf.flush()
    torch.cuda.synchronize()
    got = unpack_records(sf.records()[0].cpu().numpy(), S)
    ref = torch.zeros((b.n_pairs, record_bytes(S)), dtype=torch.uint8, device="cuda")
    Frontend(cfg).run(b, records=ref)
    want = unpack_records(ref.cpu().numpy(), S)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_sharded_frontend_lanes_equal_single_stream():
    """A batch of 5 launch sets alternating over 3 streams (own workspaces per lane, fork / join inside the captured
    graph) writes the records the single-stream schedule writes, eagerly and replayed."""
    import torch
    from b200slam.frontend import FrontendConfig
    from b200slam.sharding import ShardedFrontend
    S = 300
    bl = [_seq_batch(9, 700, 60 + j)[0] for j in range(5)]
    cfg = FrontendConfig(hypotheses=256, max_matches=S, with_pose=True)
    got = {}
    for lanes in (1, 3):
        sf = ShardedFrontend(cfg, bl[0].n_pairs, sets_per_gather=5, lanes=lanes)
        assert sf.lanes == lanes
        for b in bl:
            sf.step(b)
        sf.flush()
        torch.cuda.synchronize()
        eager = sf.records().clone()
        sf.capture(bl)
        sf.gather.buf.zero_()
        for _ in range(2):
            sf.replay()
        sf.flush()
        torch.cuda.synchronize()
        assert torch.equal(sf.records(), eager)
        got[lanes] = eager
        sf.close()
    assert torch.equal(got[1], got[3])
    assert len({bytes(got[1][j].cpu().numpy().tobytes()) for j in range(5)}) == 5      # five different launch sets


@pytest.mark.parametrize("config", [2, 3, 5])
def test_two_gpus_equal_one_gpu(config, tmp_path):
    """world_size 2 over NCCL (torchrun): the gathered records of the pair-sharded loop-closure batch (config #3
    shape, small) and of the keyframe-sharded sweep (config #5 shape, small) equal the 1-GPU run bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out1, out2 = tmp_path / "w1.npz", tmp_path / "w2.npz"
    env = dict(os.environ, PYTHONPATH=f"{ROOT}{os.pathsep}{ROOT / 'monocular-visual-slam_b200'}")
    tool = str(ROOT / "tools" / "sharded_check.py")
    r1 = subprocess.run([sys.executable, tool, "--config", str(config), "--out", str(out1)], env=env, capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0, r1.stderr[-3000:]
    r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                         "--master-port", "29517", tool, "--config", str(config), "--out", str(out2)], env=env, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr[-3000:]
    a, b = np.load(out1), np.load(out2)
    assert set(a.files) == set(b.files) and len(a.files) > 3
    for k in a.files:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    info = json.loads(r2.stdout.strip().splitlines()[-1])
    assert info["world"] == 2
