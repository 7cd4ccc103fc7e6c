"""GPU voxelisation / correspondence search (pointcontrast_b200/voxel.py) against numpy and scipy's cKDTree --
the CPU work of `pretrain/pointcontrast/lib/ddp_data_loaders.py:36-49,228-245`.  Integer results: exact."""
import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,voxel", [(200_000, 0.025), (1000, 0.05), (1, 0.025), (50_000, 0.3)])
def test_voxelize_matches_numpy_unique(n, voxel):
    from pointcontrast_b200 import voxel as V
    rng = np.random.default_rng(n)
    xyz = (rng.normal(size=(n, 3)) * np.array([2.0, 1.5, 0.8])).astype(np.float32)
    c = np.floor(xyz / np.float32(voxel)).astype(np.int64)
    key = ((c[:, 0] + (1 << 20)) << 42) | ((c[:, 1] + (1 << 20)) << 21) | (c[:, 2] + (1 << 20))
    _, sel = np.unique(key, return_index=True)
    coords, gsel = V.voxelize(torch.from_numpy(xyz).cuda(), voxel)
    assert (gsel.cpu().numpy() == sel).all()
    assert (coords.cpu().numpy() == c[sel]).all()


@pytest.mark.parametrize("ns,nd,r", [(30_000, 28_000, 0.0375), (500, 20_000, 0.1), (1, 1, 0.5), (4000, 10, 0.05)])
def test_radius_pairs_match_kdtree(ns, nd, r):
    from pointcontrast_b200 import voxel as V
    rng = np.random.default_rng(ns + nd)
    src = (rng.random((ns, 3)) * np.array([3.0, 3.0, 1.0])).astype(np.float32)
    dst = np.concatenate([src[rng.integers(0, ns, nd // 2 + 1)] + rng.normal(0, r / 3, (nd // 2 + 1, 3)).astype(np.float32),
                          (rng.random((nd, 3)) * 3).astype(np.float32)])[:nd].astype(np.float32)
    got = V.radius_pairs(torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda(), r).cpu().numpy()
    # reference: candidates from a (slightly wider) KD-tree query, then the same fp32 predicate the kernel evaluates
    tree = cKDTree(dst.astype(np.float64))
    ref = []
    r2 = np.float32(r) * np.float32(r)
    for i, nb in enumerate(tree.query_ball_point(src.astype(np.float64), r * 1.001 + 1e-6)):
        for j in sorted(nb):
            e = dst[j] - src[i]
            if np.float32(np.float32(e[0] * e[0]) + np.float32(e[1] * e[1])) + np.float32(e[2] * e[2]) < r2:
                ref.append((i, j))
    ref = np.asarray(ref, np.int32).reshape(-1, 2)
    assert got.shape == ref.shape and (got == ref).all()


def test_make_pair_on_the_synthetic_generator_views():
    """`voxel.make_pair` (the loader's per-sample work, `ddp_data_loaders.py:196-265`) on the raw views of the synthetic generator:
    voxel coordinates and kept points exact vs numpy (fp32), correspondences vs a KD-tree on the moved points (borderline pairs may
    differ: the transform is applied in fp32 on the GPU), and about as many pairs as the generator's own float64 matching."""
    from pointcontrast_b200 import synth, voxel as V
    raw = synth.synth_pair_raw(3, scale=0.3)
    host = synth.synth_pair(3, scale=0.3)
    out = V.make_pair(torch.from_numpy(raw["p0"]).cuda(), torch.from_numpy(raw["p1"]).cuda(), raw["T01"], 0.025)
    sel = []
    for v, p in (("0", raw["p0"]), ("1", raw["p1"])):
        c = np.floor(p / np.float32(0.025)).astype(np.int64)
        key = ((c[:, 0] + (1 << 20)) << 42) | ((c[:, 1] + (1 << 20)) << 21) | (c[:, 2] + (1 << 20))
        _, s_ = np.unique(key, return_index=True)
        sel.append(s_)
        assert (out["coords" + v].cpu().numpy() == c[s_]).all() and (out["xyz" + v].cpu().numpy() == p[s_]).all()
    T = raw["T01"]
    moved = raw["p0"][sel[0]].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    nb = cKDTree(raw["p1"][sel[1]].astype(np.float64)).query_ball_point(moved, 1.5 * 0.025)
    ref = {(i, j) for i, x in enumerate(nb) for j in x}
    got = {tuple(x) for x in out["corr"].cpu().numpy().tolist()}
    assert len(got ^ ref) <= max(2, len(ref) // 1000), (len(got), len(ref), len(got ^ ref))
    assert abs(len(got) - len(host["corr"])) <= 0.05 * len(host["corr"])


def test_scannet_pair_dataset_end_to_end(tmp_path):
    """`.npz{pcd}` frames + pair list on disk -> `ScanNetMatchPairDataset` (GPU voxelisation / matching) -> `PairLoader` batch dict
    -> one PointInfoNCE training iteration of the trainer (`ddp_data_loaders.py:144-309` feeding `ddp_trainer.py:380-440`)."""
    from pointcontrast_b200 import scannet_pairs as SP, synth
    from pointcontrast_b200.config import default_config
    from pointcontrast_b200.trainer import get_trainer
    from tests import refload
    (tmp_path / "scene0000_00" / "pcd").mkdir(parents=True)
    lines = []
    for s in range(2):
        raw = synth.synth_pair_raw(40 + s, scale=0.2, n_raw=60_000)
        # frames are stored in a common world frame: put view 1's points into view 0's frame (x0 = R^T (x1 - t))
        T = raw["T01"]
        np.savez(tmp_path / "scene0000_00" / "pcd" / f"{2 * s}.npz", pcd=raw["p0"].astype(np.float64))
        np.savez(tmp_path / "scene0000_00" / "pcd" / f"{2 * s + 1}.npz", pcd=(raw["p1"].astype(np.float64) - T[:3, 3]) @ T[:3, :3])
        lines.append(f"scene0000_00/pcd/{2 * s}.npz scene0000_00/pcd/{2 * s + 1}.npz 0.4")
    (tmp_path / "overlap-30-full.txt").write_text("\n".join(lines) + "\n")
    dcfg = refload.Cfg(data=dict(voxel_size=0.025, dataset_root_dir=str(tmp_path), scannet_match_dir="overlap-30-full.txt"),
                       trainer=dict(positive_pair_search_voxel_size_multiplier=1.5, min_scale=0.8, max_scale=1.2, rotation_range=360))
    ds = SP.ScanNetMatchPairDataset("train", transform=SP.Jitter(), config=dcfg, manual_seed=True, device="cuda")
    assert len(ds) == 2
    xyz0, xyz1, c0, c1, f0, f1, matches, trans = ds[0]
    assert c0.shape[1] == 3 and len(c0) == len(xyz0) == len(f0) and len(matches) > 100 and trans.shape == (4, 4)
    assert (c0 == np.floor(xyz0 / np.float32(0.025)).astype(np.int32)).all()
    d = np.linalg.norm((xyz0[matches[:, 0]] @ trans[:3, :3].T + trans[:3, 3]) - xyz1[matches[:, 1]], axis=1)
    assert d.max() < 1.5 * 0.025 * 1.2 * 1.001                 # every pair within the (scaled) search radius
    loader = SP.PairLoader(ds, batch_size=2, shuffle=False)
    batch = next(iter(loader))
    assert batch["sinput0_C"].dtype == torch.int32 and batch["sinput0_C"][:, 0].max() == 1 and batch["correspondences"].shape[1] == 2
    cfg = default_config(["trainer.batch_size=2", "misc.nceT=0.4"])
    torch.manual_seed(0)
    tr = get_trainer("PointNCELossTrainer")(cfg, loader)
    loss = tr._train_iter(iter(loader), None)
    assert np.isfinite(loss) and 0 < loss < 20
