"""More CPU-side checks: oracle kernel-map properties under hypothesis, synthetic-data contract, config overrides,
and the JSON contract of bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import me_cpu as OR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@settings(max_examples=25, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 2), st.integers(-9, 9), st.integers(-9, 9), st.integers(-9, 9)), min_size=1, max_size=200,
                unique=True), st.sampled_from([2, 4]))
def test_oracle_stride_and_maps_properties(cs, ts):
    c = np.asarray(cs, np.int32)
    coarse = OR.stride_coords(c, ts)
    # unique, sorted by (b, x, y, z), every fine voxel has exactly one parent, parents are multiples of ts
    keys = OR.pack_keys(coarse)
    assert (np.diff(keys.astype(np.int64)) > 0).all()
    assert (coarse[:, 1:] % ts == 0).all()
    parents = np.concatenate([c[:, :1], np.floor_divide(c[:, 1:], ts) * ts], 1)
    assert set(map(tuple, parents.tolist())) == set(map(tuple, coarse.tolist()))
    # k2s2-style map: offsets {0..ts-1 step ts/2}... use the 8 child offsets of a stride-2 step on a ts/2 grid
    half = ts // 2
    fine = np.unique(np.concatenate([c[:, :1], np.floor_divide(c[:, 1:], half) * half], 1), axis=0).astype(np.int32)
    coarse2 = OR.stride_coords(fine, ts)
    offs = OR.hypercube_offsets([2, 2, 2]) * half
    maps = OR.kernel_map(fine, coarse2, offs)
    assert sum(len(i) for i, _ in maps) == len(fine)                       # each fine row feeds exactly one coarse row
    seen = np.concatenate([i for i, _ in maps])
    assert sorted(seen.tolist()) == list(range(len(fine)))
    # 3x3x3 map on the fine level: symmetric under offset negation, centre = identity
    m3 = OR.kernel_map(fine, fine, OR.hypercube_offsets([3, 3, 3]) * half)
    assert (m3[13][0] == m3[13][1]).all() and len(m3[13][0]) == len(fine)
    for k in range(27):
        assert set(zip(m3[k][0].tolist(), m3[k][1].tolist())) == set(zip(m3[26 - k][1].tolist(), m3[26 - k][0].tolist()))


def test_synthetic_batch_contract_and_determinism():
    from pointcontrast_b200 import synth
    a = synth.synth_batch(3, 2, scale=0.12)
    b = synth.synth_batch(3, 2, scale=0.12)
    for k in ("sinput0_C", "sinput1_C", "sinput0_F", "sinput1_F", "correspondences"):
        assert (a[k] == b[k]).all()
    C0, P = a["sinput0_C"], a["correspondences"]
    assert C0.dtype == np.int32 and C0.shape[1] == 4 and a["sinput0_F"].dtype == np.float32 and a["sinput0_F"].shape[1] == 3
    assert set(C0[:, 0].tolist()) == {0, 1}                                # batch index first (`ddp_data_loaders.py:68-70`)
    assert len(np.unique(OR.pack_keys(C0))) == len(C0)                     # voxelised: unique per scene
    assert P.dtype == np.int32 and P[:, 0].max() < len(C0) and P[:, 1].max() < len(a["sinput1_C"])
    assert (np.diff(P[:, 0]) >= 0).all()                                   # grouped / ascending in column 0


def test_config_defaults_and_overrides():
    from pointcontrast_b200.config import default_config
    c = default_config(["trainer.batch_size=32", "misc.num_gpus=8", "misc.nceT=0.4", "misc.weight=None", "net.normalize_feature=False"])
    assert c.trainer.batch_size == 32 and c.misc.num_gpus == 8 and c.misc.nceT == 0.4 and c.misc.weight is None
    assert c.net.normalize_feature is False and c.opt.momentum == 0.8 and c.opt.bn_momentum == 0.05 and c.misc.npos == 4096
    assert c.trainer.num_pos_per_batch == 1024 and c.trainer.num_hn_samples_per_batch == 256 and c.data.voxel_size == 0.025


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c0", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_stacked_views_are_block_diagonal_on_the_oracle():
    """The invariant behind `fused.stack_views`: with view 1's batch indices shifted by 2^14, every strided level keeps view
    0's rows first, and every kernel map of the stacked coordinates is the union of the two per-view maps with view 1's row
    indices shifted -- scenes never interact."""
    from pointcontrast_b200 import synth
    from pointcontrast_b200.fused import VIEW1_BATCH_OFFSET
    b = synth.synth_batch(5, 2, scale=0.15)
    c0, c1 = b["sinput0_C"], b["sinput1_C"].copy()
    c1s = c1.copy(); c1s[:, 0] += VIEW1_BATCH_OFFSET
    lv0, lv1, lvs = [c0], [c1], [np.concatenate([c0, c1s])]
    for ts in (2, 4, 8, 16):
        lv0.append(OR.stride_coords(lv0[-1], ts)); lv1.append(OR.stride_coords(lv1[-1], ts)); lvs.append(OR.stride_coords(lvs[-1], ts))
    for l in range(5):
        n0 = len(lv0[l])
        assert len(lvs[l]) == n0 + len(lv1[l])
        assert (lvs[l][:n0] == lv0[l]).all()
        back = lvs[l][n0:].copy(); back[:, 0] -= VIEW1_BATCH_OFFSET
        assert (back == lv1[l]).all()
    ts = 1
    for l in range(5):
        n0 = len(lv0[l])
        offs = OR.hypercube_offsets([3, 3, 3]) * ts
        ms, m0, m1 = (OR.kernel_map(c, c, offs) for c in (lvs[l], lv0[l], lv1[l]))
        for k in range(27):
            want = set(zip(m0[k][0].tolist(), m0[k][1].tolist())) | set(zip((m1[k][0] + n0).tolist(), (m1[k][1] + n0).tolist()))
            assert set(zip(ms[k][0].tolist(), ms[k][1].tolist())) == want
        if l < 4:
            n0c = len(lv0[l + 1])
            offs2 = OR.hypercube_offsets([2, 2, 2]) * ts
            ms, m0, m1 = (OR.kernel_map(f, c, offs2) for f, c in ((lvs[l], lvs[l + 1]), (lv0[l], lv0[l + 1]), (lv1[l], lv1[l + 1])))
            for k in range(8):
                want = set(zip(m0[k][0].tolist(), m0[k][1].tolist())) | set(zip((m1[k][0] + n0).tolist(), (m1[k][1] + n0c).tolist()))
                assert set(zip(ms[k][0].tolist(), ms[k][1].tolist())) == want
        ts *= 2


def test_poly_lr_and_lenient_loader_host_logic():
    """`downstream/semseg/lib/solvers.py:27-32` and `lib/utils.py:19-43` (host-side pieces of the finetune step; no GPU needed)."""
    import torch
    from pointcontrast_b200 import semseg
    from pointcontrast_b200.model import load_model
    from pointcontrast_b200.optim import PolyLR
    from tests import refload
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.SGD([p], lr=0.01)
    sch = PolyLR(opt, max_iter=60000, power=0.9)
    for s in range(1, 4):
        opt.step(); sch.step()
        assert abs(sch.get_last_lr()[0] - 0.01 * (1 - s / 60001) ** 0.9) < 1e-15
    cfg13 = refload.default_config(); cfg13["net"]["normalize_feature"] = False
    pre = load_model("Res16UNet34C")(3, 32, refload.default_config(), D=3)
    net = load_model("Res16UNet34C")(3, 13, cfg13, D=3)
    kept = semseg.load_state_with_same_shape(net, {"module.encoder." + k if False else "module." + k: v for k, v in pre.state_dict().items()})
    assert set(pre.state_dict()) - set(kept) == {"final.kernel", "final.bias"}
    assert all(torch.equal(net.state_dict()[k], pre.state_dict()[k]) for k in kept)


def test_scannet_pair_formats_and_samplers(tmp_path):
    """On-disk formats of the reference's preprocessing (`point_cloud_extractor.py:80`, `generate_list.py:20-28`), the collate
    layout (`ddp_data_loaders.py:52-112`) and the distributed infinite sampler (`lib/data_sampler.py:40-70`) -- host logic."""
    import torch
    from pointcontrast_b200 import scannet_pairs as SP
    rng = np.random.default_rng(0)
    (tmp_path / "scene0000_00" / "pcd").mkdir(parents=True)
    for i in range(3):
        np.savez(tmp_path / "scene0000_00" / "pcd" / f"{i}.npz", pcd=rng.normal(size=(50, 3)))
    with open(tmp_path / "overlap-30-full.txt", "w") as f:
        f.write("scene0000_00/pcd/0.npz scene0000_00/pcd/1.npz 0.45\nscene0000_00/pcd/1.npz scene0000_00/pcd/2.npz 0.31\n")
    pairs = SP.read_pair_list(tmp_path / "overlap-30-full.txt")
    assert pairs == [("scene0000_00/pcd/0.npz", "scene0000_00/pcd/1.npz"), ("scene0000_00/pcd/1.npz", "scene0000_00/pcd/2.npz")]
    assert SP.load_frame(tmp_path / pairs[0][0]).shape == (50, 3)
    # rotation law == expm(cross(I, axis * theta))
    from scipy.linalg import expm
    ax, th = np.array([0.3, -0.2, 0.4]), 1.1
    assert np.allclose(SP.rotation_about(ax, th), expm(np.cross(np.eye(3), ax / np.linalg.norm(ax) * th)), atol=1e-12)
    T = SP.sample_random_trans(rng.normal(size=(20, 3)), np.random.RandomState(0))
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
    # collate: batch index first, correspondences offset by the rows of the previous samples
    s0 = (np.zeros((4, 3), np.float32), np.zeros((5, 3), np.float32), np.arange(12).reshape(4, 3), np.arange(15).reshape(5, 3),
          np.ones((4, 3), np.float32), np.ones((5, 3), np.float32), np.array([[0, 1], [3, 4]]), np.eye(4))
    s1 = (np.zeros((2, 3), np.float32), np.zeros((3, 3), np.float32), np.arange(6).reshape(2, 3), np.arange(9).reshape(3, 3),
          np.ones((2, 3), np.float32), np.ones((3, 3), np.float32), np.zeros((0, 2), np.int64), np.eye(4))
    b = SP.default_collate_pair_fn([s0, s1])
    assert b["sinput0_C"].shape == (6, 4) and b["sinput0_C"][:, 0].tolist() == [0, 0, 0, 0, 1, 1] and b["sinput0_C"].dtype == torch.int32
    assert b["correspondences"].tolist() == [[0, 1], [3, 4], [4, 5]] and b["len_batch"] == [[4, 5], [2, 3]] and b["T_gt"].shape == (8, 4)
    # two ranks partition each permutation pass
    a, c = SP.DistributedInfSampler(10, 2, 0, shuffle=False), SP.DistributedInfSampler(10, 2, 1, shuffle=False)
    assert [next(a) for _ in range(5)] == [0, 2, 4, 6, 8] and [next(c) for _ in range(5)] == [1, 3, 5, 7, 9]


def test_train_loop_chunks_and_pipelining_host_logic(monkeypatch):
    """`Trainer.train()` host logic without a GPU: the iterations between two LR / checkpoint boundaries (iteration 1 and every
    `lr_update_freq`-th, `ddp_trainer.py:258-263`) run through `iter_losses`, which enqueues iteration i+1 before it waits for the loss
    of iteration i and never runs ahead of the last iteration of a chunk."""
    import types
    from pointcontrast_b200 import trainer as T
    monkeypatch.setattr(T, "quiesce_gc", lambda: None)
    log = []

    class Sched:
        def get_last_lr(self):
            return [0.1]

        def step(self):
            log.append("lr")

    tr = object.__new__(T.PointNCELossTrainer)
    tr.curr_iter, tr.data_loader, tr.lr_update_freq, tr.stat_freq, tr.is_master = 0, [None], 3, 1000, True
    tr.config = types.SimpleNamespace(opt=types.SimpleNamespace(max_iter=7))
    tr.scheduler = Sched()
    count = [0]

    def enqueue(it):
        count[0] += 1
        log.append(f"e{count[0]}")
        return count[0]

    tr._enqueue_iter = enqueue
    tr._finish_iter = lambda pending: log.append(f"f{pending}") or float(pending)
    tr._save_checkpoint = lambda i, name: log.append(f"ckpt{i}")
    tr.train()
    assert tr.curr_iter == 7 and count[0] == 7
    assert log == ["e1", "f1", "lr", "ckpt1",                                 # first iteration alone (boundary: iteration 1)
                   "e2", "e3", "f2", "f3", "lr", "ckpt3",                     # 2..3: iteration 3 enqueued before loss 2 is read
                   "e4", "e5", "f4", "e6", "f5", "f6", "lr", "ckpt6",        # 4..6
                   "e7", "f7"], log                                           # 7 = max_iter, not a boundary
    # the single-iteration entry point of the reference
    log.clear()
    assert tr._train_iter(iter([None]), None) == 8.0 and log == ["e8", "f8"]


def test_arena_hint_and_step_throttle_host_logic():
    """Two small pieces of host logic behind the timed step: the arena size hint only grows, in 64 MiB steps (so the caching allocator sees
    one block size per pass); at most `MAX_STEPS_IN_FLIGHT` training steps are enqueued and unfinished."""
    import types
    import torch
    from pointcontrast_b200 import fused, trainer as T
    step = 64 << 20
    h = 0
    for used in (10 << 20, 100 << 20, 90 << 20, 129 << 20, 1 << 20):
        nh = fused._grow_hint(h, used)
        assert nh >= h and nh % step == 0 and nh >= used
        h = nh
    assert h == 3 * step

    synced, made = [], []

    class Ev:
        def __init__(self):
            made.append(self)

        def record(self):
            self.rec = getattr(self, "rec", 0) + 1

        def synchronize(self):
            synced.append(self)

    tr = object.__new__(T.PointNCELossTrainer)
    tr.timing = None
    orig = torch.cuda.Event
    torch.cuda.Event = Ev
    try:
        for _ in range(5):
            tr._step_timing(True)
            tr._step_timing(False)
    finally:
        torch.cuda.Event = orig
    assert len(made) == T.ContrastiveLossTrainer.MAX_STEPS_IN_FLIGHT + 1          # a fixed pool, reused
    assert len(tr._steps_in_flight) == T.ContrastiveLossTrainer.MAX_STEPS_IN_FLIGHT
    assert len(synced) == 3                                                        # steps 3, 4, 5 each waited for the step two before
