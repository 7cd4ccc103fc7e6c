"""Pins the oracle's *operators* against dense torch convolutions in fp64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import me_cpu as ME


def _rand_coords(rng, n, lo=-6, hi=7, batches=2):
    c = np.stack([rng.integers(0, batches, n), rng.integers(lo, hi, n), rng.integers(lo, hi, n),
                  rng.integers(lo, hi, n)], 1)
    return np.unique(c, axis=0).astype(np.int32)


def _dense(coords, feats, lo, size):
    """[B, C, Z, Y, X] dense grid (x fastest in memory == last axis)."""
    B = int(coords[:, 0].max()) + 1
    g = torch.zeros(B, feats.shape[1], size, size, size, dtype=feats.dtype)
    c = torch.from_numpy(coords).long()
    g[c[:, 0], :, c[:, 3] - lo, c[:, 2] - lo, c[:, 1] - lo] = feats
    return g


def _sample(g, coords, lo, step=1):
    c = torch.from_numpy(coords).long()
    return g[c[:, 0], :, (c[:, 3] - lo) // step, (c[:, 2] - lo) // step, (c[:, 1] - lo) // step]


@pytest.mark.parametrize("seed", [0, 1])
def test_k3_s1_matches_conv3d(seed):
    rng = np.random.default_rng(seed)
    coords = _rand_coords(rng, 300)
    Cin, Cout = 5, 7
    x = torch.randn(len(coords), Cin, dtype=torch.float64)
    conv = ME.MinkowskiConvolution(Cin, Cout, kernel_size=3, stride=1, dimension=3).double()
    y = conv(ME.SparseTensor(x, coords=torch.from_numpy(coords))).F
    lo, size = -8, 18
    W = conv.kernel.detach().view(3, 3, 3, Cin, Cout).permute(4, 3, 0, 1, 2)   # [Cout,Cin,kz,ky,kx]
    yd = F.conv3d(_dense(coords, x, lo, size), W, padding=1)
    assert torch.allclose(y, _sample(yd, coords, lo), atol=1e-12)


def test_k2_s2_and_transpose_match_dense():
    rng = np.random.default_rng(3)
    coords = _rand_coords(rng, 400, lo=-8, hi=8)
    Cin, Cout = 4, 6
    x = torch.randn(len(coords), Cin, dtype=torch.float64)
    down = ME.MinkowskiConvolution(Cin, Cout, kernel_size=[2, 2, 2], stride=2, dimension=3).double()
    st = ME.SparseTensor(x, coords=torch.from_numpy(coords))
    y = down(st)
    assert y.tensor_stride == [2, 2, 2]
    lo, size = -8, 16
    W = down.kernel.detach().view(2, 2, 2, Cin, Cout).permute(4, 3, 0, 1, 2)
    yd = F.conv3d(_dense(coords, x, lo, size), W, stride=2)
    yc = y.C.numpy()
    assert (yc[:, 1:] % 2 == 0).all()
    assert torch.allclose(y.F, _sample(yd, yc, lo, step=2), atol=1e-12)
    # each fine row has exactly one parent
    maps = st.coords_man.get_kernel_map(st.coords_key, y.coords_key, down.kernel_generator, False)
    assert sum(len(i) for i, _ in maps) == len(coords)
    # transposed conv back onto the cached fine map
    up = ME.MinkowskiConvolutionTranspose(Cout, 3, kernel_size=[2, 2, 2], stride=2, dimension=3).double()
    z = up(y)
    assert z.coords_key == st.coords_key
    Wt = up.kernel.detach().view(2, 2, 2, Cout, 3).permute(3, 4, 0, 1, 2)       # [Cin,Cout,kz,ky,kx]
    gd = torch.zeros(2, Cout, 8, 8, 8, dtype=torch.float64)
    c = torch.from_numpy(yc).long()
    gd[c[:, 0], :, (c[:, 3] - lo) // 2, (c[:, 2] - lo) // 2, (c[:, 1] - lo) // 2] = y.F.detach()
    zd = F.conv_transpose3d(gd, Wt, stride=2)
    assert torch.allclose(z.F, _sample(zd, coords, lo), atol=1e-12)


def test_stride_floor_on_negatives_and_order():
    c = np.array([[0, -1, -2, 3], [0, -3, 0, 1], [1, 0, 0, 0], [0, 1, 1, 1]], np.int32)
    out = ME.stride_coords(c, 2)
    assert out.tolist() == [[0, -4, 0, 0], [0, -2, -2, 2], [0, 0, 0, 0], [1, 0, 0, 0]]


def test_hybrid_offsets_are_a_permutation_of_the_cube():
    h = ME.hybrid_offsets([3, 3, 3], [ME.RegionType.HYPERCUBE] * 3)
    c = ME.hypercube_offsets([3, 3, 3])
    assert h.shape == (27, 3) and h[0].tolist() == [0, 0, 0]
    assert h[1].tolist() == [-1, 0, 0] and h[2].tolist() == [1, 0, 0] and h[3].tolist() == [0, -1, 0]
    assert sorted(map(tuple, h)) == sorted(map(tuple, c))
    assert c[0].tolist() == [-1, -1, -1] and c[1].tolist() == [0, -1, -1] and c[13].tolist() == [0, 0, 0]


def test_kernel_map_pairs_property():
    rng = np.random.default_rng(5)
    coords = _rand_coords(rng, 500)
    offs = ME.hypercube_offsets([3, 3, 3])
    maps = ME.kernel_map(coords, coords, offs)
    for (i, j), o in zip(maps, offs):
        assert (coords[i, 1:] == coords[j, 1:] + o).all() and (coords[i, 0] == coords[j, 0]).all()
        assert (np.diff(j) > 0).all()
    # centre offset is the identity map; opposite offsets are transposes of each other
    assert (maps[13][0] == maps[13][1]).all() and len(maps[13][0]) == len(coords)
    for k in range(27):
        a = set(zip(maps[k][0].tolist(), maps[k][1].tolist()))
        b = set(zip(maps[26 - k][1].tolist(), maps[26 - k][0].tolist()))
        assert a == b
