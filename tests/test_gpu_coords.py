"""GPU coordinate manager vs the oracle: bit-exact levels and kernel maps."""
import numpy as np
import pytest
import torch

from oracle import me_cpu as OR
from tests.helpers import rand_coords, surface_coords, table_to_pairs

pytestmark = pytest.mark.gpu


def _mgr(coords):
    from pointcontrast_b200 import me
    st = me.SparseTensor(torch.zeros(len(coords), 4, device="cuda"), coords=torch.from_numpy(coords))
    return me, st


@pytest.mark.parametrize("seed,n,gen", [(0, 2000, rand_coords), (1, 30000, surface_coords), (2, 1, rand_coords),
                                        (3, 37, rand_coords), (4, 120000, surface_coords)])
def test_levels_and_maps_bit_exact(seed, n, gen):
    rng = np.random.default_rng(seed)
    coords = gen(rng, n)
    me, st = _mgr(coords)
    cm = st.coords_man
    ocm = OR.CoordsManager(3)
    okey = ocm.initialize(coords, [1, 1, 1])
    assert (st.C.cpu().numpy() == coords).all()
    key = st.coords_key
    cube = me.KernelGenerator(3, 1, 1, region_type=me.RegionType.HYPERCUBE, dimension=3)
    hyb = me.KernelGenerator(3, 1, 1, region_type=me.RegionType.HYBRID, axis_types=[me.RegionType.HYPERCUBE] * 3, dimension=3)
    k2 = me.KernelGenerator([2, 2, 2], 2, 1, dimension=3)
    ocube = OR.KernelGenerator(3, 1, 1, region_type=OR.RegionType.HYPERCUBE, dimension=3)
    ohyb = OR.KernelGenerator(3, 1, 1, region_type=OR.RegionType.HYBRID, axis_types=[OR.RegionType.HYPERCUBE] * 3, dimension=3)
    ok2 = OR.KernelGenerator([2, 2, 2], 2, 1, dimension=3)
    assert (cube.offsets == ocube.offsets).all() and (hyb.offsets == ohyb.offsets).all() and (k2.offsets == ok2.offsets).all()
    for level in range(4):
        for kg, okg in ((cube, ocube), (hyb, ohyb)):
            plan = cm.conv_plan(key, key, kg, False)
            ref = ocm.get_kernel_map(okey, okey, okg, False)
            got = table_to_pairs(plan.fwd_tbl)
            for (gi, gj), (ri, rj) in zip(got, ref):
                assert (gi == ri.numpy()).all() and (gj == rj.numpy()).all()
            assert plan.pair_counts() == [len(r[0]) for r in ref]
            # data-gradient table: the same table read through the opposite-offset permutation
            for k in range(27):
                assert (kg.offsets[plan.dg_kmap[k]] == -kg.offsets[k]).all()
        nkey = cm.stride(key, [2, 2, 2])
        onkey = ocm.stride(okey, [2, 2, 2])
        assert (cm.get_coords(nkey).cpu().numpy() == ocm.levels[onkey.ts]).all()
        down = cm.conv_plan(key, nkey, k2, False)
        ref = ocm.get_kernel_map(okey, onkey, ok2, False)
        for (gi, gj), (ri, rj) in zip(table_to_pairs(down.fwd_tbl), ref):
            assert (gi == ri.numpy()).all() and (gj == rj.numpy()).all()
        up = cm.conv_plan(nkey, key, k2, True)
        reft = ocm.get_kernel_map(onkey, okey, ok2, True)
        for (gi, gj), (ri, rj) in zip(table_to_pairs(up.fwd_tbl), reft):
            o = np.argsort(rj.numpy(), kind="stable")
            assert (gi == ri.numpy()[o]).all() and (gj == rj.numpy()[o]).all()
        assert sum(down.pair_counts()) == cm.num_rows(key) == sum(up.pair_counts())
        key, okey = nkey, onkey


def test_duplicate_and_range_errors():
    from pointcontrast_b200 import me, _lib
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3]], dtype=torch.int32)
    with pytest.raises(_lib.PcbError):
        me.SparseTensor(torch.zeros(2, 3, device="cuda"), coords=c)
    c = torch.tensor([[0, 40000, 2, 3]], dtype=torch.int32)
    with pytest.raises(_lib.PcbError):
        me.SparseTensor(torch.zeros(1, 3, device="cuda"), coords=c)


def test_cpu_tensor_is_refused():
    from pointcontrast_b200 import me, _lib
    c = torch.tensor([[0, 1, 2, 3]], dtype=torch.int32)
    st = me.SparseTensor(torch.zeros(1, 3), coords=c)
    conv = me.MinkowskiConvolution(3, 32, kernel_size=3, dimension=3)
    with pytest.raises(_lib.PcbError):
        conv(st)
