"""Micro-benchmark of the split-operand tcgen05 kernels (forward + weight gradient) on one C1-shaped view.
    ncu --set full --clock-control none --import-source on -k regex:tcgen05 -s 4 -c 4 -o gpurun_out/tc5 python profiles/microbench_split.py --quick
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointcontrast_b200 import me, synth  # noqa: E402
from pointcontrast_b200._lib import check, lib, ptr, stream  # noqa: E402


def split(x, flags=0):
    n, C = x.shape
    planes = torch.empty(2, n * C, dtype=torch.bfloat16, device="cuda")
    check(lib.pcb_split_rows(ptr(x), C, n, C, planes[0].data_ptr(), planes[1].data_ptr(), C, flags, stream()))
    return planes


def timed(fn, flush, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def _arg(name, default=None):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def main():
    """--quick: 96->96 on the stride-1 level only.  --levels 0,4 / --shapes 96x96,256x256 / --only fwd|wgrad: pick what runs
    (for ncu captures of one kernel on one shape, e.g. the weight gradient of a 256-channel stride-16 layer)."""
    quick = "--quick" in sys.argv
    levels = [int(v) for v in _arg("--levels", "").split(",") if v] or None
    shapes = [tuple(int(c) for c in v.split("x")) for v in _arg("--shapes", "").split(",") if v] or None
    only = _arg("--only")
    batch = synth.synth_batch(0, 4)
    C = torch.from_numpy(batch["sinput0_C"])
    st = me.SparseTensor(torch.zeros(len(C), 1, device="cuda"), coords=C)
    cm = st.coords_man
    hyb = me.KernelGenerator(3, 1, 1, region_type=me.RegionType.HYBRID, axis_types=[me.RegionType.HYPERCUBE] * 3, dimension=3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    key = st.coords_key
    for level in range(1 if quick else (max(levels) + 1 if levels else 3)):
        if levels and level not in levels:
            key = cm.stride(key, [2, 2, 2])
            continue
        plan = cm.conv_plan(key, key, hyb, False)
        n = plan.n_out
        M = sum(plan.pair_counts())
        for cin, cout in (shapes or ([(96, 96)] if quick else [(96, 96), (128, 96), (32, 32), (64, 64), (128, 128), (256, 256)])):
            if level == 0 and cin > 128 and not shapes:
                continue
            X = torch.randn(n, cin, device="cuda"); dY = torch.randn(n, cout, device="cuda")
            W = torch.randn(27, cin, cout, device="cuda") * 0.05
            planes = torch.empty(4, 27 * cin * cout, dtype=torch.int16, device="cuda")
            check(lib.pcb_weight_prep(ptr(W), 27, cin, cout, ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(planes[3]), stream()))
            Xs, dYs = split(X), split(dY)
            ft = torch.empty(lib.pcb_weight_tile_bytes(27, cin, cout, 0), dtype=torch.uint8, device="cuda")
            dt = torch.empty(lib.pcb_weight_tile_bytes(27, cin, cout, 1), dtype=torch.uint8, device="cuda")
            check(lib.pcb_weight_tile(ptr(W), 27, cin, cout, ptr(ft), ptr(dt), 0, stream()))
            Y = torch.empty(n, cout, device="cuda"); dW = torch.empty(27, cin, cout, device="cuda")
            wsb = max(256, lib.pcb_conv_forward_ws_bytes(27, n, cin, cout)); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
            wsb2 = lib.pcb_conv_wgrad_split_ws_bytes(27, n, cin, cout); ws2 = torch.empty(wsb2, dtype=torch.uint8, device="cuda")

            def fwd():
                check(lib.pcb_conv_forward_split(Xs[0].data_ptr(), Xs[1].data_ptr(), cin, ptr(plan.fwd_tbl), plan.fwd_tbl.shape[1], None, 27,
                                                 n, cin, cout, ptr(ft), None, ptr(Y), cout, ptr(ws), wsb, 0, stream()))

            def wgrad():
                check(lib.pcb_conv_wgrad_split(Xs[0].data_ptr(), Xs[1].data_ptr(), cin, dYs[0].data_ptr(), dYs[1].data_ptr(), cout,
                                               ptr(plan.wg_tbl), plan.wg_tbl.shape[1], 27, n, cin, cout, ptr(dW), 0, ptr(ws2), wsb2, 0, stream()))
            alg = M * (cin + cout) * 4 + M * 8 + 27 * cin * cout * 4
            tf = timed(fwd, flush) if only != "wgrad" else float("nan")
            tw = timed(wgrad, flush) if only != "fwd" else float("nan")
            print(f"level {level} rows {n:7d} |M| {M:8d} {cin:3d}->{cout:3d}: fwd {tf*1e3:7.1f} us {alg/tf/1e6:6.0f} GB/s | wgrad {tw*1e3:7.1f} us {alg/tw/1e6:6.0f} GB/s")
        key = cm.stride(key, [2, 2, 2])


if __name__ == "__main__":
    main()
