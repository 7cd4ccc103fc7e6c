"""Micro-benchmark of the sparse-conv kernels on one C1-shaped view (4 synthetic scenes, ~160k voxels at stride 1).
Run under ncu for the per-kernel captures committed in this directory:
    ncu --set full --clock-control none --import-source on -k regex:conv_mma -s 4 -c 2 -o gpurun_out/conv_mma python profiles/microbench_conv.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointcontrast_b200 import me, synth  # noqa: E402


def timed(fn, n=5, flush=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    quick = "--quick" in sys.argv
    batch = synth.synth_batch(0, 4)
    C = torch.from_numpy(batch["sinput0_C"])
    n = len(C)
    st = me.SparseTensor(torch.zeros(n, 1, device="cuda"), coords=C)
    cm = st.coords_man
    hyb = me.KernelGenerator(3, 1, 1, region_type=me.RegionType.HYBRID, axis_types=[me.RegionType.HYPERCUBE] * 3, dimension=3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > L2
    key = st.coords_key
    shapes = [(96, 96), (128, 96)] if quick else [(96, 96), (128, 96), (32, 32), (64, 64), (128, 128), (256, 256)]
    for level in range(5):
        plan = cm.conv_plan(key, key, hyb, False)
        M = sum(plan.pair_counts())
        for cin, cout in shapes:
            if level == 0 and cin > 128:
                continue
            if level > 0 and quick:
                continue
            conv = me.MinkowskiConvolution(cin, cout, kernel_size=3, kernel_generator=hyb, dimension=3).cuda()
            x = torch.randn(plan.n_in, cin, device="cuda", requires_grad=True)
            xs = me.SparseTensor(x, coords_key=key, coords_manager=cm)
            y = conv(xs).F
            dy = torch.randn_like(y)
            alg = M * (cin + cout) * 4 + M * 8 + 27 * cin * cout * 4
            t_f, _ = timed(lambda: conv(xs), flush=flush)
            def bwd():
                yy = conv(xs).F
                yy.backward(dy)
            t_fb, _ = timed(bwd, flush=flush)
            print(f"level {level} rows {plan.n_in:7d} |M| {M:8d} {cin:3d}->{cout:3d}: fwd {t_f*1e3:8.1f} us  "
                  f"{alg / t_f / 1e6:7.0f} GB/s alg  {2 * M * cin * cout / t_f / 1e9:6.1f} TF useful | fwd+bwd(dgrad+wgrad) {t_fb*1e3:8.1f} us")
        key = cm.stride(key, [2, 2, 2])


if __name__ == "__main__":
    main()
