"""Bring-up probe for the tcgen05 conv kernel: identity-weight and integer-pattern cases whose wrong answers reveal
WHICH layout assumption is off (row/column permutation, k-chunk order, hi/lo planes).  Prints a JSON summary."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pointcontrast_b200 import me  # noqa: E402
from pointcontrast_b200._lib import check, lib, ptr, stream  # noqa: E402


def run(X, W, tbl, impl, bias=None):
    K, Cin, Cout = W.shape
    n_out = tbl.shape[1]
    planes = torch.empty(4, K * Cin * Cout, dtype=torch.int16, device="cuda")
    check(lib.pcb_weight_prep(ptr(W), K, Cin, Cout, ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), ptr(planes[3]), stream()))
    me.CONV_IMPL = impl
    try:
        y = me._conv_forward_raw(X, tbl, None, K, n_out, Cin, Cout, planes[0], planes[1], W, bias, planes[2], planes[3])
    finally:
        me.CONV_IMPL = "tcgen05"
    torch.cuda.synchronize()
    return y


def main():
    out = {}
    torch.manual_seed(0)
    for (n, cin, cout) in ((128, 32, 32), (128, 64, 96), (300, 96, 96), (1000, 128, 128), (200, 256, 256)):
        # 1. identity-ish weights, integer-coded X: Y[r, c] must equal X[r, c] for c < min(cin, cout)
        X = (torch.arange(n, device="cuda")[:, None] * 128 + torch.arange(cin, device="cuda")[None, :]).float()
        W = torch.zeros(1, cin, cout, device="cuda")
        m = min(cin, cout)
        W[0, torch.arange(m), torch.arange(m)] = 1.0
        tbl = torch.arange(n, device="cuda", dtype=torch.int32)[None, :].contiguous()
        y = run(X, W, tbl, "tcgen05")
        ref = X @ W[0]
        err = float((y - ref).abs().max())
        rec = {"identity_err": err}
        if err > 0:
            yy = y.long().cpu()
            rec["Y[0,:8]"] = yy[0, :8].tolist(); rec["Y[1,:8]"] = yy[1, :8].tolist(); rec["Y[8,:8]"] = yy[8, :8].tolist()
            rec["Y[:,0][:12]"] = yy[:12, 0].tolist()
            rec["decode(r,c)=Y//128,Y%128 row0"] = [(int(v) // 128, int(v) % 128) for v in yy[0, :16]]
            rec["decode col0"] = [(int(v) // 128, int(v) % 128) for v in yy[:16, 0]]
        # 2. random data, 3 offsets with gaps, vs the mma.sync kernel and vs fp64
        K = 3
        Xr = torch.randn(n, cin, device="cuda")
        Wr = torch.randn(K, cin, cout, device="cuda") * 0.1
        t = torch.stack([torch.arange(n), torch.roll(torch.arange(n), 1), torch.roll(torch.arange(n), -7)]).int()
        t[1, ::3] = -1
        t[2, n // 2:] = -1
        t = t.cuda().contiguous()
        b = torch.randn(cout, device="cuda")
        y5 = run(Xr, Wr, t, "tcgen05", b)
        ym = run(Xr, Wr, t, "mma", b)
        ref = b.double()[None].repeat(n, 1)
        for k in range(K):
            idx = t[k].long()
            ok = idx >= 0
            ref[ok] += Xr.double()[idx[ok]] @ Wr[k].double()
        rec["rand_vs_mma"] = float((y5 - ym).abs().max())
        rec["rand_vs_fp64_tc5"] = float((y5.double() - ref).abs().max() / ref.abs().mean())
        rec["rand_vs_fp64_mma"] = float((ym.double() - ref).abs().max() / ref.abs().mean())
        out[f"n{n}_c{cin}x{cout}"] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
