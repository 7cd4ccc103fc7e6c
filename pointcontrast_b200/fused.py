"""Fused training executor for Res16UNet on libpcb200.

The modular surface in `me.py` (one autograd node per MinkowskiConvolution / BatchNorm / ReLU, as in MinkowskiEngine)
is what makes the reference's model file run unchanged; it costs ~1.5k Python-dispatched autograd nodes per step.
This module runs the SAME graph (`pretrain/pointcontrast/model/res16unet.py:206-268`) as one autograd node per forward:

  * unit = conv -> BatchNorm statistics -> one elementwise pass doing normalise + residual add + ReLU
    (`model/modules/resnet_block.py:44-60` collapses to two units per BasicBlock);
  * `me.cat` is free: the two producers write straight into the column halves of one wider buffer (row strides);
  * backward is a hand-written reverse sweep: ReLU mask + BatchNorm backward + residual-gradient fan-out in one pass,
    data-gradient convs accumulate into their consumer's gradient buffer, weight gradients accumulate straight into
    the (flat) parameter gradient buffer -- no autograd bookkeeping, no intermediate copies.

Numerics are those of the modular path (same kernels, same order of operations per element).
"""
import ctypes
import os

import torch

from . import _lib, me
from ._lib import check, lib, ptr, stream

ENABLED = True
# Both views of a pair batch in ONE pass (see `stack_views`): half the launches, twice the rows per launch on the deep,
# latency-bound levels.  BatchNorm keeps the reference's per-view statistics through the row-segmented kernels.
PAIR = os.environ.get("PCB_PAIR", "1") == "1"
VIEW1_BATCH_OFFSET = 1 << 14      # batch indices of view 1 in a stacked tensor (packed keys hold batch < 65535)
# EXPERIMENTAL, off by default, not yet measured on a GPU: weight gradients on a second CUDA stream, concurrent with the
# data-gradient chain of the reverse sweep (they only share the read-only dz / activation planes).
WGRAD_STREAM = os.environ.get("PCB_WGRAD_STREAM", "0") == "1"


def stack_views(feats0, coords0, feats1, coords1, device):
    """One SparseTensor holding view 0's rows followed by view 1's, view 1's batch indices shifted by VIEW1_BATCH_OFFSET.
    Scenes never interact in the network (the batch index is part of the coordinate key), so every convolution of the
    stacked tensor equals the two separate forwards row for row; on every strided level (rows in packed-key order,
    batch most significant) view 0's rows still come first.  Returns (SparseTensor on `device`, rows of view 0)."""
    if not coords0.is_cuda and coords0.shape[0] and int(coords0[:, 0].max()) >= VIEW1_BATCH_OFFSET:
        raise _lib.PcbError(f"batch index >= {VIEW1_BATCH_OFFSET} cannot be stacked")
    n0 = coords0.shape[0]
    C = torch.cat([coords0.to(device, non_blocking=True).to(torch.int32), coords1.to(device, non_blocking=True).to(torch.int32)])
    C[n0:, 0] += VIEW1_BATCH_OFFSET
    F = torch.cat([feats0.to(device, non_blocking=True), feats1.to(device, non_blocking=True)])
    return me.SparseTensor(F, coords=C), n0


class Buf:
    """Matrix [n, C] with row stride ld.  `p`: fp32 storage (or 0), `hi`/`lo`: the same values as bf16 split planes (or 0)
    -- the operand format of the tensor-core kernels.  `owner` keeps the storages alive."""
    __slots__ = ("owner", "p", "hi", "lo", "n", "C", "ld", "slot", "_grad", "parent", "col", "device")

    def __init__(self, owner, p, n, C, ld, device, hi=0, lo=0, parent=None, col=0):
        self.owner, self.p, self.hi, self.lo, self.n, self.C, self.ld, self.device = owner, p, hi, lo, n, C, ld, device
        self.parent, self.col = parent, col
        self._grad = None
        self.slot = parent.slot if parent is not None else [False]     # [gradient buffer initialised?]

    @staticmethod
    def new(n, C, device, fp32=True, split=False):
        own, p, hi, lo = [], 0, 0, 0
        if fp32:
            t = torch.empty(n * C, dtype=torch.float32, device=device)
            own.append(t); p = t.data_ptr()
        if split:
            t2 = torch.empty(2, n * C, dtype=torch.bfloat16, device=device)
            own.append(t2); hi = t2.data_ptr(); lo = hi + 2 * n * C
        return Buf(own, p, n, C, C, device, hi, lo)

    def cols(self, c0, C):
        return Buf(self.owner, self.p + 4 * c0 if self.p else 0, self.n, C, self.ld, self.device, self.hi + 2 * c0 if self.hi else 0,
                   self.lo + 2 * c0 if self.lo else 0, parent=self, col=c0)

    def grad(self):
        """fp32 gradient buffer with the same geometry (column slices share their parent's buffer)."""
        if self._grad is None:
            if self.parent is not None:
                g = self.parent.grad()
                self._grad = Buf(g.owner, g.p + 4 * self.col, self.n, self.C, g.ld, self.device)
            else:
                t = torch.empty(self.n * self.ld, dtype=torch.float32, device=self.device)
                self._grad = Buf([t], t.data_ptr(), self.n, self.C, self.ld, self.device)
        return self._grad


def _kmap(plan_kmap):
    return me._c_int_array(plan_kmap) if plan_kmap is not None else None


class Runner:
    def __init__(self, model):
        self.model = model
        self.anchor = torch.zeros(1, requires_grad=True)

    # ------------------------------------------------------------------------------------------ launches
    def _conv(self, x, tbl, kmap, conv, transposed_roles, n_out, out, accumulate, bias=None, plan=None):
        kern = conv.kernel
        K, Cin, Cout = kern.shape
        ev = me._prof_begin() if plan is not None else None
        try:
            self._conv_launch(x, tbl, kmap, conv, transposed_roles, n_out, out, accumulate, bias)
        finally:
            me._prof_end(ev, "dgrad" if transposed_roles else "fwd", plan, K, Cin, Cout, Cin % 32 == 0 and Cout % 32 == 0)

    def _conv_launch(self, x, tbl, kmap, conv, transposed_roles, n_out, out, accumulate, bias):
        kern = conv.kernel
        K, Cin, Cout = kern.shape
        if transposed_roles:
            Cin, Cout = Cout, Cin
        st = stream()
        if Cin % 32 == 0 and Cout % 32 == 0:
            wt = conv._prepared.tiles(kern)[1 if transposed_roles else 0]          # pre-tiled split weights for these roles
            flags = 4 if accumulate else 0
            wsb = lib.pcb_conv_forward_ws_bytes(K, n_out, Cin, Cout)
            ws = me.workspace(wsb, self.device, slot=2)
            assert x.hi, "tensor-core conv needs the split planes of its input"
            check(lib.pcb_conv_forward_split(x.hi, x.lo, x.ld, ptr(tbl), tbl.shape[1], kmap, K, n_out, Cin, Cout, ptr(wt),
                                             ptr(bias), out.p, out.ld, ptr(ws), wsb, flags, st))
        else:
            # exact fp32 SIMT kernel: the 3-channel stem and output widths the tensor-core tiling does not cover (13 / 20
            # semantic classes); as a data gradient it runs on the per-offset transposed weights
            assert not accumulate and x.p, "the fp32 SIMT conv writes (never accumulates) and reads the fp32 plane"
            w = kern.detach().transpose(1, 2).contiguous() if transposed_roles else kern.detach()
            check(lib.pcb_conv_forward(x.p, x.ld, ptr(tbl), tbl.shape[1], kmap, K, n_out, Cin, Cout, None, None, None, None,
                                       ptr(w), ptr(bias), out.p, out.ld, None, 0, 0, st))

    def _wgrad_side(self, conv, plan, a_in, dz):
        """`_wgrad` on the side stream: ordered after everything issued so far on the current stream (dz is ready), own
        workspace slot; the caller joins the streams at the end of the sweep."""
        side = self.__dict__.get("_side_stream")
        if side is None:
            side = self._side_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            self._wgrad(conv, plan, a_in, dz, slot=3)
        for t in dz.owner:                         # dz is released by the main-stream loop while the side stream reads it
            t.record_stream(side)

    def _wgrad(self, conv, plan, a_in, dz, slot=0):
        kern = conv.kernel
        K, Cin, Cout = kern.shape
        if kern.grad is None:
            kern.grad = torch.zeros_like(kern)
        if plan.wg_gather_x:
            A, B, Ca, Cb, tr, rows = a_in, dz, Cin, Cout, 0, plan.n_out
        else:
            A, B, Ca, Cb, tr, rows = dz, a_in, Cout, Cin, 1, plan.n_in
        ev = me._prof_begin()
        if Ca % 32 == 0 and Cb % 32 == 0:
            wsb = lib.pcb_conv_wgrad_split_ws_bytes(K, rows, Ca, Cb)
            ws = me.workspace(wsb, self.device, slot=slot)
            check(lib.pcb_conv_wgrad_split(A.hi, A.lo, A.ld, B.hi, B.lo, B.ld, ptr(plan.wg_tbl), plan.wg_tbl.shape[1], K, rows, Ca, Cb,
                                           kern.grad.data_ptr(), tr, ptr(ws), wsb, 4, stream()))
        else:
            wsb = lib.pcb_conv_wgrad_ws_bytes(K, rows, Ca, Cb)
            ws = me.workspace(wsb, self.device, slot=slot)
            check(lib.pcb_conv_wgrad(A.p, A.ld, B.p, B.ld, ptr(plan.wg_tbl), plan.wg_tbl.shape[1], K, rows, Ca, Cb, kern.grad.data_ptr(),
                                     tr, ptr(ws), wsb, 4, stream()))
        me._prof_end(ev, "wgrad", plan, K, Cin, Cout, Ca % 32 == 0 and Cb % 32 == 0)

    # ------------------------------------------------------------------------------------------ forward
    def _unit(self, conv, bnm, a_in, plan, relu, residual=None, out=None):
        """out = [relu]( BN(conv(a_in)) [+ residual] )"""
        bn = bnm.bn
        K, Cin, Cout = conv.kernel.shape
        n = plan.n_out
        n0 = self.seg_of[id(plan)]                       # rows of view 0 at the output level (== n: a single view)
        nseg = 2 if n0 < n else 1
        z = Buf.new(n, Cout, self.device)
        self._conv(a_in, plan.fwd_tbl, _kmap(plan.fwd_kmap), conv, False, n, z, False, plan=plan)
        if out is None:
            out = Buf.new(n, Cout, self.device, split=True)
        mean = self._stat(nseg * Cout)
        invstd = self._stat(nseg * Cout)
        wsb = lib.pcb_bn_ws_bytes(n, Cout)
        ws = me.workspace(wsb, self.device)
        st = stream()
        check(lib.pcb_bn_stats_seg(z.p, z.ld, n, n0, Cout, bn.eps, bn.momentum, mean, invstd, bn.running_mean.data_ptr(),
                                   bn.running_var.data_ptr(), ptr(ws), wsb, st))
        check(lib.pcb_bn_apply_seg(z.p, z.ld, n, n0, Cout, mean, invstd, bn.weight.data_ptr(), bn.bias.data_ptr(),
                                   residual.p if residual is not None else None, residual.ld if residual is not None else 0,
                                   1 if relu else 0, out.p, out.ld, out.hi or None, out.lo or None, out.ld, st))
        self.bns.append(bn)
        self.tape.append((conv, bn, a_in, z, out, mean, invstd, plan, relu, residual, n0))
        return out

    def _stat(self, C):
        p = self.stats.data_ptr() + 4 * self.stat_off
        self.stat_off += C
        return p

    def _block(self, blk, x, plan3, plan1, out=None):
        h = self._unit(blk.conv1, blk.norm1, x, plan3, True)
        res = x if blk.downsample is None else self._unit(blk.downsample[0], blk.downsample[1], x, plan1, False)
        return self._unit(blk.conv2, blk.norm2, h, plan3, True, residual=res, out=out)

    def _stage(self, seq, x, plan3, plan1, out=None):
        blocks = list(seq)
        for i, blk in enumerate(blocks):
            x = self._block(blk, x, plan3, plan1, out if i == len(blocks) - 1 else None)
        return x

    def forward(self, sinput, view0_rows=None):
        """`view0_rows`: the input is a `stack_views` tensor whose first `view0_rows` rows are view 0."""
        m = self.model
        feats = sinput.F
        _lib.require_cuda(feats)
        self.device = feats.device
        cm = sinput.coords_man
        self.tape, self.bns = [], []
        self.stats = torch.empty(4 * sum(mod.bn.num_features for mod in m.modules() if isinstance(mod, me.MinkowskiBatchNorm)),
                                 dtype=torch.float32, device=self.device)
        self.stat_off = 0
        with torch.cuda.device(self.device):
            keys = [sinput.coords_key]
            for _ in range(4):
                keys.append(cm.stride(keys[-1], [2, 2, 2]))
            n = [cm.num_rows(k) for k in keys]
            kg3 = m.block1[0].conv1.kernel_generator
            kg1 = m.final.kernel_generator
            kg2 = m.conv1p1s2.kernel_generator
            p3 = [cm.conv_plan(k, k, kg3, False) for k in keys]
            p1 = [cm.conv_plan(k, k, kg1, False) for k in keys]
            down = [cm.conv_plan(keys[i], keys[i + 1], kg2, False) for i in range(4)]
            up = [cm.conv_plan(keys[i + 1], keys[i], kg2, True) for i in range(4)]
            p0 = cm.conv_plan(keys[0], keys[0], m.conv0p1s1.kernel_generator, False)
            if view0_rows is None or view0_rows >= n[0]:
                seg = list(n)
            else:                                # rows of view 0 per level: strided levels are sorted by key, batch most significant
                if view0_rows < 1:
                    raise _lib.PcbError("view 0 of a stacked pair is empty")
                thr = VIEW1_BATCH_OFFSET << 48
                cnt = torch.stack([(cm.levels[k.ts].keys < thr).sum() for k in keys[1:]]).tolist()
                seg = [int(view0_rows)] + [int(c) for c in cnt]
            self.seg_of = {id(p0): seg[0]}
            for l in range(5):
                self.seg_of[id(p3[l])] = seg[l]
                self.seg_of[id(p1[l])] = seg[l]
            for i in range(4):
                self.seg_of[id(down[i])] = seg[i + 1]
                self.seg_of[id(up[i])] = seg[i]
            calls = 2 if seg[0] < n[0] else 1
            P = m.PLANES
            x_in = feats.detach().contiguous().float()
            a0 = Buf([x_in], x_in.data_ptr(), n[0], x_in.shape[1], x_in.shape[1], self.device)
            a0.slot[0] = None                    # network input: no gradient wanted
            dev = self.device
            # concatenation buffers (left = decoder branch, right = encoder skip)
            cat8 = Buf.new(n[0], P[7] + m.INIT_DIM, dev, split=True)
            cat7 = Buf.new(n[1], P[6] + P[0], dev, split=True)
            cat6 = Buf.new(n[2], P[5] + P[1], dev, split=True)
            cat5 = Buf.new(n[3], P[4] + P[2], dev, split=True)
            out_p1 = self._unit(m.conv0p1s1, m.bn0, a0, p0, True, out=cat8.cols(P[7], m.INIT_DIM))
            x = self._unit(m.conv1p1s2, m.bn1, out_p1, down[0], True)
            b1 = self._stage(m.block1, x, p3[1], p1[1], out=cat7.cols(P[6], P[0]))
            x = self._unit(m.conv2p2s2, m.bn2, b1, down[1], True)
            b2 = self._stage(m.block2, x, p3[2], p1[2], out=cat6.cols(P[5], P[1]))
            x = self._unit(m.conv3p4s2, m.bn3, b2, down[2], True)
            b3 = self._stage(m.block3, x, p3[3], p1[3], out=cat5.cols(P[4], P[2]))
            x = self._unit(m.conv4p8s2, m.bn4, b3, down[3], True)
            x = self._stage(m.block4, x, p3[4], p1[4])
            self._unit(m.convtr4p16s2, m.bntr4, x, up[3], True, out=cat5.cols(0, P[4]))
            x = self._stage(m.block5, cat5, p3[3], p1[3])
            self._unit(m.convtr5p8s2, m.bntr5, x, up[2], True, out=cat6.cols(0, P[5]))
            x = self._stage(m.block6, cat6, p3[2], p1[2])
            self._unit(m.convtr6p4s2, m.bntr6, x, up[1], True, out=cat7.cols(0, P[6]))
            x = self._stage(m.block7, cat7, p3[1], p1[1])
            self._unit(m.convtr7p2s2, m.bntr7, x, up[0], True, out=cat8.cols(0, P[7]))
            x = self._stage(m.block8, cat8, p3[0], p1[0])
            fin = m.final
            out_t = torch.empty(n[0], fin.out_channels, dtype=torch.float32, device=dev)
            out = Buf([out_t], out_t.data_ptr(), n[0], fin.out_channels, fin.out_channels, dev)
            self._conv(x, p1[0].fwd_tbl, None, fin, False, n[0], out, False,
                       bias=fin.bias.detach().reshape(-1) if fin.bias is not None else None, plan=p1[0])
            for bn in self.bns:
                bn.num_batches_tracked += calls
        ctx = (self.tape, x, p1[0], self.stats)
        self.tape = None
        return out_t, ctx

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, ctx, d_out):
        tape, x_last, p_final, _stats = ctx
        m = self.model
        d_out = d_out.contiguous()
        dev = d_out.device
        self.device = dev
        with torch.cuda.device(dev):
            fin = m.final
            if fin.out_channels % 32 == 0:
                dfin = Buf.new(d_out.shape[0], d_out.shape[1], dev, fp32=False, split=True)
                check(lib.pcb_split_rows(d_out.data_ptr(), d_out.shape[1], d_out.shape[0], d_out.shape[1], dfin.hi, dfin.lo, dfin.ld, stream()))
            else:                                # e.g. 13 / 20 classes: the final layer's backward runs on the exact fp32 kernels
                dfin = Buf([d_out], d_out.data_ptr(), d_out.shape[0], d_out.shape[1], d_out.shape[1], dev)
            if fin.bias is not None:
                if fin.bias.grad is None:
                    fin.bias.grad = torch.zeros_like(fin.bias)
                fin.bias.grad += d_out.sum(0, keepdim=True)
            self._wgrad(fin, p_final, x_last, dfin)
            gx = x_last.grad()
            self._conv(dfin, p_final.dg_tbl, _kmap(p_final.dg_kmap), fin, True, p_final.n_in, gx, False, plan=p_final)
            x_last.slot[0] = True
            st = stream()
            for (conv, bn, a_in, z, out, mean, invstd, plan, relu, residual, n0) in reversed(tape):
                K, Cin, Cout = conv.kernel.shape
                n = plan.n_out
                g = out.grad()
                assert out.slot[0], "gradient of a unit output was never produced"
                tc = Cin % 32 == 0 and Cout % 32 == 0
                dz = Buf.new(n, Cout, dev, fp32=not tc, split=tc)       # consumed only by the conv kernels: split planes suffice
                gout_p, gout_ld, gout_mode = None, 0, 0
                if residual is not None and residual.slot[0] is not None:
                    rg = residual.grad()
                    gout_p, gout_ld = rg.p, rg.ld
                    gout_mode = 2 if residual.slot[0] else 1
                    residual.slot[0] = True
                for prm in (bn.weight, bn.bias):
                    if prm.grad is None:
                        prm.grad = torch.zeros_like(prm)
                wsb = lib.pcb_bn_ws_bytes(n, Cout)
                ws = me.workspace(wsb, dev)
                check(lib.pcb_bn_backward_seg(g.p, g.ld, z.p, z.ld, out.p if relu else None, out.ld, n, n0, Cout, mean, invstd,
                                              bn.weight.data_ptr(), dz.p or None, dz.ld, bn.weight.grad.data_ptr(),
                                              bn.bias.grad.data_ptr(), 1, gout_p, gout_ld, gout_mode, dz.hi or None, dz.lo or None,
                                              dz.ld, ptr(ws), wsb, st))
                if WGRAD_STREAM:
                    self._wgrad_side(conv, plan, a_in, dz)
                else:
                    self._wgrad(conv, plan, a_in, dz)
                if a_in.slot[0] is not None:
                    ga = a_in.grad()
                    self._conv(dz, plan.dg_tbl, _kmap(plan.dg_kmap), conv, True, plan.n_in, ga, bool(a_in.slot[0]), plan=plan)
                    a_in.slot[0] = True
            if WGRAD_STREAM and self.__dict__.get("_side_stream") is not None:
                torch.cuda.current_stream().wait_stream(self._side_stream)


class _FusedFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, runner, sinput, view0_rows):
        out, fctx = runner.forward(sinput, view0_rows)
        ctx.runner, ctx.fctx = runner, fctx
        return out

    @staticmethod
    def backward(ctx, d_out):
        ctx.runner.backward(ctx.fctx, d_out)
        ctx.fctx = None
        return None, None, None, None


_STAGES = ("block1", "block2", "block3", "block4", "block5", "block6", "block7", "block8")
_UNITS = (("conv0p1s1", "bn0"), ("conv1p1s2", "bn1"), ("conv2p2s2", "bn2"), ("conv3p4s2", "bn3"), ("conv4p8s2", "bn4"),
          ("convtr4p16s2", "bntr4"), ("convtr5p8s2", "bntr5"), ("convtr6p4s2", "bntr6"), ("convtr7p2s2", "bntr7"))


def matches(model):
    """Is `model` wired like `pretrain/pointcontrast/model/res16unet.py:36-268` (Res16UNet with BasicBlock stages)?  The
    executor reads the graph from the attribute names, so ANY class with this wiring -- this package's model file or the
    reference's own, unmodified -- runs fused.  Also checks what the tensor-core tiling needs (all hidden widths % 32)."""
    ok = model.__dict__.get("_fused_ok")
    if ok is None:
        ok = _matches(model)
        model.__dict__["_fused_ok"] = ok
    return ok


def _matches(m):
    try:
        for c, b in _UNITS:
            if not isinstance(getattr(m, c), me._ConvolutionBase) or not isinstance(getattr(m, b), me.MinkowskiBatchNorm):
                return False
        if not isinstance(m.final, me.MinkowskiConvolution) or m.final.kernel_volume != 1:
            return False
        widths = [m.INIT_DIM] + list(m.PLANES)
        for name in _STAGES:
            for blk in getattr(m, name):
                if not all(isinstance(getattr(blk, a), t) for a, t in (("conv1", me.MinkowskiConvolution), ("conv2", me.MinkowskiConvolution),
                                                                       ("norm1", me.MinkowskiBatchNorm), ("norm2", me.MinkowskiBatchNorm))):
                    return False
                if hasattr(blk, "conv3") or (blk.downsample is not None and len(blk.downsample) != 2):
                    return False
                widths += [blk.conv1.in_channels, blk.conv1.out_channels, blk.conv2.out_channels]
        if m.conv0p1s1.out_channels != m.INIT_DIM or m.final.in_channels != m.PLANES[7]:
            return False
        return all(w % 32 == 0 for w in widths) and m.conv0p1s1.in_channels % 32 != 0      # exact fp32 stem (3 input channels)
    except (AttributeError, TypeError):
        return False


def applicable_on(model, device):
    return (ENABLED and model.training and torch.is_grad_enabled() and torch.device(device).type == "cuda"
            and me.CONV_IMPL == "tcgen05" and not me.FORCE_SIMT and matches(model))


def applicable(model, sinput):
    return applicable_on(model, sinput.F.device)


def forward_pair(model, feats0, coords0, feats1, coords1, device):
    """Features (F0, F1) of the two views of a pair batch -- what `lib/ddp_trainer.py:290-297,392-398` gets from two
    calls of the model.  With the fused executor both views go through ONE stacked pass (`stack_views`), each BatchNorm
    still normalising every view with its own statistics; otherwise this is the two calls.  Works for any model class
    that `matches` (this package's or the reference's own `model/res16unet.py`)."""
    if PAIR and isinstance(model, me.MinkowskiNetwork) and applicable_on(model, device) and len(coords0) and len(coords1):
        s, n0 = stack_views(feats0, coords0, feats1, coords1, device)
        F = run(model, s, n0)
        if getattr(model, "normalize_feature", False):
            F = F / torch.norm(F, p=2, dim=1, keepdim=True)
        return F[:n0], F[n0:]
    F0 = model(me.SparseTensor(feats0, coords=coords0).to(device)).F
    F1 = model(me.SparseTensor(feats1, coords=coords1).to(device)).F
    return F0, F1


def run(model, sinput, view0_rows=None):
    """Final-layer features [N, out_channels] (before the optional L2 normalisation) as ONE autograd node."""
    runner = model.__dict__.get("_fused_runner")
    if runner is None:
        runner = Runner(model)
        model.__dict__["_fused_runner"] = runner
    return _FusedFunction.apply(runner.anchor, runner, sinput, view0_rows)
