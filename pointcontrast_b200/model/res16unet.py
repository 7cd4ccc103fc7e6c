"""Res16UNet (34C and siblings) on the MinkowskiEngine-compatible surface of `pointcontrast_b200.me`.

Same op graph, constructor signature `Model(in_channels, out_channels, config, D=3)` and state_dict keys as
`pretrain/pointcontrast/model/res16unet.py:17-275` + `model/resnet.py:99-140` + `model/modules/resnet_block.py:13-60`
(and, with `normalize_feature=False`, as `downstream/semseg/models/res16unet.py`), written data-driven rather than
stage by stage.  Quirk kept on purpose: the norms inside residual blocks use BatchNorm momentum
0.1, because the reference's `_make_layer` does not forward `bn_momentum` to the block (`model/resnet.py:121-138`).
"""
import torch
import torch.nn as nn

from .. import fused
from .. import me as ME


def _conv(cin, cout, kernel_size, stride=1, bias=False, hybrid=False, D=3):
    if hybrid:       # ConvType.SPATIAL_HYPERCUBE_TEMPORAL_HYPERCROSS (`model/modules/common.py:107-114`)
        kg = ME.KernelGenerator(kernel_size, stride, 1, region_type=ME.RegionType.HYBRID,
                                axis_types=[ME.RegionType.HYPERCUBE] * 3, dimension=D)
    else:            # ConvType.SPATIAL_HYPERCUBE / HYPERCUBE
        kernel_size = [kernel_size] * 3 if isinstance(kernel_size, int) else list(kernel_size)[:3]
        kg = ME.KernelGenerator(kernel_size, stride, 1, region_type=ME.RegionType.HYPERCUBE, dimension=D)
    return ME.MinkowskiConvolution(in_channels=cin, out_channels=cout, kernel_size=kernel_size, stride=stride, dilation=1,
                                   has_bias=bias, kernel_generator=kg, dimension=D)


def _conv_tr(cin, cout, D=3):
    kg = ME.KernelGenerator([2, 2, 2], 2, 1, region_type=ME.RegionType.HYPERCUBE, dimension=D)
    return ME.MinkowskiConvolutionTranspose(in_channels=cin, out_channels=cout, kernel_size=[2, 2, 2], stride=2, dilation=1,
                                            has_bias=False, kernel_generator=kg, dimension=D)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, downsample=None, hybrid=True, bn_momentum=0.1, D=3):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, hybrid=hybrid, D=D)
        self.norm1 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = _conv(planes, planes, 3, hybrid=hybrid, D=D)
        self.norm2 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        residual = x if self.downsample is None else self.downsample(x)
        out += residual
        return self.relu(out)


class Res16UNetBase(ME.MinkowskiNetwork):
    BLOCK = None
    PLANES = (32, 64, 128, 256, 256, 256, 256, 256)
    LAYERS = (2, 2, 2, 2, 2, 2, 2, 2)
    INIT_DIM = 32
    OUT_PIXEL_DIST = 1

    def __init__(self, in_channels, out_channels, config, D=3):
        super().__init__(D)
        assert self.BLOCK is not None and D == 3
        self.in_channels, self.out_channels, self.config = in_channels, out_channels, config
        bn_m = config.opt.bn_momentum
        P, L = self.PLANES, self.LAYERS
        self.inplanes = self.INIT_DIM
        self.conv0p1s1 = _conv(in_channels, self.inplanes, config.net.conv1_kernel_size, D=D)
        self.bn0 = ME.MinkowskiBatchNorm(self.inplanes, momentum=bn_m)
        # encoder: k2s2 down-conv + BN, then residual stage
        for i, (name_c, name_b) in enumerate((("conv1p1s2", "bn1"), ("conv2p2s2", "bn2"), ("conv3p4s2", "bn3"),
                                              ("conv4p8s2", "bn4"))):
            setattr(self, name_c, _conv(self.inplanes, self.inplanes, [2, 2, 2], stride=2, D=D))
            setattr(self, name_b, ME.MinkowskiBatchNorm(self.inplanes, momentum=bn_m))
            setattr(self, f"block{i + 1}", self._make_layer(P[i], L[i], bn_m))
        # decoder: k2s2 transposed conv + BN, concat skip, residual stage
        skips = (P[2], P[1], P[0], self.INIT_DIM)
        for i, (name_c, name_b) in enumerate((("convtr4p16s2", "bntr4"), ("convtr5p8s2", "bntr5"),
                                              ("convtr6p4s2", "bntr6"), ("convtr7p2s2", "bntr7"))):
            setattr(self, name_c, _conv_tr(self.inplanes, P[4 + i], D=D))
            setattr(self, name_b, ME.MinkowskiBatchNorm(P[4 + i], momentum=bn_m))
            self.inplanes = P[4 + i] + skips[i] * self.BLOCK.expansion
            setattr(self, f"block{i + 5}", self._make_layer(P[4 + i], L[4 + i], bn_m))
        self.final = _conv(P[7], out_channels, 1, bias=True, D=D)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.normalize_feature = config.net.normalize_feature
        for m in self.modules():                     # `model/resnet.py:93-97`
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, planes, blocks, bn_momentum):
        downsample = None
        if self.inplanes != planes * self.BLOCK.expansion:
            downsample = nn.Sequential(_conv(self.inplanes, planes * self.BLOCK.expansion, 1, D=self.D),
                                       ME.MinkowskiBatchNorm(planes * self.BLOCK.expansion, momentum=bn_momentum))
        layers = [self.BLOCK(self.inplanes, planes, downsample=downsample, D=self.D)]
        self.inplanes = planes * self.BLOCK.expansion
        layers += [self.BLOCK(self.inplanes, planes, D=self.D) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        # training on CUDA never gets here: `me.MinkowskiNetwork.__call__` runs the whole graph as one fused autograd node
        out = self._forward_modular(x)
        if self.normalize_feature:               # `model/res16unet.py:262-266` (no epsilon)
            return ME.SparseTensor(out.F / torch.norm(out.F, p=2, dim=1, keepdim=True), coords_key=out.coords_key,
                                   coords_manager=out.coords_man)
        return out

    def forward_pair(self, feats0, coords0, feats1, coords1, device):
        """Features (F0, F1) of the two views of a pair batch: see `fused.forward_pair`."""
        return fused.forward_pair(self, feats0, coords0, feats1, coords1, device)

    def _forward_modular(self, x):
        out_p1 = self.relu(self.bn0(self.conv0p1s1(x)))
        out_b1p2 = self.block1(self.relu(self.bn1(self.conv1p1s2(out_p1))))
        out_b2p4 = self.block2(self.relu(self.bn2(self.conv2p2s2(out_b1p2))))
        out_b3p8 = self.block3(self.relu(self.bn3(self.conv3p4s2(out_b2p4))))
        out = self.block4(self.relu(self.bn4(self.conv4p8s2(out_b3p8))))
        out = self.block5(ME.cat(self.relu(self.bntr4(self.convtr4p16s2(out))), out_b3p8))
        out = self.block6(ME.cat(self.relu(self.bntr5(self.convtr5p8s2(out))), out_b2p4))
        out = self.block7(ME.cat(self.relu(self.bntr6(self.convtr6p4s2(out))), out_b1p2))
        out = self.block8(ME.cat(self.relu(self.bntr7(self.convtr7p2s2(out))), out_p1))
        return self.final(out)


class Res16UNet14(Res16UNetBase):
    BLOCK = BasicBlock
    LAYERS = (1, 1, 1, 1, 1, 1, 1, 1)


class Res16UNet18(Res16UNetBase):
    BLOCK = BasicBlock
    LAYERS = (2, 2, 2, 2, 2, 2, 2, 2)


class Res16UNet34(Res16UNetBase):
    BLOCK = BasicBlock
    LAYERS = (2, 3, 4, 6, 2, 2, 2, 2)


class Res16UNet34C(Res16UNet34):
    PLANES = (32, 64, 128, 256, 256, 128, 96, 96)
