"""Donor backbones on stock PyTorch (out of the accelerated path by design).

The head kernels only consume the last feature map of a donor; these modules exist so that
`Models.models.UNet` / `HighResolutionNet` are complete drop-ins whose state-dict keys match
the reference's checkpoints (`inc0.conv.conv.0.weight`, `stage3.1.branches.2.0.conv1.weight`,
...).  Architectures: the milesial Pytorch-UNet encoder/decoder the reference embeds
(Models/models.py:108-184, :203-211, :244-255) and HRNetV2-W48 (:318-749).  Written
table-driven from the architecture definitions, cuDNN does the work.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_MOMENTUM = 0.1
# the reference maps BatchNorm2d to SyncBatchNorm on torch>=1 (Models/bn_helper.py:10); without an
# initialised process group it behaves like BatchNorm2d
HRNetNorm = nn.SyncBatchNorm


# ----------------------------------------------------------------------------- UNet
def _two_convs(cin, cout):
    layers = []
    for a, b in ((cin, cout), (cout, cout)):
        layers += [nn.Conv2d(a, b, 3, padding=1), nn.BatchNorm2d(b), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)


class double_conv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = _two_convs(in_ch, out_ch)

    def forward(self, x):
        return self.conv(x)


class inconv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = double_conv(in_ch, out_ch)

    def forward(self, x):
        return self.conv(x)


class down(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), double_conv(in_ch, out_ch))

    def forward(self, x):
        return self.mpconv(x)


class up(nn.Module):
    def __init__(self, in_ch, out_ch, bilinear=True):
        super().__init__()
        self.up = (nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) if bilinear
                   else nn.ConvTranspose2d(in_ch // 2, in_ch // 2, 2, stride=2))
        self.conv = double_conv(in_ch, out_ch)

    def forward(self, low, skip):
        low = self.up(low)
        dy, dx = skip.size(2) - low.size(2), skip.size(3) - low.size(3)
        low = F.pad(low, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return self.conv(torch.cat([skip, low], dim=1))


class outconv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, 1)

    def forward(self, x):
        return self.conv(x)


UNET_FEATURES = 64


def attach_unet_backbone(m: nn.Module, n_channels: int):
    """Registers inc0/down1-4/up1-4 on `m` under the reference's attribute names."""
    m.inc0 = inconv(n_channels, 64)
    widths = [(64, 128), (128, 256), (256, 512), (512, 512)]
    for i, (a, b) in enumerate(widths, start=1):
        setattr(m, "down%d" % i, down(a, b))
    for i, (a, b) in enumerate([(1024, 256), (512, 128), (256, 64), (128, 64)], start=1):
        setattr(m, "up%d" % i, up(a, b))


def run_unet_backbone(m: nn.Module, x):
    skips = [m.inc0(x)]
    for i in range(1, 5):
        skips.append(getattr(m, "down%d" % i)(skips[-1]))
    d = skips[4]
    for i in range(1, 5):
        d = getattr(m, "up%d" % i)(d, skips[4 - i])
    return d  # [B, 64, H, W]


# ----------------------------------------------------------------------------- HRNet
def conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, bias=False)


class _Residual(nn.Module):
    """conv/bn chain + identity (or projected) shortcut; attribute names conv{i}/bn{i}."""
    expansion = 1
    spec = ()  # (kernel, takes_stride, out_multiplier)

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        cin = inplanes
        for i, (k, strided, mult) in enumerate(self.spec, start=1):
            cout = planes * mult
            setattr(self, "conv%d" % i, nn.Conv2d(cin, cout, kernel_size=k, stride=stride if strided else 1,
                                                  padding=k // 2, bias=False))
            setattr(self, "bn%d" % i, HRNetNorm(cout, momentum=BN_MOMENTUM))
            cin = cout
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        y = x
        last = len(self.spec)
        for i in range(1, last + 1):
            y = getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(y))
            if i != last:
                y = self.relu(y)
        shortcut = x if self.downsample is None else self.downsample(x)
        return self.relu(y + shortcut)


class BasicBlock(_Residual):
    expansion = 1
    spec = ((3, True, 1), (3, False, 1))


class Bottleneck(_Residual):
    expansion = 4
    spec = ((1, False, 1), (3, True, 1), (1, False, 4))


blocks_dict = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}


def _block_chain(block, cin, planes, count, stride=1):
    proj = None
    if stride != 1 or cin != planes * block.expansion:
        proj = nn.Sequential(nn.Conv2d(cin, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                             HRNetNorm(planes * block.expansion, momentum=BN_MOMENTUM))
    chain = [block(cin, planes, stride, proj)]
    chain += [block(planes * block.expansion, planes) for _ in range(1, count)]
    return nn.Sequential(*chain)


class HighResolutionModule(nn.Module):
    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True, align_corners=True):
        super().__init__()
        if not (num_branches == len(num_blocks) == len(num_channels) == len(num_inchannels)):
            raise ValueError("HighResolutionModule: inconsistent branch description")
        self.num_branches = num_branches
        self.fuse_method = fuse_method
        self.multi_scale_output = multi_scale_output
        self.align_corners = align_corners
        self.num_inchannels = list(num_inchannels)
        self.branches = nn.ModuleList()
        for i in range(num_branches):
            self.branches.append(_block_chain(blocks, self.num_inchannels[i], num_channels[i], num_blocks[i]))
            self.num_inchannels[i] = num_channels[i] * blocks.expansion
        self.fuse_layers = self._fusion() if num_branches > 1 else None
        self.relu = nn.ReLU(inplace=True)

    def _fusion(self):
        ch = self.num_inchannels
        rows = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j == i:
                    row.append(None)
                elif j > i:  # coarser branch -> 1x1 projection (upsampled in forward)
                    row.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 1, 1, 0, bias=False),
                                             HRNetNorm(ch[i], momentum=BN_MOMENTUM)))
                else:        # finer branch -> (i-j) stride-2 convs
                    steps = []
                    for k in range(i - j):
                        final = k == i - j - 1
                        cout = ch[i] if final else ch[j]
                        mods = [nn.Conv2d(ch[j], cout, 3, 2, 1, bias=False), HRNetNorm(cout, momentum=BN_MOMENTUM)]
                        if not final:
                            mods.append(nn.ReLU(inplace=True))
                        steps.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*steps))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def get_num_inchannels(self):
        return self.num_inchannels

    def forward(self, xs):
        if self.num_branches == 1:
            return [self.branches[0](xs[0])]
        xs = [branch(x) for branch, x in zip(self.branches, xs)]
        fused = []
        for i, row in enumerate(self.fuse_layers):
            y = xs[0] if i == 0 else row[0](xs[0])
            for j in range(1, self.num_branches):
                if j == i:
                    y = y + xs[j]
                elif j > i:
                    y = y + F.interpolate(row[j](xs[j]), size=list(xs[i].shape[-2:]), mode="bilinear",
                                          align_corners=self.align_corners)
                else:
                    y = y + row[j](xs[j])
            fused.append(self.relu(y))
        return fused


HRNET_FEATURES_W48 = 720


def attach_hrnet_backbone(m: nn.Module, extra, align_corners=True):
    """Registers stem/layer1/transition1-3/stage2-4/shared_head on `m` (reference names) from the
    MODEL.EXTRA stage description; returns the channel count of the fused feature map."""
    m.relu = nn.ReLU(inplace=True)
    m.stem = nn.Sequential(nn.Conv2d(3, 64, 3, 2, 1, bias=False), HRNetNorm(64, momentum=BN_MOMENTUM), nn.ReLU(inplace=True),
                           nn.Conv2d(64, 64, 3, 2, 1, bias=False), HRNetNorm(64, momentum=BN_MOMENTUM), nn.ReLU(inplace=True))
    m.stage1_cfg = extra["STAGE1"]
    b1 = blocks_dict[m.stage1_cfg["BLOCK"]]
    m.layer1 = _block_chain(b1, 64, m.stage1_cfg["NUM_CHANNELS"][0], m.stage1_cfg["NUM_BLOCKS"][0])
    prev = [b1.expansion * m.stage1_cfg["NUM_CHANNELS"][0]]
    for s in (2, 3, 4):
        cfg = extra["STAGE%d" % s]
        setattr(m, "stage%d_cfg" % s, cfg)
        blk = blocks_dict[cfg["BLOCK"]]
        cur = [c * blk.expansion for c in cfg["NUM_CHANNELS"]]
        setattr(m, "transition%d" % (s - 1), _transition(prev, cur))
        mods, inch = [], list(cur)
        for _ in range(cfg["NUM_MODULES"]):
            hm = HighResolutionModule(cfg["NUM_BRANCHES"], blk, cfg["NUM_BLOCKS"], inch, cfg["NUM_CHANNELS"],
                                      cfg["FUSE_METHOD"], True, align_corners)
            mods.append(hm)
            inch = hm.get_num_inchannels()
        setattr(m, "stage%d" % s, nn.Sequential(*mods))
        prev = inch
    total = int(sum(prev))
    m.shared_head = nn.Sequential(nn.Conv2d(total, total, kernel_size=1, stride=1, padding=0, bias=True),
                                  HRNetNorm(total, momentum=BN_MOMENTUM), nn.ReLU(inplace=True))
    m._hrnet_align_corners = align_corners
    return total


def _transition(prev, cur):
    out = []
    for i, c in enumerate(cur):
        if i < len(prev):
            out.append(None if prev[i] == c else
                       nn.Sequential(nn.Conv2d(prev[i], c, 3, 1, 1, bias=False), HRNetNorm(c, momentum=BN_MOMENTUM),
                                     nn.ReLU(inplace=True)))
            continue
        steps = []
        for j in range(i + 1 - len(prev)):
            cout = c if j == i - len(prev) else prev[-1]
            steps.append(nn.Sequential(nn.Conv2d(prev[-1], cout, 3, 2, 1, bias=False), HRNetNorm(cout, momentum=BN_MOMENTUM),
                                       nn.ReLU(inplace=True)))
        out.append(nn.Sequential(*steps))
    return nn.ModuleList(out)


def run_hrnet_backbone(m: nn.Module, x):
    x = m.layer1(m.stem(x))
    ys = [x]
    for s in (2, 3, 4):
        trans = getattr(m, "transition%d" % (s - 1))
        nb = getattr(m, "stage%d_cfg" % s)["NUM_BRANCHES"]
        xs = []
        for i in range(nb):
            src = ys[i] if i < len(ys) else ys[-1]
            xs.append(src if trans[i] is None else trans[i](src))
        ys = getattr(m, "stage%d" % s)(xs)
    h, w = ys[0].shape[-2:]
    ups = [ys[0]] + [F.interpolate(y, size=(h, w), mode="bilinear", align_corners=m._hrnet_align_corners) for y in ys[1:]]
    return m.shared_head(torch.cat(ups, 1))  # [B, 720, H/4, W/4]
