"""MobileNetV3-large feature trunk of the landmark CNN (`stn` of Part-fViT and of
face_landmark_4simmin_glo_loc): [B,3,112,112] -> [B,160,4,4].

The trunk is OUTSIDE the hot path (stock cuDNN convolutions in the reference,
face_pre_pro/mobilenet.py, instantiated at ViT_face.py:611,1269 as MobileNetV3_backbone(mode='large')).
It is restated here, table driven, only so that the drop-in modules in vit_face.py are self-contained:
the module tree reproduces the reference's parameter / buffer names (`features.0.{0,1}.*`,
`features.{1..15}.conv.{0,1,3,4,5.fc.{0,2},7,8}.*`), so `load_state_dict(strict=True)` of a reference
checkpoint works, and the initialisation follows mobilenet.py:315-328 (kaiming-normal fan-out
convolutions, unit BatchNorm, N(0, 0.01) linears).  tests/test_oracle_vs_reference.py checks keys,
shapes and the forward against the live reference when /root/reference is present.
"""
import torch.nn as nn
import torch.nn.functional as F

# (kernel, expansion, out, squeeze-excite, hard-swish, stride): the MobileNetV3-large rows that precede
# the classifier head (Howard et al. 2019, table 1), width multiplier 1.0
_LARGE = (
    (3, 16, 16, False, False, 1),
    (3, 64, 24, False, False, 2),
    (3, 72, 24, False, False, 1),
    (5, 72, 40, True, False, 2),
    (5, 120, 40, True, False, 1),
    (5, 120, 40, True, False, 1),
    (3, 240, 80, False, True, 2),
    (3, 200, 80, False, True, 1),
    (3, 184, 80, False, True, 1),
    (3, 184, 80, False, True, 1),
    (3, 480, 112, True, True, 1),
    (3, 672, 112, True, True, 1),
    (5, 672, 160, True, True, 2),
    (5, 960, 160, True, True, 1),
    (5, 960, 160, True, True, 1),
)


class Hswish(nn.Module):
    def forward(self, x):
        return x * F.relu6(x + 3.0) / 6.0


class Hsigmoid(nn.Module):
    def forward(self, x):
        return F.relu6(x + 3.0) / 6.0


class SEModule(nn.Module):
    """Squeeze-excite with a 4x bottleneck and a hard-sigmoid gate (parameter names fc.0 / fc.2)."""

    def __init__(self, channels, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channels // reduction, channels, bias=False), Hsigmoid())

    def forward(self, x):
        b, c = x.shape[:2]
        gate = self.fc(self.avg_pool(x).view(b, c)).view(b, c, 1, 1)
        return x * gate.expand_as(x)


class MobileBottleneck(nn.Module):
    """1x1 expand -> depthwise k x k -> (SE) -> 1x1 project, residual when shapes allow.  The child
    indices of `conv` (0,1,3,4,5,7,8 carry parameters) are the reference's checkpoint keys."""

    def __init__(self, inp, oup, kernel, stride, exp, se, hswish):
        super().__init__()
        act = Hswish if hswish else (lambda: nn.ReLU(inplace=True))
        self.use_res_connect = stride == 1 and inp == oup
        self.conv = nn.Sequential(
            nn.Conv2d(inp, exp, 1, 1, 0, bias=False), nn.BatchNorm2d(exp), act(),
            nn.Conv2d(exp, exp, kernel, stride, (kernel - 1) // 2, groups=exp, bias=False), nn.BatchNorm2d(exp),
            SEModule(exp) if se else nn.Identity(), act(),
            nn.Conv2d(exp, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))

    def forward(self, x):
        return x + self.conv(x) if self.use_res_connect else self.conv(x)


class MobileNetV3LargeTrunk(nn.Module):
    """features: stem conv 3x3/2 (16 ch, hard-swish) + the 15 bottlenecks above; output stride 32."""

    def __init__(self):
        super().__init__()
        layers = [nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1, bias=False), nn.BatchNorm2d(16), Hswish())]
        inp = 16
        for kernel, exp, out, se, hs, stride in _LARGE:
            layers.append(MobileBottleneck(inp, out, kernel, stride, exp, se, hs))
            inp = out
        self.features = nn.Sequential(*layers)
        self.out_channels = inp
        for m in self.modules():                      # mobilenet.py:315-328
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)

    def forward(self, x):
        return self.features(x)
