"""Minimal stand-in for the un-vendored `timm` dependency (TEST INFRASTRUCTURE ONLY).

The reference imports `timm` at `src/pretrained_detector.py:10` and builds its trunk with
`timm.create_model(backbone_name, pretrained=...)` (`src/pretrained_detector.py:43`), then keeps
`nn.Sequential(*list(backbone.children())[:-1])` (`:46`).  timm is pinned only as `timm>=0.9.0`
(`requirements.txt:12`) and is not installed in this image, so this package restates the published
`efficientnet_b0` (non-`tf_`) architecture of timm>=0.9 with timm's child order and parameter names:

    conv_stem, bn1 (BatchNormAct2d = BN+SiLU), blocks (7 stages / 16 blocks), conv_head,
    bn2 (BatchNormAct2d), global_pool (avg + flatten), classifier

It exists so that `oracle/make_golden.py` can import the UNMODIFIED reference class in the build
container and run it as the ground truth.  It is validated against the independently written
`torchvision.models.efficientnet_b0` in `tests/test_oracle.py` (key map: SURVEY.md App. C).
Nothing under `deepfake_video_detection_b200/` imports this.
"""
import torch
import torch.nn as nn

__version__ = "0.9.0+standin"

# (repeats, kernel, stride, expand_ratio, out_channels) — timm `efficientnet_b0` arch_def
# ds_r1_k3_s1_e1_c16_se0.25 / ir_r2_k3_s2_e6_c24 / ir_r2_k5_s2_e6_c40 / ir_r3_k3_s2_e6_c80 /
# ir_r3_k5_s1_e6_c112 / ir_r4_k5_s2_e6_c192 / ir_r1_k3_s1_e6_c320, all se_ratio 0.25 of block input.
B0_STAGES = (
    (1, 3, 1, 1, 16),
    (2, 3, 2, 6, 24),
    (2, 5, 2, 6, 40),
    (3, 3, 2, 6, 80),
    (3, 5, 1, 6, 112),
    (4, 5, 2, 6, 192),
    (1, 3, 1, 6, 320),
)


class BatchNormAct2d(nn.BatchNorm2d):
    """BN followed by an activation, parameters named exactly like nn.BatchNorm2d."""

    def __init__(self, num_features, apply_act=True):
        super().__init__(num_features, eps=1e-5, momentum=0.1)
        self.act = nn.SiLU(inplace=False) if apply_act else nn.Identity()

    def forward(self, x):
        return self.act(super().forward(x))


class SqueezeExcite(nn.Module):
    def __init__(self, chs, rd_chs):
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, rd_chs, 1, bias=True)
        self.act1 = nn.SiLU()
        self.conv_expand = nn.Conv2d(rd_chs, chs, 1, bias=True)
        self.gate = nn.Sigmoid()

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(self.act1(self.conv_reduce(s)))
        return x * self.gate(s)


class DepthwiseSeparableConv(nn.Module):
    def __init__(self, cin, cout, k, stride, rd):
        super().__init__()
        self.has_skip = stride == 1 and cin == cout
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = BatchNormAct2d(cin)
        self.se = SqueezeExcite(cin, rd)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = BatchNormAct2d(cout, apply_act=False)

    def forward(self, x):
        y = self.bn1(self.conv_dw(x))
        y = self.se(y)
        y = self.bn2(self.conv_pw(y))
        return y + x if self.has_skip else y


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand, rd):
        super().__init__()
        mid = cin * expand
        self.has_skip = stride == 1 and cin == cout
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = BatchNormAct2d(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = BatchNormAct2d(mid)
        self.se = SqueezeExcite(mid, rd)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = BatchNormAct2d(cout, apply_act=False)

    def forward(self, x):
        y = self.bn1(self.conv_pw(x))
        y = self.bn2(self.conv_dw(y))
        y = self.se(y)
        y = self.bn3(self.conv_pwl(y))
        return y + x if self.has_skip else y


class SelectAdaptivePool2d(nn.Module):
    def forward(self, x):
        return x.mean((2, 3))


class EfficientNet(nn.Module):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.num_features = 1280
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = BatchNormAct2d(32)
        stages, cin = [], 32
        for (r, k, s, e, cout) in B0_STAGES:
            blocks = []
            for b in range(r):
                stride = s if b == 0 else 1
                rd = max(1, round(cin * 0.25))
                if e == 1:
                    blocks.append(DepthwiseSeparableConv(cin, cout, k, stride, rd))
                else:
                    blocks.append(InvertedResidual(cin, cout, k, stride, e, rd))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        self.conv_head = nn.Conv2d(320, 1280, 1, bias=False)
        self.bn2 = BatchNormAct2d(1280)
        self.global_pool = SelectAdaptivePool2d()
        self.classifier = nn.Linear(1280, num_classes)

    def forward_features(self, x):
        return self.bn2(self.conv_head(self.blocks(self.bn1(self.conv_stem(x)))))

    def forward(self, x):
        return self.classifier(self.global_pool(self.forward_features(x)))


# ---------------------------------------------------------------------------------------------------------
# `vit_base_patch16_224` (reference call site src/models.py:93, `num_classes=0`): timm's VisionTransformer with
# timm's module / parameter names — patch_embed.proj (Conv2d 16x16 s16), cls_token, pos_embed (1,197,768),
# blocks.{i}.{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}, norm (LayerNorm eps 1e-6), fc_norm / head =
# Identity for num_classes=0, global_pool='token' (the CLS row after the final norm), exact-erf GELU,
# qkv_bias=True, pre-norm blocks without LayerScale.
class PatchEmbed(nn.Module):
    def __init__(self, img=224, patch=16, cin=3, dim=768):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, patch, patch)
        self.num_patches = (img // patch) ** 2

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)          # (B, 196, dim), row-major patch order


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    def __init__(self, img=224, patch=16, dim=768, depth=12, heads=12, num_classes=1000):
        super().__init__()
        self.num_features = self.embed_dim = dim
        self.patch_embed = PatchEmbed(img, patch, 3, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, dim) * 0.02)
        self.blocks = nn.Sequential(*[Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head = nn.Linear(dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        return self.norm(self.blocks(x))

    def forward(self, x):
        return self.head(self.fc_norm(self.forward_features(x)[:, 0]))


def create_model(model_name, pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("timm stand-in: no pretrained weights offline (pass pretrained=False)")
    if model_name == "efficientnet_b0":
        return EfficientNet(num_classes=kwargs.get("num_classes", 1000))
    if model_name == "vit_base_patch16_224":
        return VisionTransformer(num_classes=kwargs.get("num_classes", 1000))
    raise ValueError(f"timm stand-in: unsupported model {model_name!r}")
