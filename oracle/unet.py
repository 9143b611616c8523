"""Restatement of ``segmentation_models_pytorch.Unet("resnet34", encoder_weights=None,
in_channels=3, classes=3, activation=None)`` — the model the reference builds at
d3f/train_denoiser/lit_module.py:46-52 and d3f/train_deep_fake/lit_module.py:53-59.

smp is not vendored/installed; topology follows SURVEY Appendix A1-A3.  The
encoder IS torchvision's ResNet (torchvision/models/resnet.py:59-105 BasicBlock,
:197-205 stem/layers, :266-278 forward order) with avgpool/fc removed, exactly as smp's
ResNetEncoder does.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import ResNet, BasicBlock


class ResNet34Encoder(ResNet):
    """smp ResNetEncoder(block=BasicBlock, layers=[3,4,6,3], out_channels=(3,64,64,128,256,512))."""

    def __init__(self, in_channels=3):
        super().__init__(BasicBlock, [3, 4, 6, 3])
        del self.fc
        del self.avgpool
        if in_channels != 3:
            raise NotImplementedError("reference only ever uses in_channels=3")
        self.out_channels = (3, 64, 64, 128, 256, 512)

    def forward(self, x):
        f0 = x
        f1 = self.relu(self.bn1(self.conv1(x)))       # resnet.py:268-270
        f2 = self.layer1(self.maxpool(f1))            # resnet.py:271-273
        f3 = self.layer2(f2)
        f4 = self.layer3(f3)
        f5 = self.layer4(f4)
        return [f0, f1, f2, f3, f4, f5]


class Conv2dReLU(nn.Sequential):
    """smp.base.modules.Conv2dReLU with use_batchnorm=True: conv(bias=False) -> BN -> ReLU."""

    def __init__(self, cin, cout):
        super().__init__(
            nn.Conv2d(cin, cout, 3, padding=1, bias=False),
            nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True),
        )


class DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout)
        self.conv2 = Conv2dReLU(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)           # upsampled first, skip second
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    def __init__(self, encoder_channels=(3, 64, 64, 128, 256, 512), decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]                   # 512,256,128,64,64
        ins = [enc[0]] + list(decoder_channels[:-1])             # 512,256,128,64,32
        skips = enc[1:] + [0]                                    # 256,128,64,64,0
        self.blocks = nn.ModuleList(DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, decoder_channels))

    def forward(self, *feats):
        feats = feats[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class Unet(nn.Module):
    def __init__(self, encoder_name="resnet34", encoder_weights=None, in_channels=3, classes=3, activation=None):
        super().__init__()
        if encoder_name != "resnet34" or encoder_weights is not None or activation is not None:
            raise NotImplementedError("only the configuration the reference uses is restated")
        self.encoder = ResNet34Encoder(in_channels)
        self.decoder = UnetDecoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, 3, padding=1))
        self._init()

    def _init(self):
        # smp initialize_decoder / initialize_head ("Init")
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.segmentation_head.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        h, w = x.shape[-2:]
        if h % 32 != 0 or w % 32 != 0:
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and width "
                               f"divisible by 32.")
        feats = self.encoder(x)
        return self.segmentation_head(self.decoder(*feats))
