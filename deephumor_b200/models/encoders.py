"""Image / label encoders with the reference's module layout (deephumor/models/encoders.py:7-144).

Parameters live in plain torch containers whose names reproduce the reference ``state_dict`` exactly
(SURVEY.md Appendix C); compute goes through ``runtime.encoder.EncoderRT`` (hand-written CUDA).
"""
import torch
from torch import nn

from ..runtime.encoder import EncoderRT, RESNET_BLOCKS
from ._base import RTModule, prefixed


def _stage(images, device):
    """Device fp32 (normalised) or uint8 (raw 0..255 pixels, normalised in the stem kernel with the ImageNet statistics
    of the reference's preprocessing) images; a pinned HOST batch is passed through as is: the encoder runtime streams it in
    chunks so the host->device copy overlaps the trunk (an extension -- the reference takes device tensors only)."""
    if images.device.type == 'cpu' and images.dtype in (torch.float32, torch.uint8) and images.is_contiguous() \
            and images.is_pinned() and images.shape[0] > 64:
        return images
    if images.dtype == torch.uint8:          # raw pixels: ToTensor + Normalize happen inside the stem kernel
        return images.to(device).contiguous()
    return images.to(device, torch.float32).contiguous()


class _Bottleneck(nn.Module):
    """Parameter container matching torchvision's Bottleneck attribute names (never executed by torch)."""

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))


def _resnet50_trunk():
    """children()[:-2] of torchvision resnet50: indices 0 conv1, 1 bn1, 2 relu, 3 maxpool, 4..7 layer1..4."""
    mods = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1)]
    inplanes = 64
    for li, (nblk, planes) in enumerate(zip(RESNET_BLOCKS, (64, 128, 256, 512))):
        blocks = []
        for b in range(nblk):
            blocks.append(_Bottleneck(inplanes, planes, 2 if (b == 0 and li > 0) else 1, b == 0))
            inplanes = planes * 4
        mods.append(nn.Sequential(*blocks))
    trunk = nn.Sequential(*mods)
    for m in trunk.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
    for p in trunk.parameters():
        p.requires_grad = False                      # frozen trunk (encoders.py:35-36)
    return trunk


class ImageEncoder(RTModule):
    """encoders.py:7-70.  Random-init trunk (no network for pretrained weights; load a checkpoint instead)."""

    def __init__(self, emb_dim=256, dropout=0.2, spatial_features=False):
        super().__init__()
        self.spatial_features = spatial_features
        self.resnet = _resnet50_trunk()
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.linear = nn.Linear(2048, emb_dim)
        self.bn = nn.BatchNorm1d(emb_dim)
        self.dropout = nn.Dropout(dropout)

    def _rt(self):
        return self._get_rt('enc', lambda dt, dev: EncoderRT(prefixed(self, 'm'), 'm', dt, dev,
                                                            spatial=self.spatial_features))

    def forward(self, images):
        rt = self._rt()
        start, sp = rt.forward(_stage(images, self._device()))
        if self.spatial_features:
            return start, sp.float().view(images.shape[0], 49, -1)
        return start


class LabelEncoder(RTModule):
    """encoders.py:73-106."""

    def __init__(self, num_tokens, emb_dim=256, dropout=0.2):
        super().__init__()
        self.embedding = nn.Embedding(num_tokens, emb_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, labels):
        from ..runtime import ops
        dev = self._device()
        out = torch.empty(labels.shape[0], self.embedding.weight.shape[1], dtype=torch.float32, device=dev)
        ops.embed_mean(self.embedding.weight.detach().float().contiguous(), labels.to(dev).contiguous(), out)
        return out


class ImageLabelEncoder(RTModule):
    """encoders.py:109-144."""

    def __init__(self, num_tokens, emb_dim=256, dropout=0.2):
        super().__init__()
        self.image_encoder = ImageEncoder(emb_dim, dropout)
        self.label_encoder = LabelEncoder(num_tokens, emb_dim, dropout)
        self.linear = nn.Linear(2 * emb_dim, emb_dim)
        self.dropout = nn.Dropout(dropout)

    def _rt(self):
        return self._get_rt('enc', lambda dt, dev: EncoderRT(prefixed(self, 'm'), 'm.image_encoder', dt, dev,
                                                            label_prefix='m.label_encoder', fuse_prefix='m'))

    def forward(self, images, labels):
        dev = self._device()
        start, _ = self._rt().forward(_stage(images, dev), labels.to(dev))
        return start
