"""b200unet — B200-native (sm_100a) U-Net hot path, drop-in for minghanz/pytorch-unet's `UNet`.

    from b200unet import UNet
    model = UNet(in_channels=1, n_classes=2, depth=5, wf=6, padding=False, batch_norm=False, up_mode='upconv').cuda()
    loss = F.cross_entropy(model(x), y)      # the reference's README loop works unchanged
    loss = model.loss(x, y)                  # same value, classifier fused into the loss kernel

The compute path is libb200unet.so (C ABI in include/b200unet.h); importing this package does not need a GPU,
running it does.
"""
from . import cvo, ops  # noqa: F401
from ._lib import IMPL_AUTO, IMPL_DIRECT, IMPL_UMMA, LIB_PATH, load as load_library  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .input import ImageStager, PackedImages, pack_images  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .unet import UNet, UNetConvBlock, UNetUpBlock, UNetUpBlockDeep  # noqa: F401

__all__ = ["UNet", "FusedAdam", "GraphedTrainStep", "ImageStager", "PackedImages", "pack_images", "UNetConvBlock", "UNetUpBlock", "UNetUpBlockDeep", "ops", "cvo", "load_library", "IMPL_AUTO",
           "IMPL_DIRECT", "IMPL_UMMA"]
