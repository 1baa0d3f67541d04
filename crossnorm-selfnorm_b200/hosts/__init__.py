"""Host models that call the CNSN operators (the callers either side of the hot path)."""
from .wideresnet import WideResNet  # noqa: F401
from .resnet import ResNet, resnet50, resnet50_ibn_a, resnet50_ibn_b  # noqa: F401
from .resnext import CifarResNeXt, resnext29  # noqa: F401
