"""Model zoo (API-compatible with the reference's ``simple_models``)."""
from .base import BlockPartitioned
from .classifiers import Net, Net1, Net2
from .resnet import BasicBlock, Bottleneck, ResNet, ResNet18, ResNet9
from .vae import AutoEncoderCNN, AutoEncoderCNNCL
from .cpc import EncoderCNN, ContextgenCNN, PredictorCNN

__all__ = [
    "BlockPartitioned", "Net", "Net1", "Net2", "BasicBlock", "Bottleneck", "ResNet", "ResNet18", "ResNet9",
    "AutoEncoderCNN", "AutoEncoderCNNCL", "EncoderCNN", "ContextgenCNN", "PredictorCNN",
]
