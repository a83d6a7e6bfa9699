"""Drop-in mirror of the reference's `models/` package (GIN_InfoMaxReg, MLP, Discriminator)."""
from .mlp import MLP
from .discriminator import Discriminator
from .graphcnn import GIN_InfoMaxReg, GraphCNN

__all__ = ["GIN_InfoMaxReg", "GraphCNN", "MLP", "Discriminator"]
