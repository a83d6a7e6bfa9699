from graph_neural_mapping_b200.models.graphcnn import GIN_InfoMaxReg, GraphCNN  # noqa: F401
from graph_neural_mapping_b200.models.mlp import MLP  # noqa: F401
from graph_neural_mapping_b200.models.discriminator import Discriminator  # noqa: F401

__all__ = ["GIN_InfoMaxReg", "GraphCNN", "MLP", "Discriminator"]
