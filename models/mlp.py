from graph_neural_mapping_b200.models.mlp import MLP  # noqa: F401
