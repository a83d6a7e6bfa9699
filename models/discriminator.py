from graph_neural_mapping_b200.models.discriminator import Discriminator  # noqa: F401
