from spatialrgpt_b200.multimodal_projector import MultimodalProjector  # noqa: F401
