from spatialrgpt_b200.multimodal_encoder import VisionTower  # noqa: F401
