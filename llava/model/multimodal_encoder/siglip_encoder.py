from spatialrgpt_b200.multimodal_encoder import SiglipVisionTower  # noqa: F401
