from spatialrgpt_b200.region_extractor import MaskPooling, RegionExtractor  # noqa: F401
