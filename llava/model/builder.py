from spatialrgpt_b200.builder import is_mm_model, load_pretrained_model, prepare_config_for_eval  # noqa: F401
