from spatialrgpt_b200.mm_utils import (KeywordsStoppingCriteria, boxes_to_masks, get_model_name_from_path, process_image,  # noqa: F401
                                       process_depth, process_images, process_masks, process_regions, tokenizer_image_token)
