from spatialrgpt_b200.mm_utils import (KeywordsStoppingCriteria, boxes_to_masks, expand2square, get_model_name_from_path,  # noqa: F401
                                       is_gemma_tokenizer, load_image_from_base64, process_image,
                                       process_depth, process_images, process_masks, process_regions, tokenizer_image_token)
