"""llava/mm_utils.py — the host-side image plumbing names callers import."""
from vila_b200.model.media import (dynamic_preprocess, dynamic_s2_preprocess, expand2square,  # noqa: F401
                                   find_closest_aspect_ratio, process_image, process_images)


def get_model_name_from_path(model_path: str) -> str:
    """mm_utils.py:582-588"""
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]
