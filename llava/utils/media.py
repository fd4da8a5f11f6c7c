"""llava/utils/media.py:93-123 — `extract_media(messages, config=None, draft=False)` in the reference's call
form: every message's "value" (str | list of str / PIL image / llava.Image / llava.Video) is flattened IN PLACE to
text with media tokens and the media are returned as {"image": [...]}.  The flattening itself is
vila_b200.model.media.extract_media (checked against the reference function)."""
from collections import defaultdict
from typing import Any, Dict, List, Optional

from vila_b200.model import media as _media

__all__ = ["extract_media"]


def extract_media(messages: List[Dict[str, Any]], config: Optional[Any] = None, draft: bool = False) -> Dict[str, List[Any]]:
    out: Dict[str, List[Any]] = defaultdict(list)
    for message in messages:
        if draft:  # keep llava.Image / llava.Video objects as they are (no file is opened)
            text = ""
            for part in (message["value"] if isinstance(message["value"], list) else [message["value"]]):
                if isinstance(part, str):
                    piece, _ = _media.extract_media(part, config)
                    text += piece
                elif isinstance(part, _media.Video):
                    out["image"].append(part)
                    text += _media.DEFAULT_IMAGE_TOKEN * config.num_video_frames
                else:
                    out["image"].append(part)
                    text += _media.DEFAULT_IMAGE_TOKEN
        else:
            text, images = _media.extract_media(message["value"], config)
            if images:
                out["image"].extend(images)
        message["value"] = text
    return out
