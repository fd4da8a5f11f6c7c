"""llava/media.py — prompt parts that carry a path."""
from vila_b200.model.media import File, Image, Media, Video

__all__ = ["Media", "File", "Image", "Video"]
