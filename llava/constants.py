"""llava/constants.py — the constants callers of the hot path import (media / sentinel tokens, label mask)."""
IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_VIDEO_TOKEN = "<vila/video>"
SENTINEL_TOKEN = "<vila/sentinel>"
MEDIA_TOKENS = {"image": DEFAULT_IMAGE_TOKEN, "video": DEFAULT_VIDEO_TOKEN}
