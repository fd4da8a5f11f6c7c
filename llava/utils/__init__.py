"""llava.utils — only the helpers on the inference path's boundary."""
from typing import Any, List


def make_list(obj: Any) -> List:
    return obj if isinstance(obj, list) else [obj]


def disable_torch_init() -> None:
    """The reference skips torch's default parameter initialisation here; weights on this path are
    created uninitialised and then loaded, so there is nothing to switch off."""
