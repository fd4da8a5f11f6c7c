"""llava.model — the class names of the reference's package, served by vila_b200.model."""
from vila_b200.model import (BasicImageEncoder, BasicVideoEncoder, LlavaLlamaModel, MultimodalProjector,  # noqa: F401
                             Qwen2ForCausalLM, SiglipVisionTower, TSPVideoEncoder, VILAForCausalLM)
from vila_b200.model import LlavaConfig  # noqa: F401
from vila_b200.model import LlavaConfig as LlavaLlamaConfig  # noqa: F401
