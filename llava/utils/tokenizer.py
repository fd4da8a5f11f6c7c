"""llava/utils/tokenizer.py — `tokenize_conversation` (:74-115) and `infer_stop_tokens` (:174-183) in the reference's
call form, for the chat-template route (SeparatorStyle.AUTO, what NVILA checkpoints use): messages are
{"from": "human" | "gpt", "value": str}; media tokens in the text become their token ids because they are
registered as special tokens (vila_b200.model.loading.prepare_tokenizer)."""
from typing import Dict, Optional, Sequence

import torch

from vila_b200.model.loading import infer_stop_tokens  # noqa: F401

__all__ = ["tokenize_conversation", "infer_stop_tokens"]

_ROLE = {"human": "user", "gpt": "assistant"}


def tokenize_conversation(messages: Sequence[Dict[str, str]], tokenizer, add_generation_prompt: bool = False,
                          overrides: Optional[Dict[str, str]] = None, no_system_prompt: bool = False) -> torch.Tensor:
    turns = [{"role": "system", "content": ""}] if no_system_prompt else []
    for m in messages:
        m["value"] = m["value"].strip()  # the reference normalises the caller's messages in place
        if m["from"] not in _ROLE:
            raise ValueError(f"Unexpected sender '{m['from']}' in conversation entry.")
        content = overrides[m["from"]] if overrides is not None and m["from"] in overrides else m["value"]
        turns.append({"role": _ROLE[m["from"]], "content": content})
    text = tokenizer.apply_chat_template(turns, add_generation_prompt=add_generation_prompt, tokenize=False)
    return tokenizer(text, return_tensors="pt").input_ids[0]
