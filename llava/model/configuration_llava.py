"""llava/model/configuration_llava.py:23-121 — LlavaConfig + the response-format schema."""
from typing import Literal, Optional

from pydantic import BaseModel, Field

from vila_b200.model import LlavaConfig  # noqa: F401


class JsonSchemaResponseFormat(BaseModel):
    schema_: str = Field(alias="schema")


class ResponseFormat(BaseModel):
    type: Literal["text", "json_object", "json_schema"]
    json_schema: Optional[JsonSchemaResponseFormat] = None
