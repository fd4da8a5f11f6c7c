"""`vila-infer` (llava/cli/infer.py:100-176): same flags, same flow — llava.load -> prompt of
llava.Image / llava.Video / text -> model.generate_content(prompt, response_format=...)."""
import argparse
import importlib.util
import os

import llava
from llava import conversation as clib
from llava.media import Image, Video
from llava.model.configuration_llava import JsonSchemaResponseFormat, ResponseFormat


def get_schema_from_python_path(path: str) -> str:
    spec = importlib.util.spec_from_file_location("schema_module", os.path.abspath(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Main.schema_json()


def main(argv=None) -> str:
    parser = argparse.ArgumentParser()
    parser.add_argument("--model-path", "-m", type=str, required=True)
    parser.add_argument("--lora-path", "-l", type=str, default=None)
    parser.add_argument("--conv-mode", "-c", type=str, default="auto")
    parser.add_argument("--text", type=str)
    parser.add_argument("--media", type=str, nargs="+")
    parser.add_argument("--num_video_frames", "-nf", type=int, default=-1)
    parser.add_argument("--video_max_tiles", "-vm", type=int, default=-1)
    parser.add_argument("--json-mode", action="store_true")
    parser.add_argument("--json-schema", type=str, default=None)
    args = parser.parse_args(argv)
    if args.lora_path is not None:
        raise NotImplementedError("LoRA checkpoints are merged offline; pass the merged --model-path")
    if not args.json_mode:
        response_format = None
    elif args.json_schema is None:
        response_format = ResponseFormat(type="json_object")
    else:
        response_format = ResponseFormat(type="json_schema", json_schema=JsonSchemaResponseFormat(
            schema=get_schema_from_python_path(args.json_schema)))
    model = llava.load(args.model_path, model_base=None)
    if args.num_video_frames > 0:
        model.config.num_video_frames = args.num_video_frames
    if args.video_max_tiles > 0:
        model.config.video_max_tiles = args.video_max_tiles
    clib.default_conversation = clib.conv_templates[args.conv_mode].copy()
    prompt = []
    for media in args.media or []:
        if any(media.endswith(ext) for ext in (".jpg", ".jpeg", ".png")):
            prompt.append(Image(media))
        elif any(media.endswith(ext) for ext in (".mp4", ".mkv", ".webm")) or os.path.isdir(media):
            prompt.append(Video(media))
        else:
            raise ValueError(f"Unsupported media type: {media}")
    if args.text is not None:
        prompt.append(args.text)
    response = model.generate_content(prompt, response_format=response_format)
    print(response)
    return response


if __name__ == "__main__":
    main()
