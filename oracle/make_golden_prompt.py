"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference prompt-encoding glue (`QwenImagePipeline.encode_prompt` /
`_get_qwen_prompt_embeds`, pipeline_qwen_image.py:348-434) on CPU with the deterministic tokenizer / encoder stand-ins of
oracle/prompt_stubs.py and stores its outputs in tests/golden/prompt_encode.pt.  Build container only."""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import prompt_stubs  # noqa: E402
from oracle.make_golden_diffuse import import_reference_pipeline  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prompt_encode.pt")
PROMPTS = ["a red fox", "an astronaut riding a horse on the moon, oil painting, highly detailed", "x"]


def main():
    _, RP = import_reference_pipeline()
    pipe = object.__new__(RP.QwenImagePipeline)
    torch.nn.Module.__init__(pipe)
    pipe.tokenizer, pipe.text_encoder, pipe.device = prompt_stubs.StubTokenizer(), prompt_stubs.StubTextEncoder(), torch.device("cpu")
    pipe.tokenizer_max_length = 1024
    pipe.prompt_template_encode = ("<|im_start|>system\nDescribe the image by detailing the color, shape, size, texture, quantity, text, "
                                   "spatial relationships of the objects and background:<|im_end|>\n<|im_start|>user\n{}<|im_end|>\n"
                                   "<|im_start|>assistant\n")  # set by the reference __init__ (:283), which needs a checkpoint
    pipe.prompt_template_encode_start_idx = 34
    out = {"prompts": PROMPTS}
    e, m = pipe.encode_prompt(prompt=PROMPTS, num_images_per_prompt=2, max_sequence_length=1024)
    out["embeds_x2"], out["mask_x2"] = e, m
    e, m = pipe.encode_prompt(prompt=PROMPTS[1], num_images_per_prompt=1, max_sequence_length=40)
    out["embeds_trunc40"], out["mask_trunc40"] = e, m
    torch.save(out, GOLDEN)
    print("saved", GOLDEN, tuple(out["embeds_x2"].shape), tuple(out["embeds_trunc40"].shape))


if __name__ == "__main__":
    main()
