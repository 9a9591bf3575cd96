"""Pipeline registry — same mechanism and names as the reference
(vllm_omni/diffusion/registry.py:10-139): `model_index.json::_class_name` -> lazily imported
pipeline class, plus the post-process function looked up by name."""
from __future__ import annotations

import importlib

from vllm_omni_b200.diffusion.data import OmniDiffusionConfig

_DIFFUSION_MODELS = {
    # arch: (mod_folder, mod_relname, cls_name)
    "QwenImagePipeline": ("qwen_image", "pipeline_qwen_image", "QwenImagePipeline"),
    "QwenImageEditPipeline": ("qwen_image", "pipeline_qwen_image_edit", "QwenImageEditPipeline"),
    "QwenImageEditPlusPipeline": ("qwen_image", "pipeline_qwen_image_edit", "QwenImageEditPlusPipeline"),
}
_DIFFUSION_POST_PROCESS_FUNCS = {"QwenImagePipeline": "get_qwen_image_post_process_func",
                                 "QwenImageEditPipeline": "get_qwen_image_edit_post_process_func",
                                 "QwenImageEditPlusPipeline": "get_qwen_image_edit_plus_post_process_func"}


class _Registry:
    def _try_load_model_cls(self, arch: str):
        if arch not in _DIFFUSION_MODELS:
            return None
        folder, rel, cls = _DIFFUSION_MODELS[arch]
        return getattr(importlib.import_module(f"vllm_omni_b200.diffusion.models.{folder}.{rel}"), cls)

    def get_supported_archs(self):
        return list(_DIFFUSION_MODELS)


DiffusionModelRegistry = _Registry()


def initialize_model(od_config: OmniDiffusionConfig):
    model_class = DiffusionModelRegistry._try_load_model_cls(od_config.model_class_name)
    if model_class is None:
        raise ValueError(f"Model class {od_config.model_class_name} not found in diffusion model registry.")
    model = model_class(od_config=od_config)
    if getattr(model, "vae", None) is not None:
        if hasattr(model.vae, "use_slicing"):
            model.vae.use_slicing = od_config.vae_use_slicing
        if hasattr(model.vae, "use_tiling"):
            model.vae.use_tiling = od_config.vae_use_tiling
    return model


def get_diffusion_post_process_func(od_config: OmniDiffusionConfig):
    name = _DIFFUSION_POST_PROCESS_FUNCS.get(od_config.model_class_name)
    if name is None:
        return None
    folder, rel, _ = _DIFFUSION_MODELS[od_config.model_class_name]
    return getattr(importlib.import_module(f"vllm_omni_b200.diffusion.models.{folder}.{rel}"), name)(od_config)
