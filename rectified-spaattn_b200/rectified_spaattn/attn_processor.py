"""Mirror of the reference rectified_spaattn/attn_processor.py: install helpers for diffusers-style attention
processors (any module exposing get_processor / set_processor).  diffusers itself is not required."""
from typing import Dict, Union


def get_attn_processors(module) -> Dict[str, object]:
    """{"<qualified name>.processor": processor} for every sub-module with `get_processor` (reference :6-27)."""
    found = {}

    def walk(prefix, mod):
        if hasattr(mod, "get_processor"):
            found[f"{prefix}.processor"] = mod.get_processor()
        for name, child in mod.named_children():
            walk(f"{prefix}.{name}", child)

    for name, child in module.named_children():
        walk(name, child)
    return found


def set_attn_processor(module, processor: Union[object, Dict[str, object]]):
    """Sets one processor everywhere, or a dict keyed like get_attn_processors() (reference :30-62).
    Raises ValueError when the dict size does not match the number of attention layers."""
    count = len(get_attn_processors(module))
    if isinstance(processor, dict) and len(processor) != count:
        raise ValueError(
            f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
            f" number of attention layers: {count}. Please make sure to pass {count} processor classes.")

    def walk(prefix, mod):
        if hasattr(mod, "set_processor"):
            mod.set_processor(processor if not isinstance(processor, dict) else processor.pop(f"{prefix}.processor"))
        for name, child in mod.named_children():
            walk(f"{prefix}.{name}", child)

    for name, child in module.named_children():
        walk(name, child)
