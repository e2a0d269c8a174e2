"""Model registry with the call surface the reference gets from ``nowcasting_utils.models.base``
(reference call sites: satflow/models/conv_lstm.py:7,13; satflow/models/__init__.py:1;
tests/test_models.py:64-75).

``nowcasting_utils`` is an un-vendored third-party package (requirements.txt:22) whose source is
not under /root/reference, so this shim is pinned only by those call sites: ``register_model`` is a
bare class decorator that returns the class, keys are the lower-cased class name,
``create_model(name, pretrained=False, **kwargs)`` constructs, ``get_model(name)`` returns the
class, ``list_models()`` the sorted names.  If the real package is importable the class is
registered there as well, so ``satflow.models.create_model`` finds it.
"""
from __future__ import annotations

from typing import Dict, List, Type

_REGISTRY: Dict[str, Type] = {}


def register_model(cls):
    _REGISTRY[cls.__name__.lower()] = cls
    try:  # mirror into the real registry when it exists
        from nowcasting_utils.models.base import register_model as _real  # type: ignore

        _real(cls)
    except Exception:
        pass
    return cls


def get_model(name: str):
    key = name.lower()
    if key not in _REGISTRY:
        raise KeyError(f"unknown model {name!r}; registered: {list_models()}")
    return _REGISTRY[key]


def is_model(name: str) -> bool:
    return name.lower() in _REGISTRY


def list_models() -> List[str]:
    return sorted(_REGISTRY)


def create_model(model_name: str, pretrained: bool = False, checkpoint_path=None, **kwargs):
    """Same shape of call as the reference tests make (tests/test_models.py:66-75)."""
    if model_name.startswith("hf_hub:"):
        raise NotImplementedError("hf_hub checkpoints need network access; not part of the ConvLSTM hot path")
    model = get_model(model_name)(pretrained=pretrained, **kwargs)
    if checkpoint_path:
        import torch

        state = torch.load(checkpoint_path, map_location="cpu")
        model.load_state_dict(state.get("state_dict", state))
    return model
