"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference classes from /root/reference.

The reference's hot-path modules (``satflow/models/conv_lstm.py``,
``satflow/models/layers/ConvLSTM.py``, ``satflow/models/utils.py``) import two third-party
packages that are absent here (``pytorch_lightning``, ``nowcasting_utils``) and their
package ``__init__`` files eagerly import unrelated model families.  This loader stubs
the two absent packages and pre-seeds bare package objects so that the three files above
are executed exactly as they lie on disk — no reference source is copied or edited.

It only works where ``/root/reference`` exists (the build container).  It is used to
(1) pin ``oracle/convlstm_oracle.py`` against the real reference and (2) generate the
golden fixtures under ``tests/golden/`` (see ``oracle/make_golden.py``).  Nothing that
runs on the GPU box may call it.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SATFLOW_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "satflow", "models", "conv_lstm.py"))


def _install_stubs() -> None:
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl

    if "nowcasting_utils" not in sys.modules:
        registry = {}

        def register_model(cls):
            registry[cls.__name__.lower()] = cls
            return cls

        def get_model(name):
            return registry[name.lower()]

        def create_model(name, pretrained=False, **kwargs):
            return registry[name.lower()](**kwargs)

        def list_models():
            return sorted(registry)

        def get_loss(loss="mse", **kwargs):
            if isinstance(loss, nn.Module):
                return loss
            assert loss in ("mse", "l2"), loss
            return nn.MSELoss()

        nu = types.ModuleType("nowcasting_utils")
        nu_models = types.ModuleType("nowcasting_utils.models")
        nu_base = types.ModuleType("nowcasting_utils.models.base")
        nu_loss = types.ModuleType("nowcasting_utils.models.loss")
        for fn in (register_model, get_model, create_model, list_models):
            setattr(nu_base, fn.__name__, fn)
        nu_base._registry = registry
        nu_loss.get_loss = get_loss
        nu.models = nu_models
        nu_models.base = nu_base
        nu_models.loss = nu_loss
        sys.modules.update(
            {
                "nowcasting_utils": nu,
                "nowcasting_utils.models": nu_models,
                "nowcasting_utils.models.base": nu_base,
                "nowcasting_utils.models.loss": nu_loss,
            }
        )


def load_reference():
    """Returns (ConvLSTMCell, ConvLSTM, EncoderDecoderConvLSTM) — the reference's own classes."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    sat = os.path.join(REFERENCE_ROOT, "satflow")
    if "satflow" not in sys.modules or not hasattr(sys.modules["satflow"], "_oracle_seeded"):
        for name, path in (
            ("satflow", sat),
            ("satflow.models", os.path.join(sat, "models")),
            ("satflow.models.layers", os.path.join(sat, "models", "layers")),
        ):
            mod = types.ModuleType(name)
            mod.__path__ = [path]
            mod._oracle_seeded = True
            sys.modules[name] = mod
        coord = importlib.import_module("satflow.models.layers.CoordConv")
        sys.modules["satflow.models.layers"].CoordConv = coord.CoordConv
    cell_mod = importlib.import_module("satflow.models.layers.ConvLSTM")
    model_mod = importlib.import_module("satflow.models.conv_lstm")
    return cell_mod.ConvLSTMCell, model_mod.ConvLSTM, model_mod.EncoderDecoderConvLSTM


if __name__ == "__main__":
    Cell, Net, Lit = load_reference()
    torch.manual_seed(0)
    m = Lit(hidden_dim=32, input_channels=12, out_channels=12, forecast_steps=4)
    for k, v in m.state_dict().items():
        print(k, tuple(v.shape))
    y = m(torch.randn(2, 4, 12, 16, 16), 4)
    print(tuple(y.shape))
