"""ConvLSTM encoder-forecaster with the reference's registry entry, constructors, forward
signature and state_dict layout (satflow/models/conv_lstm.py:13-228), executed as ONE rollout on
the B200 through the C ABI (clstm_rollout_forward / clstm_rollout_backward).

The four cells and the Conv3d head are real torch modules used as parameter holders only — their
``forward`` is never called on the rollout path.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from .layers import ConvLSTMCell
from .plan import RolloutPlan
from .registry import register_model

try:  # the reference derives from LightningModule (conv_lstm.py:14); use it when it is installed
    import pytorch_lightning as _pl  # type: ignore

    _Base = _pl.LightningModule
except Exception:  # pragma: no cover - pytorch_lightning is absent in this image

    class _Base(nn.Module):
        """Minimal stand-in with the hooks EncoderDecoderConvLSTM uses."""

        def save_hyperparameters(self, *args, **kwargs):
            import inspect

            frame = inspect.currentframe().f_back
            names = frame.f_code.co_varnames[1 : frame.f_code.co_argcount]
            self.hparams = {n: frame.f_locals[n] for n in names}

        def log(self, name, value, *args, **kwargs):
            if not hasattr(self, "logged"):
                self.logged = {}
            self.logged[name] = value

        def log_dict(self, d, *args, **kwargs):
            for k, v in d.items():
                self.log(k, v)


def get_loss(loss="mse", **kwargs):
    """conv_lstm.py:8,29 take this from nowcasting_utils.models.loss (un-vendored); "mse" is the only
    string the ConvLSTM configs use (configs/model/convlstm.yaml), a Module passes through."""
    if isinstance(loss, nn.Module):
        return loss
    try:
        from nowcasting_utils.models.loss import get_loss as _real  # type: ignore

        return _real(loss, **kwargs)
    except ImportError:
        pass
    if loss in ("mse", "l2"):
        return nn.MSELoss()
    if loss in ("l1", "mae"):
        return nn.L1Loss()
    raise ValueError(f"loss {loss!r} needs nowcasting_utils, which is not installed")


class _RolloutFn(torch.autograd.Function):
    """One autograd node for the whole encoder-forecaster rollout (not one per cell)."""

    @staticmethod
    def forward(ctx, plan: RolloutPlan, x: torch.Tensor, channels_last: bool, *params: torch.Tensor):
        plan.set_weights(params)
        y = plan.forward(x, channels_last=channels_last)
        ctx.plan = plan
        ctx.generation = plan.generation  # the saved gates / states live in the plan's workspace, not in ctx
        ctx.n_params = len(params)
        ctx.save_for_backward(y, *params)
        return y

    @staticmethod
    def backward(ctx, dy):
        y = ctx.saved_tensors[0]
        params = ctx.saved_tensors[1:]
        grads = [torch.empty_like(p) if ctx.needs_input_grad[3 + i] else None for i, p in enumerate(params)]
        ctx.plan.backward(dy, y, grads, accumulate=False, generation=ctx.generation)
        return (None, None, None, *grads)


class ConvLSTM(nn.Module):
    """conv_lstm.py:121-228.  ``n_layers`` / ``kernel_size`` / ``operand_dtype`` are extensions with
    reference defaults (2 encoder + 2 decoder cells, 3x3): the extra cells are named
    ``encoder_{l}_convlstm`` / ``decoder_{l}_convlstm`` following the reference's pattern."""

    def __init__(self, input_channels, hidden_dim, out_channels, conv_type: str = "standard", n_layers: int = 2,
                 kernel_size=(3, 3), operand_dtype: str = "fp16"):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_dim = hidden_dim
        self.out_channels = out_channels
        self.n_layers = n_layers
        self.kernel_size = tuple(kernel_size)
        self.operand_dtype = operand_dtype
        for l in range(n_layers):  # conv_lstm.py:128-145
            setattr(
                self,
                f"encoder_{l + 1}_convlstm",
                ConvLSTMCell(input_channels if l == 0 else hidden_dim, hidden_dim, self.kernel_size, True, conv_type),
            )
        for l in range(n_layers):  # conv_lstm.py:147-162
            setattr(self, f"decoder_{l + 1}_convlstm", ConvLSTMCell(hidden_dim, hidden_dim, self.kernel_size, True, conv_type))
        self.decoder_CNN = nn.Conv3d(  # conv_lstm.py:164-169
            in_channels=hidden_dim, out_channels=out_channels, kernel_size=(1, 3, 3), padding=(0, 1, 1)
        )
        self._plans: Dict[tuple, RolloutPlan] = {}

    # ---- plumbing ---------------------------------------------------------------------------
    def cells(self) -> List[ConvLSTMCell]:
        names = [f"encoder_{l + 1}_convlstm" for l in range(self.n_layers)]
        names += [f"decoder_{l + 1}_convlstm" for l in range(self.n_layers)]
        return [getattr(self, n) for n in names]

    def rollout_params(self) -> List[torch.Tensor]:
        """Parameter order of clstm_plan_set_weights (include/clstm.h)."""
        out: List[torch.Tensor] = []
        for cell in self.cells():
            out += [cell.conv.weight, cell.conv.bias]
        out += [self.decoder_CNN.weight, self.decoder_CNN.bias]
        return out

    def plan_for(self, x: torch.Tensor, forecast_steps: int, training: bool, channels_last: bool = False) -> RolloutPlan:
        if channels_last:
            b, seq_len, h, w, c = x.shape
        else:
            b, seq_len, c, h, w = x.shape
        key = (b, seq_len, c, h, w, forecast_steps, training, self.operand_dtype, x.device.index)

        def make():
            return RolloutPlan(
                b, h, w, c, self.hidden_dim, self.out_channels, seq_len, forecast_steps, self.n_layers,
                self.kernel_size, self.operand_dtype, training, 0.0, x.device,
            )

        plan = self._plans.get(key)
        if plan is not None and plan.pending_backward:
            # A graph built on this plan has not run its backward yet (two batches summed into one loss, a GAN /
            # consistency loss, a grad-enabled evaluation pass): its saved states must survive, so this forward gets a
            # second plan of the same shape.  If that does not fit in HBM the cached plan is reused; the older graph's
            # backward then raises (generation mismatch) instead of returning gradients of the wrong forward.
            key2 = key + ("second",)
            plan2 = self._plans.get(key2)
            if plan2 is None:
                try:
                    plan2 = make()
                    self._plans[key2] = plan2
                except torch.OutOfMemoryError:
                    plan2 = None
            if plan2 is not None and not plan2.pending_backward:
                return plan2
            return plan if plan2 is None or plan.last_use <= plan2.last_use else plan2  # sacrifice the oldest graph
        if plan is None:
            # bounded cache (a training plan pins tens of GB); plans with an outstanding backward are never evicted
            evictable = [k for k, pl in self._plans.items() if not pl.pending_backward]
            while len(self._plans) >= 3 and evictable:
                self._plans.pop(evictable.pop(0)).close()
            plan = make()
            self._plans[key] = plan
        return plan

    def release_plans(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()

    def check_gradients(self):
        """Wait for the backwards issued so far and raise ``FloatingPointError`` if a 16-bit gradient operand
        overflowed in one of them (the same check runs without waiting at the start of every later forward).
        Returns the range statistics of the newest finished backward (see RolloutPlan.poll_grad_status)."""
        out = None
        for p in self._plans.values():
            if p.training:
                st = p.poll_grad_status(block=True)
                if st is not None and (out is None or st["generation"] >= out["generation"]):
                    out = st
        return out

    # ---- reference API ----------------------------------------------------------------------
    def forward_channels_last(self, x, forecast_steps=0):
        """The same rollout on a channels-last batch (B, T_in, H, W, C) — the layout the reference's datasets emit
        (data/datasets.py:70-106) — consumed without a permute copy."""
        return self.forward(x, forecast_steps, channels_last=True)

    def forward(self, x, forecast_steps=0, hidden_state=None, channels_last: bool = False):
        """x: (B, T_in, C, H, W) float32 CUDA -> (B, out_channels, T_out, H, W)   (conv_lstm.py:205-228).
        ``hidden_state`` is accepted and ignored exactly like the reference (:205 never reads it)."""
        if x.dim() != 5:
            raise RuntimeError(f"ConvLSTM expects a 5-D (b, t, c, h, w) tensor, got {tuple(x.shape)}")
        if x.shape[4 if channels_last else 2] != self.input_channels:
            raise RuntimeError(f"expected {self.input_channels} input channels, got {x.shape[4 if channels_last else 2]}")
        if forecast_steps <= 0:
            # the reference reaches torch.stack([]) at conv_lstm.py:198
            raise RuntimeError("stack expects a non-empty TensorList")
        if not x.is_cuda:
            raise RuntimeError(
                f"satflow_b200.ConvLSTM runs on a B200 only: x is on {x.device}; there is no CPU path "
                "(move the module and its input to cuda)"
            )
        params = self.rollout_params()
        training = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        plan = self.plan_for(x, int(forecast_steps), training, channels_last)
        if x.dtype != torch.float32:
            x = x.float()
        if training and x.requires_grad:
            # the reference's flows never differentiate through the input sequence, and encoder_1's data gradient is not
            # computed here; returning None silently would be a wrong gradient for whoever asked for it
            raise RuntimeError(
                "satflow_b200.ConvLSTM does not produce the gradient w.r.t. its input sequence x; detach() it "
                "(parameter gradients are unaffected)"
            )
        return _RolloutFn.apply(plan, x, bool(channels_last), *params)


@register_model
class EncoderDecoderConvLSTM(_Base):
    """conv_lstm.py:13-118 (registry entry "encoderdecoderconvlstm")."""

    def __init__(
        self,
        hidden_dim: int = 64,
        input_channels: int = 12,
        out_channels: int = 1,
        forecast_steps: int = 48,
        lr: float = 0.001,
        visualize: bool = False,
        loss="mse",
        pretrained: bool = False,
        conv_type: str = "standard",
    ):
        super().__init__()
        self.forecast_steps = forecast_steps
        self.criterion = get_loss(loss)
        self.lr = lr
        self.visualize = visualize
        self.model = ConvLSTM(input_channels, hidden_dim, out_channels, conv_type=conv_type)
        self.save_hyperparameters()

    @classmethod
    def from_config(cls, config):
        return EncoderDecoderConvLSTM(  # conv_lstm.py:35-43 (note forecast_steps default 1 here)
            hidden_dim=config.get("num_hidden", 64),
            input_channels=config.get("in_channels", 12),
            out_channels=config.get("out_channels", 1),
            forecast_steps=config.get("forecast_steps", 1),
            lr=config.get("lr", 0.001),
        )

    @classmethod
    def from_yaml(cls, path):
        """Build from a Hydra model YAML of the reference's shape (configs/model/convlstm.yaml: ``_target_`` plus
        constructor kwargs); unknown keys are rejected like a wrong kwarg would be."""
        import yaml

        with open(path) as f:
            cfg = yaml.safe_load(f) or {}
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        return cls(**cfg)

    def load_lightning_checkpoint(self, ckpt, strict: bool = True):
        """Load a reference Lightning checkpoint (path or dict): ``state_dict`` has the keys of SURVEY.md §8(b);
        ``hyper_parameters`` (from save_hyperparameters, conv_lstm.py:33) are returned for inspection."""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
        self.load_state_dict(ckpt.get("state_dict", ckpt), strict=strict)
        return ckpt.get("hyper_parameters", {})

    def lightning_checkpoint(self, epoch: int = 0, global_step: int = 0) -> dict:
        """The dict Lightning's ModelCheckpoint(save_weights_only=True) writes for this module
        (configs/callbacks/default.yaml:1-17): reference state_dict keys + ``hyper_parameters``.  ``torch.save`` it
        and the reference class loads it (``load_state_dict(ckpt["state_dict"])`` strict, or ``load_from_checkpoint``)."""
        return {"epoch": epoch, "global_step": global_step, "pytorch-lightning_version": "1.4.9",
                "state_dict": {k: v.detach().cpu().clone() for k, v in self.state_dict().items()},
                "hyper_parameters": dict(self.hparams)}

    @classmethod
    def load_from_checkpoint(cls, ckpt, map_location=None, strict: bool = True, **overrides):
        """Lightning's ``LightningModule.load_from_checkpoint`` for this class: construct from the checkpoint's
        ``hyper_parameters`` (keyword overrides win) and load its ``state_dict``."""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location=map_location or "cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
        names = ("hidden_dim", "input_channels", "out_channels", "forecast_steps", "lr", "visualize", "loss", "pretrained",
                 "conv_type")
        model = cls(**{k: v for k, v in hp.items() if k in names})
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model

    def forward(self, x, future_seq=0, hidden_state=None):
        return self.model.forward(x, future_seq, hidden_state)

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.lr)  # conv_lstm.py:48-51

    def _frame_losses(self, y_hat, y, prefix):
        # conv_lstm.py:66-69: the reference does one .item() host sync per frame; here the per-frame
        # means are reduced on the device and fetched with a single copy.
        with torch.no_grad():
            if isinstance(self.criterion, nn.MSELoss):
                per = ((y_hat - y) ** 2).mean(dim=(0, 2, 3, 4))
            else:
                per = torch.stack([self.criterion(y_hat[:, f], y[:, f]) for f in range(y_hat.shape[1])])
            vals = per.tolist()
        return {f"{prefix}/frame_{f}_loss": v for f, v in enumerate(vals)}

    def loss_and_frame_losses(self, out, y):
        """out: forward output (B, C, T, H, W); y: target (B, T, C, H, W).  MSE (the reference default) runs as one
        fused kernel producing loss, gradient and the per-frame losses of conv_lstm.py:66-69 without host syncs."""
        if isinstance(self.criterion, nn.MSELoss) and out.is_cuda and self.criterion.reduction == "mean":
            from .loss import fused_mse

            return fused_mse(out, y)
        y_hat = torch.permute(out, dims=(0, 2, 1, 3, 4))  # conv_lstm.py:56
        loss = self.criterion(y_hat, y)
        with torch.no_grad():
            frames = torch.stack([self.criterion(y_hat[:, f], y[:, f]) for f in range(y_hat.shape[1])])
        return loss, frames

    def training_step(self, batch, batch_idx):
        x, y = batch
        out = self(x, self.forecast_steps)
        loss, frames = self.loss_and_frame_losses(out, y)
        self.log("train/loss", loss, on_step=True)
        # conv_lstm.py:65-69: the reference log_dict()s one python float per frame, each fetched with its own .item()
        # host sync.  Here the T_out per-frame means come out of the fused loss kernel as ONE device tensor and are
        # logged as 0-d views of it (Lightning accepts tensors and reduces them over the epoch on the device), so no
        # synchronisation happens inside the step; ``frame_losses`` keeps the whole vector for a single D2H copy.
        self.frame_losses = frames.detach()
        self.log_dict({f"train/frame_{f}_loss": self.frame_losses[f] for f in range(self.frame_losses.shape[0])},
                      on_step=False, on_epoch=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        y_hat = self(x, self.forecast_steps)
        y_hat = torch.permute(y_hat, dims=(0, 2, 1, 3, 4))
        val_loss = self.criterion(y_hat, y)
        self.log("val/loss", val_loss, on_step=True, on_epoch=True)
        self.log_dict(self._frame_losses(y_hat, y, "val"))
        return val_loss

    def test_step(self, batch, batch_idx):
        x, y = batch
        y_hat = self(x, self.forecast_steps)
        y_hat = torch.permute(y_hat, dims=(0, 2, 1, 3, 4))  # the reference forgets this (conv_lstm.py:89-90)
        loss = self.criterion(y_hat, y)
        return loss
