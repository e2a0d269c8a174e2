"""satflow_b200 — B200-native ConvLSTM encoder-forecaster hot path of openclimatefix/satflow.

Drop-in surface: ``ConvLSTMCell``, ``ConvLSTM``, ``EncoderDecoderConvLSTM`` and the
``register_model / get_model / create_model / list_models`` registry.  All compute goes through
``libclstm.so`` (hand-written sm_100a kernels behind the C ABI of include/clstm.h); there is no
CPU or eager fallback.
"""
from .registry import create_model, get_model, is_model, list_models, register_model  # noqa: F401
from .layers import ConvLSTMCell, NativeCellStepper, get_conv_layer  # noqa: F401
from .conv_lstm import ConvLSTM, EncoderDecoderConvLSTM, get_loss  # noqa: F401
from .plan import CellPlan, RolloutPlan  # noqa: F401
from .loss import fused_mse  # noqa: F401

__version__ = "0.1.0"
