"""CPU: the C-ABI library loads and exports every symbol include/clstm.h declares; host-side validation
and the drop-in Python surface (registry, constructors, state_dict layout) behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

import satflow_b200 as S
from satflow_b200 import _lib
from gpu_checks import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "clstm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clstm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.exported_symbols()) == declared
    assert L.clstm_abi_version() == 2


def _create(**kw):
    base = dict(batch=2, height=16, width=16, in_channels=12, hidden=32, out_channels=12, n_layers=2, kernel_h=3,
                kernel_w=3, t_in=4, t_out=4, dtype=0, training=1, grad_scale=0.0)
    base.update(kw)
    cfg = _lib.Config(*[base[f[0]] for f in _lib.Config._fields_])
    h = ctypes.c_void_p()
    rc = _lib.lib().clstm_plan_create(ctypes.byref(cfg), ctypes.byref(h))
    return rc, h


def test_plan_create_validates_without_a_gpu():
    L = _lib.lib()
    rc, h = _create()
    assert rc == 0 and h.value
    ws_train = L.clstm_plan_workspace_bytes(h)
    L.clstm_plan_destroy(h)
    rc, h = _create(training=0)
    ws_inf = L.clstm_plan_workspace_bytes(h)
    L.clstm_plan_destroy(h)
    assert 0 < ws_inf < ws_train
    for bad in (dict(t_out=0), dict(kernel_h=4), dict(hidden=1024), dict(batch=0), dict(dtype=7), dict(n_layers=0)):
        rc, _ = _create(**bad)
        assert rc == -1, bad
        assert L.clstm_last_error()
    rc, _ = _create(t_out=0)
    assert b"forecast_steps" in L.clstm_last_error()


def test_workspace_scales_with_saved_states():
    L = _lib.lib()
    sizes = []
    for t_out in (2, 4):
        rc, h = _create(t_out=t_out)
        sizes.append(L.clstm_plan_workspace_bytes(h))
        L.clstm_plan_destroy(h)
    npix, HP = 2 * 16 * 16, 64
    per_step = 2 * (npix * HP * 2 + npix * HP * 4 + npix * 4 * HP * 2)  # two decoder cells: h + c + gates
    assert sizes[1] - sizes[0] >= 2 * per_step
    assert sizes[1] - sizes[0] < 2 * per_step + 64 * 1024


def test_large_fp16_plans_keep_the_cell_state_in_16_bits(monkeypatch):
    """CLSTM_C16 (DESIGN.md finding 19) is decided when the plan is created: fp16 rollouts with more tiles than SMs and
    at most 200 steps halve their c stacks; small plans (persistent chain), bf16 plans and long rollouts keep fp32."""
    L = _lib.lib()

    def ws(**kw):
        cfg = _lib.Config(kw.get("batch", 16), 256, 256, 12, 64, 12, 2, 3, 3, kw.get("t_in", 12), kw.get("t_out", 24),
                          kw.get("dtype", _lib.CLSTM_F16), 1, 0.0)
        h = ctypes.c_void_p()
        assert L.clstm_plan_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
        n = L.clstm_plan_workspace_bytes(h)
        L.clstm_plan_destroy(h)
        return n

    npix, HP = 16 * 256 * 256, 64
    c_fp32 = (2 * 13 + 2 * 25) * npix * HP * 4  # 2 encoder + 2 decoder cells, T + 1 slots each
    monkeypatch.setenv("CLSTM_C16", "0")
    base = ws()
    monkeypatch.setenv("CLSTM_C16", "1")
    assert abs((base - ws()) - c_fp32 // 2) < (1 << 20)
    monkeypatch.setenv("CLSTM_C16", "0")
    base_bf16, base_long = ws(dtype=_lib.CLSTM_BF16), ws(batch=1, t_out=201)
    monkeypatch.setenv("CLSTM_C16", "1")
    assert ws(dtype=_lib.CLSTM_BF16) == base_bf16
    assert ws(batch=1, t_out=201) == base_long


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    assert _lib.lib().clstm_device_check(0) == -3
    m = S.ConvLSTM(12, 8, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 2, 12, 8, 8), 2)
    cell = S.ConvLSTMCell(3, 4, (3, 3), True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cell(torch.randn(1, 3, 8, 8), cell.init_hidden(1, (8, 8)))


def test_registry_mirrors_reference_contract():
    # reference tests/test_models.py:64-75: every registered name constructs with no kwargs
    assert "encoderdecoderconvlstm" in S.list_models()
    for name in S.list_models():
        assert S.create_model(name) is not None
    assert S.get_model("EncoderDecoderConvLSTM") is S.EncoderDecoderConvLSTM
    with pytest.raises(KeyError):
        S.get_model("nope")


def test_constructor_defaults_and_from_config():
    m = S.EncoderDecoderConvLSTM()
    assert (m.forecast_steps, m.lr, m.model.hidden_dim, m.model.input_channels, m.model.out_channels) == (48, 0.001, 64, 12, 1)
    assert isinstance(m.criterion, torch.nn.MSELoss)
    m2 = S.EncoderDecoderConvLSTM.from_config({"num_hidden": 8, "in_channels": 3})
    assert (m2.forecast_steps, m2.model.hidden_dim, m2.model.input_channels) == (1, 8, 3)  # conv_lstm.py:41 default 1
    assert isinstance(m.configure_optimizers(), torch.optim.Adam)
    with pytest.raises(ValueError):
        S.get_conv_layer("bogus")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 12, 4, 4), 0)  # future_seq = 0 raises like the reference


def test_state_dict_layout_matches_reference_fixture():
    z = load_golden("rollout_h16_12x10")
    ref = {k[len("param."):]: v for k, v in z.items() if k.startswith("param.")}
    m = S.EncoderDecoderConvLSTM(hidden_dim=16, input_channels=12, out_channels=5, forecast_steps=4)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    m.load_state_dict(ref)  # strict
    bare = S.ConvLSTM(12, 16, 5)
    assert list(bare.state_dict().keys()) == [k[len("model."):] for k in ref]
    cell = bare.encoder_1_convlstm
    assert (cell.input_dim, cell.hidden_dim, cell.kernel_size, cell.padding, cell.bias) == (12, 16, (3, 3), (1, 1), True)
    h, c = cell.init_hidden(2, (5, 7))
    assert h.shape == (2, 16, 5, 7) and float(h.abs().sum() + c.abs().sum()) == 0.0
    # CloudGAN's init_net matches sub-modules whose class name contains "Conv" and that own .weight (gan/common.py:46-64)
    assert "Conv" in type(cell.conv).__name__ and hasattr(cell.conv, "weight")


def test_yaml_and_lightning_checkpoint_helpers(tmp_path):
    # configs/model/convlstm.yaml has exactly these keys (reference defaults: 17 channels, forecast 24, lr 1e-4)
    y = tmp_path / "convlstm.yaml"
    y.write_text("# @package _group_\n_target_: satflow.models.conv_lstm.EncoderDecoderConvLSTM\ninput_channels: 17\n"
                 "hidden_dim: 64\nout_channels: 1\nforecast_steps: 24\nlr: 0.0001\nvisualize: True\n")
    m = S.EncoderDecoderConvLSTM.from_yaml(str(y))
    assert (m.model.input_channels, m.model.hidden_dim, m.forecast_steps, m.lr, m.visualize) == (17, 64, 24, 0.0001, True)
    z = load_golden("rollout_h16_12x10")
    sd = {k[len("param."):]: v for k, v in z.items() if k.startswith("param.")}
    ck = tmp_path / "best.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": {"hidden_dim": 16, "forecast_steps": 4}}, ck)
    m2 = S.EncoderDecoderConvLSTM(hidden_dim=16, input_channels=12, out_channels=5, forecast_steps=4)
    hp = m2.load_lightning_checkpoint(str(ck))
    assert hp["hidden_dim"] == 16
    assert torch.equal(m2.state_dict()["model.decoder_CNN.bias"], sd["model.decoder_CNN.bias"])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) prints ONE JSON line with the same
    metric / unit / config as the CUDA arm plus the reference-arm keys; it must work without a GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_lightning_format_checkpoint_round_trip(tmp_path):
    """tests/golden/lightning_best.ckpt was written by the REFERENCE class in the format Lightning's
    ModelCheckpoint(save_weights_only=True) uses (state_dict + hyper_parameters of save_hyperparameters,
    conv_lstm.py:33).  load_from_checkpoint rebuilds the module from it; lightning_checkpoint() writes the same format
    back, bit for bit."""
    path = os.path.join(ROOT, "tests", "golden", "lightning_best.ckpt")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    m = S.EncoderDecoderConvLSTM.load_from_checkpoint(path)
    hp = ck["hyper_parameters"]
    assert (m.model.hidden_dim, m.model.input_channels, m.model.out_channels, m.forecast_steps, m.lr) == (
        hp["hidden_dim"], hp["input_channels"], hp["out_channels"], hp["forecast_steps"], hp["lr"])
    assert dict(m.hparams) == hp
    for k, v in ck["state_dict"].items():
        assert torch.equal(m.state_dict()[k], v), k
    out = tmp_path / "resaved.ckpt"
    torch.save(m.lightning_checkpoint(epoch=ck["epoch"], global_step=ck["global_step"]), out)
    ck2 = torch.load(out, map_location="cpu", weights_only=False)
    assert list(ck2["state_dict"].keys()) == list(ck["state_dict"].keys())
    assert all(torch.equal(ck2["state_dict"][k], ck["state_dict"][k]) for k in ck["state_dict"])
    assert ck2["hyper_parameters"] == hp and ck2["epoch"] == 3 and ck2["global_step"] == 1234
    # keyword overrides win over the stored hyper-parameters, like Lightning's load_from_checkpoint(**kwargs)
    assert S.EncoderDecoderConvLSTM.load_from_checkpoint(path, forecast_steps=2).forecast_steps == 2


@pytest.mark.reference
def test_reference_class_loads_a_checkpoint_written_here(tmp_path):
    """The other direction: a checkpoint written by this package loads strictly into the unmodified reference class and
    reproduces the output stored in the fixture bit for bit (CPU, build container only)."""
    from oracle.reference_loader import load_reference

    path = os.path.join(ROOT, "tests", "golden", "lightning_best.ckpt")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    ours = S.EncoderDecoderConvLSTM.load_from_checkpoint(path)
    out = tmp_path / "from_b200.ckpt"
    torch.save(ours.lightning_checkpoint(), out)
    ck2 = torch.load(out, map_location="cpu", weights_only=False)
    _, _, Lit = load_reference()
    ref = Lit(**ck2["hyper_parameters"])
    ref.load_state_dict(ck2["state_dict"])  # strict
    with torch.no_grad():
        y = ref(ck["golden"]["x"], ck2["hyper_parameters"]["forecast_steps"])
    assert torch.equal(y, ck["golden"]["y"])
