"""not-gpu: the .pack checkpoint format (dqn/network.py:27-47) stays byte-compatible."""
import hashlib
import os

import numpy as np
import pytest
import torch

from multimodal_drl_rmc_b200 import Networks, packfmt
from multimodal_drl_rmc_b200.macro_config import ObsSpace, network_config
from oracle import refharness
from tests import recipes as R

PACK = os.path.join(R.GOLDEN_DIR, "macro_with_lane.pack")


def test_codec_roundtrip_is_byte_identical():
    raw = open(PACK, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == R.golden_meta()["pack_sha"]
    obj = packfmt.loads(raw)
    assert list(obj) == ["parameters", "step", "episode_count", "rew_mean", "len_mean"]
    assert list(obj["parameters"]) == ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "fc_val.weight",
                                       "fc_val.bias", "fc_adv.weight", "fc_adv.bias"]
    assert obj["parameters"]["net.0.weight"].shape == (256, 14) and obj["parameters"]["net.0.weight"].dtype == np.float32
    assert packfmt.dumps(obj) == raw


def test_network_load_save_reproduces_shipped_file(tmp_path):
    net = Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, network_config, ObsSpace(14), 8)
    step, episodes, rew_mean, len_mean = net.load(PACK)
    assert (step, episodes) == (2000000, 22222) and abs(len_mean - 90.0) < 1e-9
    out = str(tmp_path / "sub" / "m.pack")
    net.save(out, step, episodes, rew_mean, len_mean)
    assert open(out, "rb").read() == open(PACK, "rb").read()
    with pytest.raises(FileNotFoundError):
        net.load(str(tmp_path / "missing.pack"))


def test_state_dict_keys_and_shapes_match_reference_layout():
    net = Networks.DeepQNetwork(torch.device("cpu"), 1e-4, network_config, ObsSpace(8), 8)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == [
        ("net.0.weight", (256, 8)), ("net.0.bias", (256,)), ("net.2.weight", (128, 256)), ("net.2.bias", (128,)),
        ("fc_out.weight", (8, 128)), ("fc_out.bias", (8,))]
    assert hasattr(net, "optimizer") and hasattr(net, "loss") and net.fc_out_dim == 128


@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree only exists in the build container")
def test_cross_load_with_the_reference_implementation(tmp_path):
    dqn = refharness.import_reference()
    ours = Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, network_config, ObsSpace(14), 8)
    path = str(tmp_path / "ours" / "x.pack")
    ours.save(path, 123, 4, np.float64(1.5), np.float64(90.0))
    ref = dqn.Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, refharness.macro_network_config, refharness.ObsBox(14), 8)
    assert ref.load(path) == (123, 4, 1.5, 90.0)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    path2 = str(tmp_path / "ref" / "y.pack")
    ref.save(path2, 123, 4, np.float64(1.5), np.float64(90.0))
    assert open(path, "rb").read() == open(path2, "rb").read()


def test_hybrid_state_dict_layout():
    """The hybrid body's parameter names/shapes are those of the reference's shipped hybrid checkpoints."""
    from multimodal_drl_rmc_b200.macro_config import network_config_hybrid, HYBRID_OBS_DIM
    net = Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, network_config_hybrid, ObsSpace(HYBRID_OBS_DIM), 8)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert got == [("net.cnn_stream.0.weight", (32, 2, 3, 3)), ("net.cnn_stream.0.bias", (32,)),
                   ("net.cnn_stream.2.weight", (64, 32, 3, 3)), ("net.cnn_stream.2.bias", (64,)),
                   ("net.cnn_stream.4.weight", (64, 64, 3, 3)), ("net.cnn_stream.4.bias", (64,)),
                   ("net.dense_stream.0.weight", (512, 1358)), ("net.dense_stream.0.bias", (512,)),
                   ("net.dense_stream.2.weight", (256, 512)), ("net.dense_stream.2.bias", (256,)),
                   ("fc_val.weight", (1, 256)), ("fc_val.bias", (1,)), ("fc_adv.weight", (8, 256)), ("fc_adv.bias", (8,))]
    assert sum(v.numel() for v in net.state_dict().values()) == 885481 and net._hybrid is not None


@pytest.mark.skipif(not refharness.reference_available(), reason="reference tree only exists in the build container")
def test_shipped_hybrid_checkpoint_loads_and_resaves_byte_identically(tmp_path):
    """save/1ramp_1x3/DuelingDoubleDQNAgent_lr0.0001_model.pack (3.5 MB, the repo-HEAD network) through our Network."""
    from multimodal_drl_rmc_b200.macro_config import network_config_hybrid, HYBRID_OBS_DIM
    src = os.path.join(refharness.REFERENCE_ROOT, "save", "1ramp_1x3", "DuelingDoubleDQNAgent_lr0.0001_model.pack")
    if not os.path.exists(src):
        pytest.skip("shipped hybrid checkpoint not present")
    net = Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, network_config_hybrid, ObsSpace(HYBRID_OBS_DIM), 8)
    meta = net.load(src)
    out = str(tmp_path / "h.pack")
    net.save(out, *meta)
    assert open(out, "rb").read() == open(src, "rb").read()
