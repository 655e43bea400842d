"""Transfer-learning flow on the host package, numpy device, against the fixture the reference
produced (oracle/make_golden_transfer.py): `load_weights` of a partial dict, frozen stem / first stage, Adam over the
trainable parameters only, two steps."""
import numpy as np
import pytest

import parity
import workloads


def test_transfer_learning_matches_reference(cpu_device):
    parity.check_transfer("cpu")


def test_state_dict_round_trip_and_strict_errors(cpu_device):
    from DeepFlows import tensor
    model, weights, g = parity.check_transfer("cpu")
    # strict loading of the partial dict: same complaint as the reference (module.py:527-535)
    with pytest.raises(RuntimeError) as err:
        model.load_state_dict(weights, strict=True)
    assert "fc.weight" in str(err.value) and "Missing key(s)" in str(err.value)
    assert str(g["strict_error"]).splitlines()[0] == str(err.value).splitlines()[0]
    with pytest.raises(RuntimeError) as err:
        model.load_state_dict(dict(model.state_dict(), bogus=np.zeros(1, np.float32)), strict=True)
    assert "Unexpected key(s) in state_dict: bogus." in str(err.value)
    # a parameters-only dict (what the reference's modules hold: its BatchNorm statistics are not buffers) loads strictly
    model.load_state_dict({k: p.data.numpy() for k, p in model.named_parameters()}, strict=True)
    # full round trip, parameters and BatchNorm running statistics
    state = model.state_dict()
    tensor.Graph.free_graph_all()
    np.random.seed(77)
    df = parity.df_namespace()
    twin = workloads.resnet_cifar(df, "cpu", widths=(4, 8, 8, 16), layers=(1, 1, 1, 1))
    twin.load_state_dict(state, strict=True)
    for (k, a), (_, b) in zip(model.state_dict().items(), twin.state_dict().items()):
        assert np.array_equal(a, b), k
    assert any(k.endswith("running_mean") for k in state)
