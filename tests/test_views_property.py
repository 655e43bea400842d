"""Property tests of the host-side view algebra (BackendTensor on the oracle's numpy device): random chains of
permute / slice / broadcast / flip / reshape / setitem / reductions must agree with numpy on the same data. These are
the strided views that every device kernel call is built from (reference: backend_tensor.py:320-528)."""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

F32 = np.float32
SET = dict(max_examples=120, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


def _array(shape, seed):
    return np.random.RandomState(seed).randint(-50, 50, size=shape).astype(F32)


shapes = st.lists(st.integers(1, 5), min_size=1, max_size=4).map(tuple)


@st.composite
def slice_for(draw, n):
    if draw(st.booleans()):
        return draw(st.integers(0, n - 1))                      # an integer index KEEPS the axis (bt.py:491-496)
    start = draw(st.integers(0, n - 1))
    stop = draw(st.integers(start + 1, n))                      # empty slices assert in the reference (bt.py:455)
    step = draw(st.integers(1, 3))
    return slice(start, stop, step)


def as_numpy_index(idx):
    return tuple(slice(i, i + 1) if isinstance(i, int) else i for i in idx)


@settings(**SET)
@given(shape=shapes, seed=st.integers(0, 1000), data=st.data())
def test_chain_of_views_matches_numpy(cpu_device, shape, seed, data):
    from DeepFlows import backend_api
    a = _array(shape, seed)
    t = backend_api.Btensor(a, device=cpu_device)
    for _ in range(data.draw(st.integers(1, 4))):
        op = data.draw(st.sampled_from(["permute", "slice", "broadcast", "flip", "reshape"]))
        if op == "permute" and a.ndim > 1:
            perm = tuple(data.draw(st.permutations(range(a.ndim))))
            a, t = a.transpose(perm), t.permute(perm)
        elif op == "slice":
            idx = tuple(data.draw(slice_for(n)) for n in a.shape)
            a, t = a[as_numpy_index(idx)], t[idx]
        elif op == "broadcast" and a.ndim < 4:
            lead = data.draw(st.integers(1, 3))
            a, t = np.broadcast_to(a, (lead,) + a.shape), t.broadcast_to((lead,) + t.shape)
        elif op == "flip" and a.ndim > 0:
            axes = tuple(sorted(set(data.draw(st.lists(st.integers(0, a.ndim - 1), min_size=1, max_size=a.ndim)))))
            a, t = np.flip(a, axes), t.flip(axes)
        elif op == "reshape":
            a, t = np.ascontiguousarray(a).reshape(-1), t.compact().reshape((t.size,))
        assert t.shape == a.shape
    assert np.array_equal(t.numpy(), a)
    assert np.array_equal(t.compact().numpy(), a)
    assert np.array_equal((t + 1.0).numpy(), a + 1)              # element-wise ops see the same view


@settings(**SET)
@given(shape=shapes, seed=st.integers(0, 1000), data=st.data())
def test_reductions_match_numpy(cpu_device, shape, seed, data):
    from DeepFlows import backend_api
    a = _array(shape, seed)
    t = backend_api.Btensor(a, device=cpu_device)
    if a.ndim > 1 and data.draw(st.booleans()):
        perm = tuple(data.draw(st.permutations(range(a.ndim))))
        a, t = a.transpose(perm), t.permute(perm)
    axis = data.draw(st.integers(0, a.ndim - 1))
    keep = data.draw(st.booleans())
    assert np.array_equal(t.sum(axis=axis, keepdims=keep).numpy(), a.sum(axis=axis, keepdims=keep))   # small integers: exact
    assert np.array_equal(t.max(axis=axis, keepdims=keep).numpy(), a.max(axis=axis, keepdims=keep))
    # reference quirk Q3: mean over one axis still divides by the total element count
    assert np.allclose(t.mean(axis=axis, keepdims=keep).numpy(), a.sum(axis=axis, keepdims=keep) / a.size, rtol=1e-6)


@settings(**SET)
@given(shape=shapes, seed=st.integers(0, 1000), data=st.data())
def test_setitem_matches_numpy(cpu_device, shape, seed, data):
    from DeepFlows import backend_api
    a = _array(shape, seed)
    t = backend_api.Btensor(a.copy(), device=cpu_device)
    idx = tuple(data.draw(slice_for(n)) for n in a.shape)
    region = a[as_numpy_index(idx)]
    if data.draw(st.booleans()):
        a[as_numpy_index(idx)] = 7.0
        t[idx] = 7.0
    else:
        src = _array(region.shape, seed + 1)
        a[as_numpy_index(idx)] = src
        t[idx] = backend_api.Btensor(src, device=cpu_device)
    assert np.array_equal(t.numpy(), a)
