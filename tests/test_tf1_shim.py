"""Self-checks of oracle/tf1_shim (the eager stand-in for TensorFlow 1.x that tests/golden/make_reference_vectors.py runs
the reference's own modules on): each op whose semantics matter for the golden vectors against an explicit loop or a
closed form that does not go through the same torch call."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("tf1_shim_under_test", os.path.join(ROOT, "oracle", "tf1_shim", "tensorflow", "__init__.py"))
tf = importlib.util.module_from_spec(_spec)  # not registered as `tensorflow`: other packages probe for that name
_spec.loader.exec_module(tf)


def test_dilated_convolution_is_a_valid_cross_correlation():
    rng = np.random.default_rng(0)
    x, w = rng.standard_normal((2, 11, 3)), rng.standard_normal((2, 3, 4))
    for dil in (1, 2, 4):
        y = tf.nn.convolution(torch.as_tensor(x), torch.as_tensor(w), "VALID", [1], [dil]).numpy()
        assert y.shape == (2, 11 - dil, 4)
        for t in range(11 - dil):
            want = x[:, t, :] @ w[0] + x[:, t + dil, :] @ w[1]
            assert np.allclose(y[:, t, :], want, atol=1e-12)


def test_conv1d_transpose_with_width_equal_to_stride():
    rng = np.random.default_rng(1)
    x, w = rng.standard_normal((2, 5, 3)), rng.standard_normal((4, 6, 3))  # filter [width, out, in]
    y = tf.contrib.nn.conv1d_transpose(torch.as_tensor(x), torch.as_tensor(w), [2, 20, 6], 4).numpy()
    for t in range(5):
        for k in range(4):
            assert np.allclose(y[:, t * 4 + k, :], x[:, t, :] @ w[k].T, atol=1e-12)


def test_one_hot_and_the_fused_xent_gradient():
    oh = tf.one_hot(torch.tensor([[0, 3, -1, 4]]), 4).numpy()
    assert oh.shape == (1, 4, 4) and oh[0, 0, 0] == 1 and oh[0, 1, 3] == 1 and not oh[0, 2].any() and not oh[0, 3].any()
    logits = torch.tensor([[0.5, -1.0, 2.0], [0.1, 0.2, 0.3]], dtype=torch.float64, requires_grad=True)
    labels = torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
    loss = tf.nn.softmax_cross_entropy_with_logits_v2(labels=labels, logits=logits, dim=1)
    sm = torch.softmax(logits.detach(), 1)
    assert np.allclose(loss.detach().numpy(), [-(np.log(sm[0, 1])), 0.0])
    loss.sum().backward()
    # grad = softmax - labels, also for the all-zero label row (xent_op.h backprop, nn_grad.py)
    assert np.allclose(logits.grad.numpy(), (sm - labels).numpy())


def test_variables_scopes_and_integer_mean():
    tf._reset()
    with tf.variable_scope("a"):
        v = tf.get_variable("w", [2, 3])
        with tf.name_scope("ignored"):
            assert tf.get_variable("w", [2, 3]) is v  # reuse by scoped name; name_scope does not prefix variables
    assert v.name == "a/w:0" and v.trainable and tuple(v.shape) == (2, 3)
    lim = np.sqrt(6.0 / 5.0)
    assert np.abs(v.numpy()).max() <= lim
    s = tf.get_variable("step", [], initializer=tf.zeros_initializer, dtype=tf.int32, trainable=False)
    tf.assign(s, s + 1)
    assert int(s.numpy()) == 1 and v in [v] and s not in [v]
    assert int(tf.reduce_mean(torch.tensor([1, 2, 2], dtype=torch.int32))) == 1  # integer Mean truncates
    assert float(tf.nn.l2_loss(torch.tensor([3.0, 4.0]))) == 12.5
    tf._reset()
