"""CPU: the TensorFlow-1 graph interpreter the reference-graph goldens are generated on (oracle/tf1_shim.py, test
infrastructure) against hand-computed answers of the TF rules it restates: optimiser steps and their slot ownership,
fetch order of Session.run, l2_normalize with and without an axis, gradient conventions at kinks, the sparse ops."""
import math

import numpy as np
import pytest

from oracle import tf1_shim as tf


@pytest.fixture(autouse=True)
def fresh_graph():
    tf.reset_default_graph()
    yield
    tf.reset_default_graph()


def _quadratic(opt_cls, **kw):
    v = tf.Variable(np.array([[1.0, -2.0]]), name="v")
    loss = tf.reduce_sum(tf.square(v))                  # gradient 2v
    step = opt_cls(0.1, **kw).minimize(loss)
    return v, loss, step


def test_optimiser_rules_one_and_two_steps():
    sess = tf.Session()
    v, loss, step = _quadratic(tf.train.GradientDescentOptimizer)
    assert sess.run([loss, step])[0] == pytest.approx(5.0)              # the fetched loss is the pre-update value
    np.testing.assert_allclose(v.value.detach().numpy(), [[0.8, -1.6]])

    tf.reset_default_graph()
    v, loss, step = _quadratic(tf.train.AdagradOptimizer)
    sess.run(step)
    g = np.array([2.0, -4.0])
    acc = 0.1 + g * g                                                    # initial_accumulator_value = 0.1
    want = np.array([1.0, -2.0]) - 0.1 * g / np.sqrt(acc)
    np.testing.assert_allclose(v.value.detach().numpy()[0], want, rtol=1e-12)
    sess.run(step)
    g2 = 2 * want
    want2 = want - 0.1 * g2 / np.sqrt(acc + g2 * g2)
    np.testing.assert_allclose(v.value.detach().numpy()[0], want2, rtol=1e-12)

    tf.reset_default_graph()
    v, loss, step = _quadratic(tf.train.AdamOptimizer)
    sess.run(step)
    m, s = 0.1 * g, 0.001 * g * g
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(v.value.detach().numpy()[0], np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(s) + 1e-8), rtol=1e-12)

    tf.reset_default_graph()
    v, loss, step = _quadratic(tf.train.AdadeltaOptimizer)
    sess.run(step)
    acc = 0.05 * g * g
    up = np.sqrt(1e-8) / np.sqrt(acc + 1e-8) * g
    np.testing.assert_allclose(v.value.detach().numpy()[0], np.array([1.0, -2.0]) - 0.1 * up, rtol=1e-12)


def test_every_optimiser_instance_owns_its_slots_and_untouched_variables_stay():
    a, b = tf.Variable(np.ones((2, 2)), name="a"), tf.Variable(np.ones((2, 2)), name="b")
    loss1, loss2 = tf.reduce_sum(tf.square(a)), tf.reduce_sum(tf.square(a)) * 3.0
    s1, s2 = tf.train.AdagradOptimizer(0.1).minimize(loss1), tf.train.AdagradOptimizer(0.1).minimize(loss2)
    sess = tf.Session()
    sess.run(s1)
    sess.run(s2)
    assert len(a.slots) == 2 and len(b.slots) == 0                       # b is not reachable from either loss
    np.testing.assert_allclose(b.value.detach().numpy(), 1.0)
    acc1, acc2 = [slots["accumulator"] for slots in a.slots.values()]
    assert not np.allclose(acc1.numpy(), acc2.numpy())                   # different gradients accumulated separately


def test_l2_normalize_axes_and_kink_gradients():
    x = tf.placeholder(tf.float32)
    sess = tf.Session()
    arr = np.array([[3.0, 4.0], [0.0, 0.0]])
    np.testing.assert_allclose(sess.run(tf.nn.l2_normalize(x, 1), {x: arr}), [[0.6, 0.8], [0.0, 0.0]])
    np.testing.assert_allclose(sess.run(tf.nn.l2_normalize(x), {x: arr}), arr / 5.0)          # all elements
    v = tf.Variable(np.array([0.0, -1.0, 2.0]), name="k")
    step = tf.train.GradientDescentOptimizer(1.0).minimize(tf.reduce_sum(tf.nn.relu(v)) + tf.reduce_sum(tf.abs(v)))
    sess.run(step)
    np.testing.assert_allclose(v.value.detach().numpy(), [0.0, 0.0, 0.0])    # relu'(0) = 0, |x|' at 0 = 0; −1 → +1·(−1)… = 0; 2 − (1 + 1) = 0
    w = tf.Variable(np.array([1.0]), name="w")
    step = tf.train.GradientDescentOptimizer(1.0).minimize(tf.reduce_sum(tf.maximum(w, 1.0)))
    sess.run(step)
    np.testing.assert_allclose(w.value.detach().numpy(), [0.0])              # a tie sends the gradient to the first argument


def test_sparse_ops_and_named_placeholders():
    sess = tf.Session()
    idx = np.array([[0, 1], [0, 0], [1, 1], [0, 1]])                         # a duplicate coordinate, unordered
    sp = tf.SparseTensor(idx, tf.constant([1.0, 2.0, 3.0, 1.0]), [2, 2])
    soft = sess.run(tf.sparse_softmax(sp))
    e = np.exp(np.array([1.0, 2.0, 1.0]) - 2.0)
    np.testing.assert_allclose(soft.values.numpy(), [e[0] / e.sum(), e[1] / e.sum(), 1.0, e[2] / e.sum()], rtol=1e-12)
    dense = np.array([[1.0, 2.0], [10.0, 20.0]])
    np.testing.assert_allclose(sess.run(tf.sparse_tensor_dense_matmul(sp, tf.constant(dense))), [[2 + 10 + 10, 4 + 20 + 20], [30, 60]])
    scaled = sess.run(tf.cast(sp, tf.float32) * tf.constant(np.array([[2.0], [5.0]])))         # [n, 1] broadcasts over rows
    np.testing.assert_allclose(scaled.values.numpy(), [2.0, 4.0, 15.0, 2.0])
    ph = tf.placeholder(tf.int32, name="neg_left")
    assert tf.get_default_graph().get_tensor_by_name("neg_left:0") is ph
    assert sess.run(ph + 1, {"neg_left:0": np.array([1, 2])}).tolist() == [2, 3]
    with pytest.raises(KeyError):
        tf.get_default_graph().get_tensor_by_name("absent:0")


def test_while_loop_unrolls_a_constant_countdown():
    x = tf.placeholder(tf.float32)
    total = tf.while_loop(lambda i, s: tf.greater(i, 0), lambda i, s: (tf.subtract(i, 1), tf.add(s, x * tf.cast(i, tf.float32))),
                          [tf.constant(3), tf.zeros([2])])[1]
    np.testing.assert_allclose(tf.Session().run(total, {x: np.array([1.0, 2.0])}), [6.0, 12.0])
