"""Numerical gradient check through the drop-in's public API — TEST INFRASTRUCTURE.

Follows the protocol of the reference's harness (tests/coverage/test_gradient_checking.py:15-127 with
dnpy/utils.py:112-148): one analytic pass, then for a random sample of the trainable weights two perturbed passes
``f(theta +- eps)`` driven ONLY through the boundary a dnpy user has — ``feed_input / forward / compute_losses /
reset_grads / do_delta / backward``, ``get_params(only_trainable=True) / get_grads / set_params(only_trainable=True)``
and ``utils.params2vector / vector2params / sample_params_slices`` — and the reference's norm-wise error
``|a - n| / (|a| + |n|)``.  The reference probes with eps = 1e-6 in float64; the drop-in computes in fp32 (loss
resolution ~1e-7), so the probe is eps = 1e-2 and the bound 5e-3 (central differences: O(eps^2) truncation plus
~1e-7/eps round-off), stated where the test calls it.
"""
import numpy as np


def analytic_and_numeric(model, utils, x_mb, y_mb, max_samples, eps):
    def passes():
        model.feed_input(x_mb)
        model.forward()
        losses, deltas = model.compute_losses(y_target=y_mb)
        model.reset_grads()
        model.do_delta(deltas)
        model.backward()
        return losses

    passes()
    theta0 = model.get_params(only_trainable=True)
    g0 = model.get_grads()
    vec, slices = utils.params2vector(theta0)
    gvec, _ = utils.params2vector(g0)
    picked = np.array(utils.sample_params_slices(slices, max_samples=max_samples))
    analytic = gvec[picked].astype(np.float64).ravel()
    numeric = np.zeros_like(analytic)
    for i, j in enumerate(picked):
        side = []
        for sign in (+1.0, -1.0):
            v = np.array(vec, dtype=np.float64)
            v[j] += sign * eps
            model.set_params(utils.vector2params(v, theta0), only_trainable=True)
            side.append(float(passes()[0]))
        numeric[i] = (side[0] - side[1]) / (2.0 * eps)
    model.set_params(theta0, only_trainable=True)
    return analytic, numeric


def gradient_check(model, utils, x_train, y_train, batch_size, max_error, max_samples=5, burn_in_passes=0, eps=1e-2):
    """True when every minibatch's norm-wise error stays under ``max_error`` (the reference's pass criterion)."""
    n_batches = int(len(x_train[0]) / batch_size)
    if model.mode != "train":
        model.set_mode("train")
    for _ in range(burn_in_passes):
        for b in range(n_batches):
            x_mb, y_mb = model.get_minibatch(x_train, y_train, b, batch_size)
            model.feed_input(x_mb)
            model.forward()
            _, deltas = model.compute_losses(y_target=y_mb)
            model.reset_grads()
            model.do_delta(deltas)
            model.backward()
            model.apply_grads()
    worst = 0.0
    for b in range(n_batches):
        x_mb, y_mb = model.get_minibatch(x_train, y_train, b, batch_size)
        a, n = analytic_and_numeric(model, utils, x_mb, y_mb, max_samples, eps)
        err = float(np.linalg.norm(a - n) / (np.linalg.norm(a) + np.linalg.norm(n)))
        worst = max(worst, err)
        if err > max_error:
            return False, worst
    return True, worst
