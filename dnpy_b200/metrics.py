"""Metrics — host mirror of dnpy/metrics.py (same class names, constructor arguments and attributes).

Two families: metrics that ARE a loss value (MSE, RMSE, MAE: one factory over the loss class, evaluated by that loss's
fused kernel) and the accuracies (one tiny count kernel + a 4-byte read).  They are evaluated on print epochs only.
"""
from . import _lib, losses
from .darray import DArray, dev_empty, stream_ptr


class Metric:
    def __init__(self, name="Base loss"):
        self.name = name

    def compute_metric(self, y_pred, y_target):
        return None


def _loss_metric(loss_name):
    """metrics.py:15-41: a metric whose value is ``losses.<loss_name>().compute_loss``; the loss object is kept in ``cls``."""
    loss_cls = getattr(losses, loss_name)

    class _LossMetric(Metric):
        def __init__(self, name=loss_name):
            Metric.__init__(self, name=name)
            self.cls = loss_cls()

        def compute_metric(self, y_pred, y_target):
            return self.cls.compute_loss(y_pred, y_target)

    _LossMetric.__name__ = _LossMetric.__qualname__ = loss_name
    return _LossMetric


MSE = _loss_metric("MSE")
RMSE = _loss_metric("RMSE")
MAE = _loss_metric("MAE")


class _CountMetric(Metric):
    """Fraction of rows on which prediction and target agree; the agreement test is a device kernel."""
    _kernel = None

    def _extra(self):
        return ()

    def _agreeing_rows(self, y_pred, y_target):
        p, t = DArray.wrap(y_pred), DArray.wrap(y_target)
        rows, cols = p.shape
        count = dev_empty((1,))
        _lib.call(self._kernel, p.ptr(), t.ptr(), count.data_ptr(), rows, cols, *self._extra(), stream_ptr())
        return float(count.item()), rows, cols


class BinaryAccuracy(_CountMetric):
    """metrics.py:44-60: classes by thresholding both tensors; float(mean(axis=0)) => one output column only."""
    _kernel = "dnb_metric_binary"

    def __init__(self, threshold=0.5, name="BinaryAccuracy"):
        super().__init__(name=name)
        self.threshold = threshold

    def _extra(self):
        return (float(self.threshold),)

    def compute_metric(self, y_pred, y_target):
        hits, rows, cols = self._agreeing_rows(y_pred, y_target)
        if cols != 1:       # the reference's float() of a per-column mean
            raise TypeError("only length-1 arrays can be converted to Python scalars")
        return hits / float(rows)


class CategoricalAccuracy(_CountMetric):
    """metrics.py:63-80: first-argmax of prediction vs target per row."""
    _kernel = "dnb_metric_categorical"

    def __init__(self, name="CategoricalAccuracy"):
        super().__init__(name=name)

    def compute_metric(self, y_pred, y_target):
        hits, rows, _ = self._agreeing_rows(y_pred, y_target)
        return hits / float(rows)
