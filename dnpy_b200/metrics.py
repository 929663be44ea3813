"""Metrics (dnpy/metrics.py): evaluated only on print epochs; the accuracies are one tiny
count kernel + a 4-byte read."""
from . import _lib, losses
from .darray import DArray, dev_empty, stream_ptr


class Metric:
    def __init__(self, name="Base loss"):
        self.name = name

    def compute_metric(self, y_pred, y_target):
        pass


class MSE(Metric):
    """metrics.py:15-22."""
    def __init__(self, name="MSE"):
        super().__init__(name=name)
        self.cls = losses.MSE()

    def compute_metric(self, y_pred, y_target):
        return self.cls.compute_loss(y_pred, y_target)


class RMSE(Metric):
    def __init__(self, name="RMSE"):
        super().__init__(name=name)
        self.cls = losses.RMSE()

    def compute_metric(self, y_pred, y_target):
        return self.cls.compute_loss(y_pred, y_target)


class MAE(Metric):
    def __init__(self, name="MAE"):
        super().__init__(name=name)
        self.cls = losses.MAE()

    def compute_metric(self, y_pred, y_target):
        return self.cls.compute_loss(y_pred, y_target)


class _CountMetric(Metric):
    def _count(self, kernel, y_pred, y_target, *extra):
        p, t = DArray.wrap(y_pred), DArray.wrap(y_target)
        b, n = p.shape
        out = dev_empty((1,))
        _lib.call(kernel, p.ptr(), t.ptr(), out.data_ptr(), b, n, *extra, stream_ptr())
        return float(out.item()), b, n


class BinaryAccuracy(_CountMetric):
    """metrics.py:44-60."""
    def __init__(self, threshold=0.5, name="BinaryAccuracy"):
        super().__init__(name=name)
        self.threshold = threshold

    def compute_metric(self, y_pred, y_target):
        cnt, b, n = self._count("dnb_metric_binary", y_pred, y_target, float(self.threshold))
        if n != 1:
            raise TypeError("only length-1 arrays can be converted to Python scalars")
        return cnt / float(b)


class CategoricalAccuracy(_CountMetric):
    """metrics.py:63-80."""
    def __init__(self, name="CategoricalAccuracy"):
        super().__init__(name=name)

    def compute_metric(self, y_pred, y_target):
        cnt, b, _ = self._count("dnb_metric_categorical", y_pred, y_target)
        return cnt / float(b)
