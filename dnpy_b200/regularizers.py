"""Weight regularizers (dnpy/regularizers.py).

On the hot path only their gradient matters, and it is fused into the dW/db
epilogues of the Dense/Conv kernels: ``coeffs()`` hands the (lambda_l1,
lambda_l2) pair to the C ABI (dnb_reg_t).  ``forward``/``backward`` keep the
reference's host semantics for user code that calls them directly.
"""
import numpy as np


class Regularizer:
    def __init__(self, name="Base regularizer"):
        self.name = name

    def coeffs(self):
        return (0.0, 0.0)

    def forward(self, param):
        pass

    def backward(self, param):
        pass


class L1(Regularizer):
    """regularizers.py:16-26."""
    def __init__(self, lmda=0.01):
        super().__init__(name="L1")
        self.lmda_l1 = lmda

    def coeffs(self):
        return (float(self.lmda_l1), 0.0)

    def forward(self, param):
        return self.lmda_l1 * np.abs(np.asarray(param))

    def backward(self, param):
        return self.lmda_l1 * np.sign(np.asarray(param))


class L2(Regularizer):
    """regularizers.py:29-39."""
    def __init__(self, lmda=0.01):
        super().__init__(name="L2")
        self.lmda_l2 = lmda

    def coeffs(self):
        return (0.0, float(self.lmda_l2))

    def forward(self, param):
        return self.lmda_l2 * (np.asarray(param) ** 2)

    def backward(self, param):
        return self.lmda_l2 * (2 * np.asarray(param))


class L1L2(Regularizer):
    """regularizers.py:42-53."""
    def __init__(self, lmda_l1=0.01, lmda_l2=0.01):
        super().__init__(name="L1L2")
        self.l1_reg = L1(lmda_l1)
        self.l2_reg = L2(lmda_l2)

    def coeffs(self):
        return (float(self.l1_reg.lmda_l1), float(self.l2_reg.lmda_l2))

    def forward(self, param):
        return self.l1_reg.forward(param) + self.l2_reg.forward(param)

    def backward(self, param):
        return self.l1_reg.backward(param) + self.l2_reg.backward(param)


def reg_coeffs(reg):
    """dnb_reg_t pair for an optional regularizer object."""
    return reg.coeffs() if reg is not None else (0.0, 0.0)
