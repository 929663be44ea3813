"""Weight regularizers — host mirror of dnpy/regularizers.py (same class names, constructor arguments and attributes).

Every penalty here is the elastic-net family  l1*|p| + l2*p^2 ; a regularizer object is just its coefficient pair.
On the hot path only the gradient matters and it is fused into the dW/db epilogues of the Dense/Conv kernels:
``coeffs()`` hands (lambda_l1, lambda_l2) to the C ABI as a ``dnb_reg_t``.  ``forward``/``backward`` evaluate the same
pair on the host for user code that calls them directly (regularizers.py:25-26, 38-39, 52-53).
"""
import numpy as np


class Regularizer:
    _label = "Base regularizer"

    def __init__(self, name=None):
        self.name = self._label if name is None else name

    # (lambda_l1, lambda_l2) of this penalty; the base class is the null penalty
    def coeffs(self):
        return (0.0, 0.0)

    def _is_null(self):
        return type(self) is Regularizer

    def forward(self, param):
        if self._is_null():
            return None
        l1, l2 = self.coeffs()
        p = np.asarray(param)
        terms = []
        if isinstance(self, (L1, L1L2)):
            terms.append(l1 * np.abs(p))
        if isinstance(self, (L2, L1L2)):
            terms.append(l2 * (p ** 2))
        return sum(terms[1:], terms[0])

    def backward(self, param):
        if self._is_null():
            return None
        l1, l2 = self.coeffs()
        p = np.asarray(param)
        terms = []
        if isinstance(self, (L1, L1L2)):
            terms.append(l1 * np.sign(p))          # d|p|/dp
        if isinstance(self, (L2, L1L2)):
            terms.append(l2 * (2 * p))             # d(p^2)/dp
        return sum(terms[1:], terms[0])


class L1(Regularizer):
    _label = "L1"

    def __init__(self, lmda=0.01):
        super().__init__()
        self.lmda_l1 = lmda

    def coeffs(self):
        return (float(self.lmda_l1), 0.0)


class L2(Regularizer):
    _label = "L2"

    def __init__(self, lmda=0.01):
        super().__init__()
        self.lmda_l2 = lmda

    def coeffs(self):
        return (0.0, float(self.lmda_l2))


class L1L2(Regularizer):
    _label = "L1L2"

    def __init__(self, lmda_l1=0.01, lmda_l2=0.01):
        super().__init__()
        # the reference keeps two member regularizers; their coefficients stay the single source of truth
        self.l1_reg, self.l2_reg = L1(lmda_l1), L2(lmda_l2)

    def coeffs(self):
        return (self.l1_reg.coeffs()[0], self.l2_reg.coeffs()[1])


def reg_coeffs(reg):
    """dnb_reg_t pair for an optional regularizer object."""
    return (0.0, 0.0) if reg is None else reg.coeffs()
