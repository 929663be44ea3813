"""Parameter initializers (dnpy/initializers.py) — host NumPy on purpose.

They run once, at ``Net.build``; keeping them on the host and drawing from the
GLOBAL ``np.random`` stream in the same call order as the reference makes the
initial weights bit-identical to dnpy's for a given ``np.random.seed`` (the
arrays are uploaded afterwards).  Like the reference they REBIND ``params[k]``.
"""
import numpy as np


class Initializer:
    def __init__(self, name="Base initializer", seed=None):
        self.name = name
        self.seed = seed

    def _draw(self, shape):
        raise NotImplementedError

    def apply(self, params, keys=None):
        for k in (keys if keys else list(params.keys())):
            params[k] = self._draw(tuple(params[k].shape))


class Constant(Initializer):
    """initializers.py:15-24."""
    def __init__(self, fill_value=1.0, name="Constant"):
        super().__init__(name=name)
        self.fill_value = fill_value

    def _draw(self, shape):
        return np.full(shape, fill_value=self.fill_value)


class Zeros(Constant):
    def __init__(self, name="Zeros"):
        super().__init__(fill_value=0.0, name=name)


class Ones(Constant):
    def __init__(self, name="Ones"):
        super().__init__(fill_value=1.0, name=name)


class RandomNormal(Initializer):
    """initializers.py:39-49."""
    def __init__(self, mean=0.0, stddev=0.05, name="RandomNormal", seed=None):
        super().__init__(name=name, seed=seed)
        self.mean, self.stddev = mean, stddev

    def _draw(self, shape):
        return self.stddev * np.random.randn(*shape) + self.mean


class RandomUniform(Initializer):
    """initializers.py:52-62."""
    def __init__(self, minval=-0.05, maxval=0.05, name="RandomUniform", seed=None):
        super().__init__(name=name, seed=seed)
        self.minval, self.maxval = minval, maxval

    def _draw(self, shape):
        return (self.maxval - self.minval) * np.random.random(shape) + self.minval


class GlorotNormal(Initializer):
    """initializers.py:65-77."""
    def __init__(self, fan_in=None, fan_out=None, name="GlorotNormal", seed=None):
        super().__init__(name=name, seed=seed)
        self.fan_in, self.fan_out = fan_in, fan_out

    def _draw(self, shape):
        return np.random.randn(*shape) * np.sqrt(2.0 / (self.fan_in + self.fan_out))


class GlorotUniform(Initializer):
    """initializers.py:80-94."""
    def __init__(self, fan_in=None, fan_out=None, minval=-0.05, maxval=0.05, name="GlorotUniform", seed=None):
        super().__init__(name=name, seed=seed)
        self.minval, self.maxval = minval, maxval
        self.fan_in, self.fan_out = fan_in, fan_out

    def _draw(self, shape):
        u = (self.maxval - self.minval) * np.random.random(shape) + self.minval
        return u * np.sqrt(2.0 / (self.fan_in + self.fan_out))


class HeNormal(Initializer):
    """initializers.py:97-109."""
    def __init__(self, fan_in=None, fan_out=None, name="HeNormal", seed=None):
        super().__init__(name=name, seed=seed)
        self.fan_in, self.fan_out = fan_in, fan_out

    def _draw(self, shape):
        return np.random.randn(*shape) * np.sqrt(2.0 / self.fan_in)
