"""Process-wide execution settings of the backend (one process drives one GPU).

``math``          "fp32" — SIMT fp32 FMA kernels, parity rel 1e-5 vs the float64 reference;
                  "tf32" — tcgen05 tensor-core kernels for Dense / Conv2D / Pointwise (fp32
                  storage, TF32 operands, fp32 accumulation in TMEM), parity rel 2e-2;
                  "bf16" — as "tf32", with BF16 operands (fp32 storage and accumulation) in the
                  shifted-view convolution kernels: half the shared-memory operand traffic those
                  kernels are bound by; parity rel 2e-2.
                  Chosen once, before ``Net.build`` (env DNPY_B200_MATH or ``set_math``).
``maxpool_route`` "literal" (default) reproduces dnpy/layers/pooling.py:95 exactly;
                  "per_sample" is the intended scatter-add.
``world``/``rank`` data-parallel layout (set by dnpy_b200.parallel.init); gradients that the
                  reference averages over the batch are scaled by 1/world so that the all-reduced
                  sum equals the single-process value.
"""
import os

from . import _lib


_MATH_CODES = {"fp32": _lib.MATH_FP32, "tf32": _lib.MATH_TF32, "bf16": _lib.MATH_BF16}


class _Config:
    def __init__(self):
        self.math = _MATH_CODES.get(os.environ.get("DNPY_B200_MATH", "fp32").lower(), _lib.MATH_FP32)
        self.maxpool_route = (_lib.ROUTE_PER_SAMPLE if os.environ.get("DNPY_B200_MAXPOOL", "literal").lower() == "per_sample"
                              else _lib.ROUTE_LITERAL)
        self.world = 1
        self.rank = 0

    @property
    def route(self):
        """The MaxPool2D backward routing in effect: the literal routing (every delta into sample 0 of the minibatch,
        pooling.py:95) does not shard, so data-parallel runs always use the per-sample routing."""
        return self.maxpool_route if self.world == 1 else _lib.ROUTE_PER_SAMPLE

    @property
    def inv_world(self):
        return 1.0 / float(self.world)

    @property
    def reg_scale(self):
        # regulariser terms are batch independent: contribute 1/world per rank so the sum is exact
        return 1.0 / float(self.world)


config = _Config()


def set_math(mode):
    mode = str(mode).lower()
    if mode not in _MATH_CODES:
        raise ValueError("math must be 'fp32', 'tf32' or 'bf16'")
    config.math = _MATH_CODES[mode]


def get_math():
    return {v: k for k, v in _MATH_CODES.items()}[config.math]


def set_maxpool_route(route):
    route = str(route).lower()
    if route not in ("literal", "per_sample"):
        raise ValueError("maxpool route must be 'literal' or 'per_sample'")
    config.maxpool_route = _lib.ROUTE_PER_SAMPLE if route == "per_sample" else _lib.ROUTE_LITERAL


def get_maxpool_route():
    return "per_sample" if config.maxpool_route == _lib.ROUTE_PER_SAMPLE else "literal"
