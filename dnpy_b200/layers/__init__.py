"""Layer classes with the reference's names and constructors (dnpy/layers/__init__.py:1-8)."""
from .base import Layer, MLayer
from .activations import LeakyRelu, Relu, PRelu, Sigmoid, Tanh, Softmax, LogSoftmax
from .convolution import Conv, Conv2D, PointwiseConv2D, DepthwiseConv2D
from .core import Input, Embedding, Dense, Reshape, Flatten
from .merge import MPWLayer, Add, Subtract, Multiply, Divide, Power, Maximum, Minimum, Average, Concatenate
from .pooling import Pool, MaxPool2D, GlobalMaxPool2D, AvgPool2D, GlobalAvgPool2D
from .recurrent import RNNCell, BaseRNN, SimpleRNN
from .regularization import Dropout, GaussianNoise, BatchNorm
from .. import initializers, utils  # the reference re-exports these through `from dnpy.layers import *`
