"""Minimal stand-in for `scityping` (typing aliases only) -- see ../README.md."""
import numbers
import numpy as _np


class Serializable:
    """No-op base: the reference only uses it for (de)serialisation, which is out of scope."""

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)


Number = numbers.Number
Real = numbers.Real


class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls


class TorchTensor(_Subscriptable):
    pass


class DType(_Subscriptable):
    pass
