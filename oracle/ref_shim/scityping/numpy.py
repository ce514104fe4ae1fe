class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls


class Array(_Subscriptable):
    pass


class DType(_Subscriptable):
    pass
