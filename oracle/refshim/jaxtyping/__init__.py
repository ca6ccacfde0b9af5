"""Minimal stand-in for jaxtyping (annotations only); see ../README.md."""


class _Sub:
    def __class_getitem__(cls, item):
        return cls


class Array(_Sub):
    pass


class Float(_Sub):
    pass


class Int(_Sub):
    pass


class Bool(_Sub):
    pass


class PyTree(_Sub):
    pass
