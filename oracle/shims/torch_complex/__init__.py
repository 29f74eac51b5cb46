from . import tensor, functional  # noqa
