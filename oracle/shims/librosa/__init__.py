from . import filters  # noqa
