"""Stub: the A3T hot path never calls editdistance (imported by espnet's ErrorCalculator)."""
def eval(a, b):  # noqa: A001
    raise NotImplementedError("editdistance stub")
