"""Permissive typeguard shim (typeguard 4.x rejects ESPnet's `str = None` defaults)."""
_A3T_SHIM = True
def check_argument_types(*a, **k):
    return True
def check_return_type(*a, **k):
    return True
def check_type(value, *a, **k):
    return value
def typechecked(f=None, **k):
    if f is None:
        return lambda g: g
    return f
def __getattr__(name):
    return lambda *a, **k: True
