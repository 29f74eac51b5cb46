"""Stub for humanfriendly (only used for log formatting in the reference host code)."""
def parse_size(s):
    return int(float(s))
def format_size(n, *a, **k):
    return str(n)
def format_timespan(n, *a, **k):
    return str(n)
