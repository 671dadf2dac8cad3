"""Stub: mmh3 is absent; only graphs/createAttributeSum.py (offline summariser,
out of scope) calls it."""
def hash128(*a, **k):
    raise NotImplementedError('mmh3 stub: attribute summariser is out of scope')
