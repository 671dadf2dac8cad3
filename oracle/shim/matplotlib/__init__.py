"""Stub: matplotlib is absent in this image; reference helpers/results.py:2 and
helpers/vizEmb.py:2 import pyplot only for plotting (out of scope)."""
