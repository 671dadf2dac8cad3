"""Stub: distutils was removed in Python 3.12; reference main.py:4 imports strtobool."""
