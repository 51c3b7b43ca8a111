"""Biopython stand-in (TEST INFRASTRUCTURE): only what the reference scripts import."""
__version__ = "0.0-shim"
