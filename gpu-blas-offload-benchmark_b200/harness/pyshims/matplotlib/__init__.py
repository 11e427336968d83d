"""Minimal stand-in for the slice of matplotlib that the reference's createGflopsGraphs.py
uses (reference createGflopsGraphs.py:3,144-210,312-378), drawn with Pillow.

matplotlib is not installed in the build image and there is no network; this directory is
appended to PYTHONPATH *behind* site-packages by scripts/run_reference_scripts.py, so a real
matplotlib always wins when it exists.  Not part of the product: post-processing only.
"""
__version__ = "0.0-b200-standin"
