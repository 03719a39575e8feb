"""Import shims that let the reference's unmodified Python files run on this package."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
