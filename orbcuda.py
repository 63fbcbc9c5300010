"""Import alias: `import orbcuda` -> the package in cooperative-orb-slam_b200/ (the directory name the
project layout prescribes is not a valid Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("cooperative-orb-slam_b200")
sys.modules[__name__] = _pkg
