"""wmf_py.models_py-compatible model backed by the B200 library."""
from .core import ModelParams, ModelState, shia_v1  # noqa: F401
