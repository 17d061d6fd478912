"""wmf_py.cu_py-compatible operators backed by the B200 library."""
from . import basics, metrics, streams  # noqa: F401
