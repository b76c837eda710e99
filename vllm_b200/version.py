"""Version string (reference: vllm/version.py:1-13)."""
__version__ = "0.1.0.dev0"
__version_tuple__ = (0, 1, 0, "dev0")
