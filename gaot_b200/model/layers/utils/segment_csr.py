"""Drop-in for reference src/model/layers/utils/segment_csr.py:14-19."""
from ....ops import segment_csr  # noqa: F401
