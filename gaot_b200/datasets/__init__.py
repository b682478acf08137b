"""Device-side counterparts of the two reference dataset utilities that sit directly on the hot path's input side
(SURVEY.md section 8f): the vx-mode graph builder and the time-pair generator of sequential training."""
from .graph_builder import CachedGraphBuilder, GraphBuilder, rescale  # noqa: F401
from .data_utils import DynamicPairDataset  # noqa: F401
