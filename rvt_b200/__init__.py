"""rvt_b200 — B200-native (sm_100a) implementation of RVT's hot path: the per-timestep recurrent
backbone forward (``RNNDetector``) and the event -> ``StackedHistogram`` voxelizer, behind the
reference's own module API.  See DESIGN.md / INTEGRATION.md."""
from .backbone import RNNDetector, RNNDetectorStage, build_recurrent_backbone  # noqa: F401
from .representations import MixedDensityEventStack, StackedHistogram  # noqa: F401
from . import preprocessing  # noqa: F401
from .graph import GraphedCallable, capture_sequence  # noqa: F401

MaxViTRNNDetector = RNNDetector
