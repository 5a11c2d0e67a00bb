"""subzero_b200 -- B200-native contact-force step of the SubZero sea-ice model."""
from .abi import FloesSoA, Boundary, SzParams, SzSummary, SzError, default_params, LIB_PATH  # noqa: F401
from .contact import ContactContext, floe_interactions_all, floes_to_soa  # noqa: F401
from .field import voronoi_field  # noqa: F401
