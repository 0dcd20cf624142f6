"""B200-native fitness engine for the Evolutionary Illusion Generator (hot path only)."""
from .grid import StructureType, create_grid  # noqa: F401

__all__ = ["StructureType", "create_grid"]
