class NeighborOverflowError(Exception):
    """Raised when a row needs more than max_neighbors slots (call sites: aimnet/calculators/neighbors.py:127)."""
