"""Drop-in for the reference's utils/coma_occupancy.py (hot-path symbols only)."""
from coma_b200.coma_occupancy import ComA_Occupancy, load_voxelgrid  # noqa: F401
