"""Drop-in for the reference's utils/misc.py (hot-path symbols only)."""
from coma_b200.misc import get_3d_indexgrid_ijk, to_np_torch_recursive  # noqa: F401
