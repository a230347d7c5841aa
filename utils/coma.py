"""Drop-in for the reference's utils/coma.py (hot-path symbols only), backed by coma_b200's sm_100a kernels.
`negative_exp` must stay importable from here: every exported ComA pickle references `utils.coma.negative_exp`."""
from coma_b200.coma import (ComA, get_aggregated_contact, get_nonphysical_score, nearest_vertex_indices,  # noqa: F401
                            negative_exp, simplify_mesh_and_get_indices)
from coma_b200.misc import get_uniform_points_on_sphere  # noqa: F401
