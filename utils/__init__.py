"""Import-path shim: the reference's scripts (and its pickles) address the ComA code as `utils.coma`,
`utils.coma_occupancy`, `utils.misc`. These modules re-export the B200 implementations from `coma_b200` under the
reference's names, so `from utils.coma import ComA, get_aggregated_contact` keeps working unchanged."""
