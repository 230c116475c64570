"""Mirror of the part of `wfd/wfc/__init__.py:4` that meets the wave: the prior-distribution WFC state and its
information propagation, on the device."""
from .wfc_prior_dist import WFCGeneratorPriorDist, connections_from_wfc_mats, pack_bits, unpack_bits

__all__ = ["WFCGeneratorPriorDist", "connections_from_wfc_mats", "pack_bits", "unpack_bits"]
