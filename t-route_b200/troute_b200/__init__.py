"""troute_b200 -- B200-native compute path behind T-Route's routing entry points.

    troute_b200.routing.compute.compute_nhd_routing_v02          <-> troute.routing.compute
    troute_b200.routing.fast_reach.mc_reach.compute_network_structured
    troute_b200.routing.fast_reach.reach.compute_reach_kernel / compute_reach
    troute_b200.network.RoutingNetwork                            flat-array fast path (ctypes C ABI)
"""
from ._lib import LIB_PATH, TrouteB200Error  # noqa: F401

__all__ = ["LIB_PATH", "TrouteB200Error"]
