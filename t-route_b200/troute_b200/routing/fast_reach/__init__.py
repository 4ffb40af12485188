"""troute_b200.routing.fast_reach -- GPU twins of troute.routing.fast_reach.{mc_reach, reach, simple_da}."""
