"""troute_b200.routing -- mirrors the troute.routing package layout for the hot path
(/root/reference/src/troute-routing/troute/routing)."""
