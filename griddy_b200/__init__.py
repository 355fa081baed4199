"""griddy_b200 — B200-native rebuild of griddy's per-tick actor update loop.

Only the hot path of the reference (griddy/simulate.scm:141-153 and griddy/simulate/route.scm) is
rebuilt here, as hand-written sm_100a CUDA behind the C ABI of include/griddy_cuda.h.  The Python
modules mirror the reference's Scheme API for that path (core.py, simulate.py) and flatten a world
into the structure-of-arrays buffers the ABI takes (flat.py, flatten.py).
"""
from .flat import (ActorState, FlatActors, FlatNetwork, BACK, FORW, ITEM_SLEEP, ITEM_TRAVEL,
                   K_JLANE, K_JUNCTION, K_OTHER, K_SEGLANE, K_SEGMENT, ROUTE_DONE, ROUTE_NONE,
                   ROUTE_SOME)

__all__ = ["ActorState", "FlatActors", "FlatNetwork", "FORW", "BACK", "ITEM_SLEEP", "ITEM_TRAVEL",
           "K_JUNCTION", "K_SEGMENT", "K_SEGLANE", "K_JLANE", "K_OTHER", "ROUTE_NONE", "ROUTE_SOME",
           "ROUTE_DONE"]
