"""pmvs_b200 — Python front-end of the B200-native pais-mvs patch-refinement path (tests / benchmarks)."""
from . import abi  # noqa: F401
