"""B200-native Poisson surface reconstruction (sm_100a).  The compute path lives in
``libprb.so`` (hand-written CUDA behind the C ABI of ``include/prb.h``); this package is the
thin host-side mirror used by tests, bench.py and the multi-GPU launcher.  There is no CPU
fallback: importing works anywhere, but every compute call raises if the library or a B200 is
missing."""
from .api import PoissonRecon, PrbError, assemble_mesh, deal_passes, lib_path, load_library, shard_plan  # noqa: F401
