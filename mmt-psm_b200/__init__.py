"""mmtpsm_b200 — B200-native (sm_100a) implementation of the MMT-PSM mean-teacher training hot path.

Host side mirrors the reference's `maskrcnn_benchmark` operator surface (cfg / registry / BoxList /
`_C`); the compute is hand-written CUDA behind the C ABI of include/mmtpsm_b200.h.  The directory is
called `mmt-psm_b200/`; import it as `mmt_psm_b200` (see the loader `mmt_psm_b200.py` at the repo root).
"""
__version__ = "0.1"
