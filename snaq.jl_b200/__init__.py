"""snaq.jl_b200 — B200-native quartet-pseudolikelihood engine for SNaQ.

Host-side mirror of the reference's hot-path interface (same names, argument
meaning and error behaviour) over the C ABI of ``include/snaq_b200.h``.
"""
from .network import (HybridNetwork, Node, Edge, SnaqError, readnewicklevel1,  # noqa: F401
                      network_from_lists, hybridEdges, getOtherNode)
from .datacf import DataCF, fittedquartetCF, readtableCF, writeExpCF  # noqa: F401
from .engine import (Engine, load_library, LIB_PATH, extractQuartet_, calculateExpCFAll_, logPseudoLik,  # noqa: F401
                     topologyQpseudolik_, parameters_, optBL_, optBL_candidates_, updateBL_)
from . import simulate, sharded  # noqa: F401
