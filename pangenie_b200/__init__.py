"""pangenie_b200 — B200-native implementation of PanGenie's genotyping hot path
(k-mer counting -> unique-k-mer emissions -> forward-backward HMM) behind a C-ABI
(include/pangenie_b200.h).  The compute lives in csrc/*.cu (sm_100a); this package is the thin
host-side mirror of the reference's interface plus multi-GPU plumbing over torch.distributed.
"""
from .capi import PG_OP_COUNT, PG_OP_PRIME, PG_OP_UPDATE, load  # noqa: F401
from .model import HMM, Engine, Index, KmerCounter, PgError, ProbabilityTable, UniqueKmerSelection, copy_number, haplotype_sample  # noqa: F401
from .panel import Panel, PanelBuilder, Result  # noqa: F401

__all__ = ["HMM", "Engine", "Index", "KmerCounter", "PgError", "ProbabilityTable", "copy_number", "Panel", "PanelBuilder",
           "Result", "UniqueKmerSelection", "haplotype_sample", "PG_OP_COUNT", "PG_OP_PRIME", "PG_OP_UPDATE", "load"]
