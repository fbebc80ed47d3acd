"""spn4cir_b200 -- B200-native (sm_100a) hot path of SPN4CIR: memory-bank contrastive loss
(forward + backward, incl. BLIP-2 max-over-tokens) and query x gallery similarity with top-K / rank.

Layout: ``csrc/`` hand-written CUDA + the C ABI (``include/spn_b200.h``); ``_lib`` ctypes loader;
``ops`` / ``retrieval`` torch-facing operators; ``compat`` mirrors of the reference call sites;
``dist`` bank/gallery row sharding across ranks.  There is no CPU fallback anywhere in this package.
"""
from ._lib import SpnError, build_library, load_library  # noqa: F401
from .ops import DeviceBank, load_bank_file, bank_ce_stats, bank_ce_step, bank_infonce, part_infonce, cast_to_bf16, merge_stats, query_assemble  # noqa: F401
from .retrieval import merge_topk, recall_at, release_workspaces, sim_rank, sim_target_score, sim_topk  # noqa: F401

from . import compat, dist, synth  # noqa: E402,F401

__version__ = "0.1.0"
