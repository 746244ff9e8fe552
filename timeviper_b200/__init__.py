"""timeviper_b200 -- B200-native (sm_100a) Mamba-2 mixer prefill path of xiaomi-research/timeviper.

Public surface = the reference's own operator names for this path (see ops.py) plus the mixer drop-in.
Importing this package requires the in-tree CUDA extension (libtimeviper_b200.so); there is no fallback.
"""
from . import _lib

_lib.load()

from .config import Mamba2Config  # noqa: E402
from .mixer import Mamba2MixerPrefill, MambaRMSNormGated, patch_reference  # noqa: E402
from .ops import (causal_conv1d_fn, causal_conv1d_update, fold_boundary_states, mamba_chunk_scan_combined,  # noqa: E402
                  mamba_chunk_state_summary, mamba_split_conv1d_scan_combined, rmsnorm_fn,
                  selective_state_update, ssd_kernel_family, launch_count)
from .sharded import (sharded_mixer_forward, sharded_prefill_from_host, sharded_scan_core,  # noqa: E402
                      sharded_scan_core_graph)
from .hybrid import HybridCausalLM, HybridPrefillStack  # noqa: E402

__all__ = ["Mamba2Config", "Mamba2MixerPrefill", "MambaRMSNormGated", "patch_reference", "causal_conv1d_fn",
           "causal_conv1d_update", "mamba_chunk_scan_combined", "mamba_split_conv1d_scan_combined",
           "rmsnorm_fn", "selective_state_update", "mamba_chunk_state_summary", "fold_boundary_states",
           "ssd_kernel_family", "sharded_mixer_forward", "sharded_scan_core", "sharded_prefill_from_host", "sharded_scan_core_graph",
           "HybridPrefillStack", "HybridCausalLM", "launch_count"]
