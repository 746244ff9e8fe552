"""Dimensions of the Mamba-2 mixer, with the reference's attribute names
(timeviper/model/llm/llm_repo/nano/configuration_nano.py:133-258; read by the mixer at
modeling_nano.py:393-411)."""
from dataclasses import dataclass, field
from typing import Tuple


@dataclass
class Mamba2Config:
    hidden_size: int = 4480
    mamba_num_heads: int = 128
    mamba_head_dim: int = 80
    n_groups: int = 8
    ssm_state_size: int = 128
    chunk_size: int = 128
    conv_kernel: int = 4
    layer_norm_epsilon: float = 1e-5
    time_step_limit: Tuple[float, float] = (0.0, float("inf"))
    time_step_min: float = 0.001
    time_step_max: float = 0.1
    time_step_floor: float = 1e-4
    use_conv_bias: bool = True
    use_bias: bool = False
    mamba_hidden_act: str = "silu"
    num_hidden_layers: int = 56
    # the other two block types of the hybrid stack (timeviper_b200/hybrid.py; configuration_nano.py:137-175)
    hybrid_override_pattern: str = "M"          # one character per layer: M = Mamba-2, * = attention, - = MLP
    num_attention_heads: int = 40
    num_key_value_heads: int = 8
    head_dim: int = 128                         # attention head dim
    attention_bias: bool = False
    intermediate_size_mlp: int = 15680          # NemotronHConfig.intermediate_size (ours is d_inner of the mixer)
    mlp_bias: bool = False
    mlp_hidden_act: str = "relu2"
    residual_in_fp32: bool = False
    vocab_size: int = 131072
    # TransV (configuration_nano.py:177-179): cross-attention merge of the dropped vision tokens into the text tokens
    merge_module: str = "no_merge"
    pdrop_type: str = None

    @property
    def intermediate_size(self):
        return self.mamba_num_heads * self.mamba_head_dim

    @property
    def conv_dim(self):
        return self.intermediate_size + 2 * self.n_groups * self.ssm_state_size

    @property
    def projection_size(self):
        return self.intermediate_size + self.conv_dim + self.mamba_num_heads

    @classmethod
    def nanov2_9b(cls):
        """NVIDIA-Nemotron-Nano-9B-v2 Mamba-2 layer (SURVEY.md section 8 header: the checkpoint's config.json
        is not in the reference tree; every field is overridable)."""
        return cls()

    @classmethod
    def nanov2_9b_hybrid(cls, **kw):
        """The 56-layer Nanov2-9B-shaped hybrid stack of BASELINE.json configs[3] (SURVEY.md section 8 header): attention at
        layers 14 / 21 / 30 / 39, Mamba-2 and MLP layers alternating around them: 27 M / 4 * / 25 -."""
        attn, pat, k = {14, 21, 30, 39}, [], 0
        for i in range(56):
            if i in attn:
                pat.append("*")
            else:
                pat.append("M-"[k % 2]); k += 1
        pat[len(pat) - 1 - pat[::-1].index("-")] = "M"
        return cls(num_hidden_layers=56, hybrid_override_pattern="".join(pat), **kw)

    @classmethod
    def small(cls, n_groups=1):
        """BASELINE.json configs[0]: the reference's CPU-runnable small config."""
        return cls(hidden_size=512, mamba_num_heads=16, mamba_head_dim=80, n_groups=n_groups,
                   ssm_state_size=128, chunk_size=128)

    @classmethod
    def from_hf(cls, cfg):
        """From a reference NemotronHConfig (or anything with the same attributes)."""
        names = [f for f in cls.__dataclass_fields__]
        kw = {n: getattr(cfg, n) for n in names if hasattr(cfg, n) and getattr(cfg, n) is not None}
        if hasattr(cfg, "intermediate_size"):       # the reference's name for the MLP width
            kw["intermediate_size_mlp"] = cfg.intermediate_size
        return cls(**kw)
