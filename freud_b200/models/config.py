"""Mirror of src/models/config.py:5-28 -- same dataclasses, fields and defaults.

The reference derives from simple_parsing.Serializable; only from_dict / to_dict are used by its callers
(train_sae.py:357,360; dataset/activations.py:21-27).  from_dict drops unknown keys (the train configs carry
`dead_feature_threshold` in the same dict, configs/train/tiny_topk.json:13)."""
import dataclasses
from dataclasses import dataclass


class _Serializable:
    @classmethod
    def from_dict(cls, d, drop_extra_fields=True):
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in dict(d).items() if k in names})

    def to_dict(self):
        return dataclasses.asdict(self)


@dataclass
class AutoEncoderConfig(_Serializable):
    expansion_factor: int = 32
    """Multiple of the input dimension to use as the SAE dimension."""
    n_dict_components: int = 0
    """Number of latents to use. If 0, use `expansion_factor`."""


@dataclass
class L1AutoEncoderConfig(AutoEncoderConfig):
    recon_alpha: float = 1.0
    """Weight of the reconstruction loss."""


@dataclass
class TopKAutoEncoderConfig(AutoEncoderConfig):
    normalize_decoder: bool = True
    """Whether to normalize the decoder weights to unit norm."""
    k: int = 32
    """Number of top latents to keep."""
    multi_topk: bool = False
    """Whether to use multi-topk."""
    auxk_alpha: float = 0.0
    """Weight of the auxk loss."""
