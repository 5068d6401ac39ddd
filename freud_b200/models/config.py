"""SAE hyper-parameter records with the names, field order, defaults and `from_dict` / `to_dict` behaviour that the
reference's callers rely on (src/models/config.py:5-28; used at train_sae.py:357,360 and
dataset/activations.py:21-27).  The reference gets (de)serialisation from simple_parsing; here the three records
are generated from one field table and carry their own two-method serialiser.  `from_dict` ignores keys that are
not fields, because the train configs keep `dead_feature_threshold` in the same JSON object
(configs/train/tiny_topk.json:13)."""
import dataclasses

# (field, type, default, meaning) -- positional order is part of the interface
_COMMON = (
    ("expansion_factor", int, 32, "dictionary size as a multiple of the activation size"),
    ("n_dict_components", int, 0, "explicit dictionary size; 0 means activation_size * expansion_factor"),
)
_L1_ONLY = (
    ("recon_alpha", float, 1.0, "weight of the masked-MSE reconstruction term against the L1 term"),
)
_TOPK_ONLY = (
    ("normalize_decoder", bool, True, "unit-norm the decoder rows at construction"),
    ("k", int, 32, "latents kept per token"),
    ("multi_topk", bool, False, "add the 4k-selection FVU / 8 to the loss"),
    ("auxk_alpha", float, 0.0, "weight of the dead-latent (AuxK) loss"),
)


class _Record:
    """from_dict / to_dict of the generated records."""

    @classmethod
    def from_dict(cls, d, drop_extra_fields=True):
        known = [f.name for f in dataclasses.fields(cls)]
        return cls(**{name: d[name] for name in known if name in d})

    def to_dict(self):
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}


def _record(name, base, table):
    cls = dataclasses.make_dataclass(name, [(n, t, dataclasses.field(default=v)) for n, t, v, _ in table],
                                     bases=(base,), module=__name__)
    cls.__doc__ = name + ": " + "; ".join(f"{n} = {v!r} ({why})" for n, _, v, why in table)
    return cls


AutoEncoderConfig = _record("AutoEncoderConfig", _Record, _COMMON)
L1AutoEncoderConfig = _record("L1AutoEncoderConfig", AutoEncoderConfig, _L1_ONLY)
TopKAutoEncoderConfig = _record("TopKAutoEncoderConfig", AutoEncoderConfig, _TOPK_ONLY)
