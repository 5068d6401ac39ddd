"""Import shims that let the *unmodified* reference import in this image.

Used only by ``tests/golden/make_golden.py`` and by tests that compare the
oracle with the live reference when ``/root/reference`` is present (build
container only; the GPU box has no reference tree).

* ``simple_parsing`` is not installed -> a module exposing ``Serializable`` with
  ``from_dict`` (keeps dataclass fields, drops unknown keys such as
  ``dead_feature_threshold``) and ``to_dict``  (needed by src/models/config.py:2).
* ``whisper`` is not installed -> empty module (src/models/hooked_model.py:7 is
  imported by l1autoencoder.py:7 at module import; no symbol is used on the SAE path).
* ``trim_activation`` (src/utils/activations.py:19-29) decodes audio; it is
  replaced by the same arithmetic fed from a {filename: num_samples} table.
"""
import dataclasses
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """The reference tree: $FREUD_REFERENCE_ROOT, else /root/reference (build container), else oracle/_ref -- the
    copy oracle/make_ref.sh makes so that the unmodified reference travels to the GPU box."""
    for cand in (os.environ.get("FREUD_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "models")):
            return cand
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def install():
    if "simple_parsing" not in sys.modules:
        sp = types.ModuleType("simple_parsing")

        class Serializable:
            @classmethod
            def from_dict(cls, d, drop_extra_fields=True):
                names = {f.name for f in dataclasses.fields(cls)}
                return cls(**{k: v for k, v in d.items() if k in names})

            def to_dict(self):
                return dataclasses.asdict(self)

        sp.Serializable = Serializable
        sys.modules["simple_parsing"] = sp
    if "whisper" not in sys.modules:
        sys.modules["whisper"] = types.ModuleType("whisper")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def patch_trim(num_samples_by_file, sample_rate=16000):
    """Replace src.utils.activations.trim_activation with its own arithmetic
    (utils/activations.py:26-29) on a length table instead of an audio decode."""
    install()
    import src.utils.activations as ua
    from src.utils.constants import TIMESTEP_S

    def trim(fname, act):
        dur = num_samples_by_file[fname] / sample_rate
        return act[: int(dur / TIMESTEP_S)]

    ua.trim_activation = trim
    return ua
