"""Make the reference's own scripts import this build: `install_as_src()` registers the freud_b200 mirrors under
the reference's module paths (src.models.config, src.models.l1autoencoder, src.models.topkautoencoder,
src.utils.models, src.utils.constants) and patches src.utils.activations.top_activations /
src.dataset.activations.init_sae_from_checkpoint, so `python -m src.scripts.train_sae` and `gui_server` run on the
CUDA kernels unmodified -- and `torch.save(model, ...)` pickles (train_sae.py:594-595) resolve by class path.
See INTEGRATION.md."""
import importlib
import sys


def install_as_src(patch_search: bool = True):
    from .models import config, l1autoencoder, topkautoencoder
    from .utils import constants, models

    mapping = {
        "src.models.config": config,
        "src.models.l1autoencoder": l1autoencoder,
        "src.models.topkautoencoder": topkautoencoder,
        "src.utils.models": models,
    }
    for name, mod in mapping.items():
        sys.modules[name] = mod
    if patch_search:
        try:
            ua = importlib.import_module("src.utils.activations")
            from .utils import activations as ours

            ua.top_activations = ours.top_activations
        except ImportError:
            pass  # reference tree not on sys.path: only the model classes are aliased
    return mapping
