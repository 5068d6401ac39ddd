"""Make the reference's own scripts run on this build.

`install_as_src()` registers the freud_b200 mirrors under the reference's module paths
    src.models.config, src.models.l1autoencoder, src.models.topkautoencoder, src.utils.models
so that `python -m src.scripts.train_sae` / `gui_server` construct OUR classes, and `torch.save(model, ...)` pickles
(train_sae.py:594-595) resolve by class path.  With the reference tree importable it additionally
  * rebinds the class names inside reference modules that were imported before the call
    (src.dataset.activations -> init_sae_from_checkpoint / FlyActivationDataLoader isinstance checks,
     src.scripts.train_sae -> train() / validate()),
  * replaces src.utils.activations.top_activations (the `top_fn` of gui_server.py:91-99) by the CUDA search,
  * replaces src.scripts.train_sae.topk_feature_extraction (validate(), train_sae.py:70-118,180-182) by the
    scatter-max kernel,
  * replaces the `Adam` / `RAdam` names train() resolves (train_sae.py:20,374-381) by the fused multi-tensor
    optimisers (same constructor arguments, same state_dict layout); `clip_grad_norm_` stays torch's own (it is
    called through `torch.nn.utils`, train_sae.py:449, and works on the CUDA gradients as is).
`src.utils.constants` is left alone (train_sae.py needs its get_n_mels).  See INTEGRATION.md."""
import importlib
import sys

_CLASS_NAMES = ("AutoEncoderConfig", "L1AutoEncoderConfig", "TopKAutoEncoderConfig", "L1AutoEncoder", "L1EncoderOutput",
                "L1ForwardOutput", "mse_loss", "TopKAutoEncoder", "TopKEncoderOutput", "TopKForwardOutput",
                "eager_decode", "get_n_dict_components")


def install_as_src(patch_search: bool = True, patch_validation: bool = True, patch_optim: bool = True):
    from .models import config, l1autoencoder, topkautoencoder
    from .utils import models

    mapping = {
        "src.models.config": config,
        "src.models.l1autoencoder": l1autoencoder,
        "src.models.topkautoencoder": topkautoencoder,
        "src.utils.models": models,
    }
    ours = {}
    for mod in mapping.values():
        for name in _CLASS_NAMES:
            if hasattr(mod, name):
                ours[name] = getattr(mod, name)
    for name, mod in mapping.items():
        sys.modules[name] = mod
    # reference modules imported BEFORE this call hold the reference classes by name: rebind them
    for name, mod in list(sys.modules.items()):
        if mod is None or not name.startswith("src.") or name in mapping:
            continue
        for cname, obj in ours.items():
            if cname in getattr(mod, "__dict__", {}):
                setattr(mod, cname, obj)
    if patch_search:
        try:
            ua = importlib.import_module("src.utils.activations")
            from .utils import activations as our_search

            ua.top_activations = our_search.top_activations
        except ImportError:
            pass  # reference tree not on sys.path: only the model classes are aliased
    if patch_validation:
        try:
            ts = importlib.import_module("src.scripts.train_sae")
            from .utils import validation

            ts.topk_feature_extraction = validation.topk_feature_extraction
        except ImportError:
            pass
    if patch_optim:
        try:
            ts = importlib.import_module("src.scripts.train_sae")
            from . import optim

            ts.Adam, ts.RAdam = optim.FusedAdam, optim.FusedRAdam
        except ImportError:
            pass
    return mapping
