"""Drop the B200 encoder into an unmodified reference checkout.

    import efficientconformer_b200 as ec; ec.patch_reference()     # before `from functions import create_model`
    python main.py -c configs/EfficientConformerCTCSmall.json --mode validation-clean --gready   # unchanged CLI

`models.model_ctc.ModelCTC.__init__` looks `ConformerEncoder` up in its own module namespace
(reference models/model_ctc.py:28,43-44) and `models.transducer` does the same (models/transducer.py:59), so rebinding
the name in those modules (and in models.encoders) is all that is needed."""
import importlib
import sys


def patch_reference(modules=("models.encoders", "models.model_ctc", "models.transducer")):
    from .encoders import ConformerEncoder
    patched = []
    for name in modules:
        try:
            mod = sys.modules.get(name) or importlib.import_module(name)
        except Exception:
            continue
        if hasattr(mod, "ConformerEncoder"):
            setattr(mod, "ConformerEncoder", ConformerEncoder)
            patched.append(name)
    return patched
