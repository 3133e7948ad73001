"""efficientconformer_b200 -- B200-native (sm_100a) Efficient Conformer encoder hot path behind the reference's API.

Public surface (mirrors burchim/EfficientConformer for the one accelerated path):
    ConformerEncoder(params)           reference models/encoders.py:44-142
    ModelCTC(...), LossCTC             reference models/model_ctc.py:37-136, models/losses.py:48-71
    patch_reference()                  swap the encoder class into an imported reference checkout (main.py unchanged)
    DeviceBatchPrefetcher(loader, dev) pinned, stream-overlapped replacement of the loop's `[elt.to(device) for elt in batch]`
Importing the package does not load CUDA; the shared library is loaded on first use and there is no fallback."""
from .config import resolve_blocks, state_dict_layout, CTC_SMALL_ENCODER_PARAMS, CTC_SMALL_VOCAB  # noqa: F401


def __getattr__(name):  # lazy: keeps `import efficientconformer_b200.config` torch-free and CPU-only friendly
    if name in ("ConformerEncoder",):
        from .encoders import ConformerEncoder
        return ConformerEncoder
    if name in ("ModelCTC", "LossCTC", "ctc_loss", "greedy_ids"):
        from . import model_ctc
        return getattr(model_ctc, name)
    if name == "patch_reference":
        from .dropin import patch_reference
        return patch_reference
    if name == "DeviceBatchPrefetcher":
        from .loader import DeviceBatchPrefetcher
        return DeviceBatchPrefetcher
    raise AttributeError(name)
