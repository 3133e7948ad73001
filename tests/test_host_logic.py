"""CPU-only checks of the host side: the drop-in module's state_dict equals the reference's (names, shapes, order),
the C-ABI library loads and exports every symbol declared in include/effconf_b200.h, shape bookkeeping, loud failures."""
import json
import os
import re

import pytest
import torch

from efficientconformer_b200.config import CTC_SMALL_ENCODER_PARAMS as P, CTC_SMALL_VOCAB as V, stage_lengths, resolve_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_state_dict_equals_reference(golden_dir):
    from efficientconformer_b200 import ModelCTC
    layouts = json.load(open(os.path.join(golden_dir, "state_dict_layouts.json")))
    ref = layouts["EfficientConformerCTCSmall"]["keys"]
    m = ModelCTC(P, {"vocab_size": V})
    mine = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert mine == ref
    # Medium config (different dims / block count) through the same holders
    med = layouts["EfficientConformerCTCMedium"]
    m2 = ModelCTC(med["encoder_params"], {"vocab_size": med["vocab_size"]})
    assert [[k, list(v.shape)] for k, v in m2.state_dict().items()] == med["keys"]
    assert [b.stride for b in m.encoder.blocks] == [1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1]


def test_library_exports_every_declared_symbol():
    import ctypes
    from efficientconformer_b200 import _lib
    header = open(os.path.join(ROOT, "include", "effconf_b200.h")).read()
    declared = set(re.findall(r"\b(ec_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(_lib.LIB_PATH)          # built in-tree by __graft_entry__.build()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    L = _lib.lib()
    assert L.ec_version() == 100


def test_engine_shape_queries_without_gpu():
    """Pure host logic of the C library (no CUDA calls): workspace / weight sizes and frame bookkeeping."""
    import ctypes as C
    from efficientconformer_b200 import ConformerEncoder, _lib
    enc = ConformerEncoder(P)
    cfg = enc._config_struct()
    L = _lib.lib()
    h = C.c_void_p()
    _lib.check(L.ec_engine_create(C.byref(cfg), _lib.PREC_TF32, C.byref(h)))
    try:
        for T in (1, 2, 7, 500, 999, 1000, 1600):
            lens, t_final = stage_lengths(P, T)
            assert L.ec_engine_out_frames(h, T) == t_final
            rows = (C.c_int32 * 15)(); frames = (C.c_int32 * 15)()
            _lib.check(L.ec_engine_relpos_rows(h, T, rows, frames))
            assert list(frames) == lens
            for i, s in enumerate(resolve_blocks(P)):
                tp = lens[i] + (-lens[i]) % s.group_size
                assert rows[i] == 2 * tp - s.group_size
        wb = L.ec_engine_weight_bytes(h)
        assert 13_281_856 * 4 * 0.95 < wb < 13_281_856 * 4 * 1.2      # fp32-storage arena ~ parameter bytes
        ws = L.ec_engine_workspace_bytes(h, 32, 1000)
        assert 300e6 < ws < 800e6
    finally:
        L.ec_engine_destroy(h)
    bad = _lib.Config()
    assert L.ec_engine_create(C.byref(bad), _lib.PREC_TF32, C.byref(h)) == 1
    assert b"num_blocks" in L.ec_last_error()


def test_relative_table_matches_oracle():
    from efficientconformer_b200.encoders import relative_sinusoid_rows
    from oracle.conformer_oracle import relative_sinusoid_rows as oracle_rows
    for (tp, d, g, ml) in [(501, 120, 3, 10000), (250, 168, 1, 5000), (125, 240, 1, 2500), (3, 120, 3, 10000)]:
        assert torch.equal(relative_sinusoid_rows(tp, d, g, ml), oracle_rows(tp, d, g, ml))


def test_product_fails_loudly_without_gpu():
    from efficientconformer_b200 import ConformerEncoder
    from efficientconformer_b200 import ops
    enc = ConformerEncoder(P)
    with pytest.raises(RuntimeError):                     # neither mode has a CPU path
        enc.train().forward_mel(torch.zeros(1, 80, 16))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):                    # the operators themselves reject host memory / a missing device
            ops.cast(torch.zeros(4), "bf16")
    with pytest.raises(RuntimeError):
        enc.eval().forward_mel(torch.zeros(1, 80, 16))
    with pytest.raises(NotImplementedError):
        ConformerEncoder(dict(P, subsampling_module="VGG"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "efficientconformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
