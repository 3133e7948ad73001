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


def test_flat_parameter_arena_and_operand_views_bookkeeping():
    """Host side of the native training step (efficientconformer_b200/trainer.py), no GPU: the flat arena keeps every parameter's
    values and state_dict name, Wq | Wk | Wv become one contiguous [3D, D] matrix, every GEMM weight gets a forward and a
    transposed operand view at its own offset, and the descriptor table of the multi-tensor transpose covers them exactly once."""
    from efficientconformer_b200.model_ctc import ModelCTC
    from efficientconformer_b200 import trainer
    from efficientconformer_b200.training import TrainingPath
    torch.manual_seed(0)
    m = ModelCTC(P, {"vocab_size": V})
    before = {k: v.clone() for k, v in m.state_dict().items()}
    path = TrainingPath(m.encoder, m.fc)
    named = trainer._qkv_adjacent_order(path.param_list())
    assert sorted(n for n, _ in named) == sorted(n for n, _ in path.param_list())
    flat = trainer.FlatParams(named, "cpu")
    after = m.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)      # same names, same values
    assert flat.total % 64 == 0 and all(o % 64 == 0 for o in flat.offsets)
    for n, p in named:                                                                             # parameters alias the arena
        i = flat.index[n]
        assert p.data_ptr() == flat.params[flat.offsets[i]:].data_ptr() and tuple(p.shape) == flat.shapes[i]
    with torch.no_grad():
        flat.params.add_(1.0)                                                                      # an optimiser step on the arena ...
    assert torch.equal(m.state_dict()["fc.bias"], before["fc.bias"] + 1.0)                         # ... is visible through the modules
    w = trainer.ArenaWeights(flat, "bf16", "cpu")
    assert w.supports(m.encoder) and len(w._qkv) == len(m.encoder.blocks)
    covered = torch.zeros(flat.total, dtype=torch.int32)
    for o, rows, cols, dst in w.desc.tolist():
        assert o == dst
        covered[o:o + rows * cols] += 1
    assert int(covered.max()) == 1                                                                 # no tensor transposed twice
    blk = m.encoder.blocks[4]
    mh = blk.multi_head_self_attention_module.mhsa
    D = mh.query_layer.weight.shape[0]
    assert w.qkv_act(mh).shape == (3 * D, D) and w.qkv_act_t(mh).shape == (D, 3 * D)
    assert w.qkv_act(mh).data_ptr() == w.act(mh.query_layer.weight).data_ptr()
    assert w.act(mh.key_layer.weight).data_ptr() == w.qkv_act(mh)[D:].data_ptr()
    for weight in (blk.conv_res[1].weight, blk.convolution_module.layers[2].weight, m.encoder.linear.weight, m.fc.weight):
        N, K = weight.shape[0], weight.numel() // weight.shape[0]
        assert w.act(weight).shape == (N, K) and w.act_t(weight).shape == (K, N)
    with pytest.raises(KeyError):
        w.act(blk.convolution_module.layers[4].weight)                                             # the depthwise taps are not a GEMM operand


def test_every_programmatically_launched_kernel_waits_for_its_predecessor():
    """Source-level invariant behind programmatic dependent launch (csrc/ec_common.cuh launch_pdl / launch_dep): a kernel launched
    with the attribute may start while its predecessor still runs, so its body must execute griddepcontrol.wait -- for the training kernels
    (launch_dep) before anything but a few declarations, with no pointer parameter mentioned ahead of it; for the hand-written
    pipelines (launch_pdl) somewhere before the first access to predecessor data (checked by the GPU parity tests).  A kernel added to a launch_dep call without the wait would race silently; this test fails instead."""
    src_dir = os.path.join(ROOT, "efficientconformer_b200", "csrc")
    texts = {f: open(os.path.join(src_dir, f)).read() for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))}

    def body_of(text, start):
        i = text.index("(", start)
        depth = 0
        while True:                                            # parameter list (skipping __launch_bounds__(...))
            depth += text[i] == "("; depth -= text[i] == ")"
            i += 1
            if depth == 0:
                j = i
                while text[j] in " \n\t":
                    j += 1
                if text[j] == "{":
                    break
                if text[j] == ";":
                    return None
                i = text.index("(", j)
        depth, k = 0, j
        while True:
            depth += text[k] == "{"; depth -= text[k] == "}"
            k += 1
            if depth == 0:
                return text[j + 1:k - 1]

    bodies = {}
    for f, text in texts.items():
        for m in re.finditer(r"__global__", text):
            b = body_of(text, m.end())
            if b is None:
                continue
            head = text[m.end():text.index("{", m.end())]
            head = re.sub(r"__launch_bounds__\s*\([^)]*\)", "", head)
            name = re.findall(r"(\w+)\s*\(", head)[0]
            pointers = set(re.findall(r"\*\s*(?:__restrict__\s+)?(\w+)\s*[,)/]", head))      # pointer parameters of the kernel
            bodies[name] = (b, pointers)
    launched_dep, launched_pdl = set(), set()
    for text in texts.values():
        launched_dep |= set(re.findall(r"launch_dep\(\s*(?:\w+::)*([A-Za-z_]\w*)", text))
        launched_pdl |= set(re.findall(r"launch_pdl(?:_cluster)?\(\s*(?:\w+::)*([A-Za-z_]\w*)", text))
    launched_dep -= {"kernel", "void"}; launched_pdl -= {"kernel", "void"}   # the helper templates' own declarations
    assert len(launched_dep) > 40 and len(launched_pdl) >= 8
    for k in sorted(launched_dep):
        assert k in bodies, k
        body, pointers = bodies[k]
        assert "grid_dependency_wait()" in body, k
        before = body.split("grid_dependency_wait()")[0]      # only declarations may precede the wait: no pointer parameter is touched
        before = re.sub(r"//[^\n]*", "", before)
        assert not any(re.search(r"\b%s\b" % p_, before) for p_ in pointers), (k, before)
        assert before.count(";") <= 3, (k, before)
    for k in sorted(launched_pdl):
        assert k in bodies and "grid_dependency_wait()" in bodies[k][0], k
    # and the spin-waiting peer-memory exchange is never launched programmatically
    assert "p2p_exchange_kernel" not in launched_dep | launched_pdl


def test_batch_prefetcher_refuses_non_cuda_devices():
    from efficientconformer_b200 import DeviceBatchPrefetcher
    with pytest.raises(RuntimeError, match="CUDA path only"):
        DeviceBatchPrefetcher([(torch.zeros(1),)], "cpu")
    with pytest.raises(ValueError):
        DeviceBatchPrefetcher([(torch.zeros(1),)], "cuda", depth=0)


def test_gradient_bucket_plan_covers_the_arena_in_backward_order():
    """Overlapped gradient buckets (trainer.CTCTrainStep._plan_buckets, EFFCONF_BUCKET_OVERLAP=1), no GPU: the arena is cut at block
    boundaries, bucket b holds exactly the head and the blocks >= b that no later-cut bucket holds, the buckets plus the remainder tile
    the parameter list without gaps or overlap, and a model whose arena order would break contiguity gets no plan (one bucket)."""
    from types import SimpleNamespace
    from efficientconformer_b200.model_ctc import ModelCTC
    from efficientconformer_b200 import trainer
    from efficientconformer_b200.training import TrainingPath
    m = ModelCTC(P, {"vocab_size": V})
    path = TrainingPath(m.encoder, m.fc)
    flat = trainer.FlatParams(trainer._qkv_adjacent_order(path.param_list()), "cpu")
    plan = trainer.CTCTrainStep._plan_buckets(SimpleNamespace(flat=flat), m)
    assert sorted(plan) == [5, 10]                                         # 15 blocks: the middle and the last third
    (lo1, hi1), (lo0, hi0) = plan[5], plan[10]
    assert hi0 == len(flat.names) and hi1 == lo0 and 0 < lo1 < lo0         # backward order: bucket 10 first, then 5, then [0, lo1)
    block_of = lambda n: int(n.split(".")[2]) if n.startswith("encoder.blocks.") else (99 if n.startswith("fc.") else -1)
    assert all(block_of(n) >= 10 for n in flat.names[lo0:hi0]) and all(5 <= block_of(n) < 10 for n in flat.names[lo1:hi1])
    assert all(block_of(n) < 5 for n in flat.names[:lo1])
    a0, b0 = flat.arena_range(lo0, hi0); a1, b1 = flat.arena_range(lo1, hi1); a2, b2 = flat.arena_range(0, lo1)
    assert (a2, b2, b1, b0) == (0, a1, a0, flat.total)                    # float ranges tile the arena
    assert 0.5 < (b0 - a0) / flat.total < 0.56                             # the last third of the blocks owns half of the bytes (D = 240)
    # an arena whose order interleaves the head with the encoder cannot be cut: no plan
    shuffled = SimpleNamespace(names=[flat.names[-1]] + flat.names[:-1])
    assert trainer.CTCTrainStep._plan_buckets(SimpleNamespace(flat=shuffled), m) == {}


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) in both modes on a tiny shape: one JSON line with
    the contract keys."""
    import subprocess
    import sys
    for mode in ("train", "forward"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mode", mode, "--steps", "1", "--warmup", "1",
                            "--batch", "2", "--frames", "120"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == "encoder_mel_frames_per_sec" and d["unit"] == "frames/s" and d["value"] > 0
        assert d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert "workload" in d["config"] and d["steps"] == 1


def test_bench_both_arms_describe_the_same_config():
    """The driver compares the `config` objects of the two arms: for the same flags they must be identical."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for mode in ("train", "forward"):
        for n in (1, 8):
            a = argparse.Namespace(config="EfficientConformerCTCSmall", batch=32, frames=1000, gpus=n, pdrop=0.1)
            c1, c2 = bench.workload_config(a, mode), bench.workload_config(argparse.Namespace(**vars(a)), mode)
            assert c1 == c2 and c1["global_batch"] == 32 * n and c1["frames"] == 1000 and "workload" in c1
    assert abs(bench.flops_per_frame(*bench.SHIPPED_ENCODER_PARAMS["EfficientConformerCTCSmall"], 1000, 32) / 1e6 - 6.616) < 1e-3
    assert abs(bench.flops_per_frame(*bench.SHIPPED_ENCODER_PARAMS["ConformerCTCLarge"], 4000, 8) / 1e6 - 99.888) < 1e-2


def test_oracle_block_shapes_are_independent_of_and_equal_to_the_product():
    """The oracle resolves the per-block hyper-parameters itself (no import from the product package); both agree on every shipped config."""
    import inspect
    from oracle import conformer_oracle as O
    from efficientconformer_b200.config import SHIPPED_ENCODER_PARAMS, resolve_blocks
    import re
    assert not re.search(r"^\s*(from|import)\s+efficientconformer_b200", inspect.getsource(O), re.M)
    for name, (params, _) in SHIPPED_ENCODER_PARAMS.items():
        a, b = O.resolve_blocks(params), resolve_blocks(params)
        assert len(a) == len(b) == params["num_blocks"]
        for x, y in zip(a, b):
            for f in ("dim_model", "dim_expand", "num_heads", "kernel_size", "group_size", "max_pos", "conv_stride", "ff_ratio", "dim_head", "has_conv_res_proj"):
                assert getattr(x, f) == getattr(y, f), (name, f)
