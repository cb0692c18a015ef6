"""CPU: the C-ABI library loads and exports every symbol include/dupl.h declares (no compute without a
GPU), the ctypes structs mirror the header, and the host-side modules keep the reference's interface."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dupl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dupl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dupl_b200 import _lib
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dupl.h but not exported by libdupl.so"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.dupl_version() == 1


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof() of every args struct as seen by gcc equals the ctypes mirror."""
    from dupl_b200 import _lib
    names = {"dupl_segment": _lib.Segment, "dupl_gemm_group": _lib.GemmGroup, "dupl_gemm_args": _lib.GemmArgs,
             "dupl_attention_args": _lib.AttentionArgs, "dupl_mscam_args": _lib.MscamArgs,
             "dupl_cam_to_label_args": _lib.CamToLabelArgs, "dupl_refine_prologue_args": _lib.RefinePrologueArgs,
             "dupl_refine_epilogue_args": _lib.RefineEpilogueArgs, "dupl_attention_bwd_args": _lib.AttentionBwdArgs,
             "dupl_crf_args": _lib.CrfArgs, "dupl_adamw_param": _lib.AdamwParam, "dupl_adamw_args": _lib.AdamwArgs,
             "dupl_transpose_item": _lib.TransposeItem}
    prog = '#include <stdio.h>\n#include "dupl.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    c = tmp_path / "sz.c"
    c.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = dict(zip(out[::2], map(int, out[1::2])))
    for n, cls in names.items():
        assert ctypes.sizeof(cls) == sizes[n], n


def test_invalid_arguments_are_reported_without_a_gpu():
    from dupl_b200 import _lib
    lib = _lib.lib()
    rc = lib.dupl_gemm_bf16x3(None, None)
    assert rc == -1 and b"NULL" in lib.dupl_last_error()
    a = _lib.GemmArgs()
    a.groups, a.M, a.N, a.K, a.lda, a.ldo = 1, 8, 16, 60, 64, 16
    assert lib.dupl_gemm_bf16x3(ctypes.byref(a), None) == -1
    assert b"multiple of 64" in lib.dupl_last_error()


def test_product_path_refuses_cpu_tensors():
    from dupl_b200.model.PAR import PAR
    from dupl_b200.utils import cam_helper
    par = PAR(num_iter=1, dilations=[1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        par(torch.rand(1, 3, 8, 8), torch.rand(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cam_helper.cam_to_label(torch.rand(1, 3, 4, 4), torch.ones(1, 3), bkg_thre=0.5)


def test_model_keeps_the_reference_interface():
    """state-dict schema, parameter groups (204/100/4/6 tensors), zero buffers — SURVEY §8(b)."""
    from helpers import init_state_dict
    from dupl_b200.model.model_dupl import siamese_network
    m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    assert sum(p.numel() for p in m.parameters()) == 185014736
    assert list(m.named_buffers()) == []
    assert [len(g) for g in m.get_param_groups()] == [204, 100, 4, 6]
    assert [sum(p.numel() for p in g) for g in m.get_param_groups()] == [173058512, 76800, 61440, 11817984]
    P = init_state_dict(21)
    assert set(P) == set(m.state_dict())
    m.load_state_dict(P, strict=True)
    assert not m.branch1.encoder.pos_embed.requires_grad
    assert m.branch2.encoder.aux_block_index() == 9


def test_par_module_keeps_the_reference_attributes():
    from dupl_b200.model.PAR import PAR
    par = PAR(dilations=[1, 2, 4, 8, 12, 24], num_iter=10)
    assert par.kernel.shape == (8, 1, 3, 3) and list(par.state_dict()) == ["kernel"]
    assert par.pos.shape == (1, 1, 48, 1, 1) and (par.dim, par.w1, par.w2) == (2, 0.3, 0.01)


def test_segments_are_packed():
    from dupl_b200 import ops
    segs, M, Mp = ops.make_segments([(8, 28, 28), (8, 14, 14), (8, 42, 42)])
    assert (M, Mp) == (8 * (785 + 197 + 1765), 8 * (784 + 196 + 1764))
    assert [s.row_offset for s in segs] == [0, 6280, 7856] and [s.patch_row_offset for s in segs] == [0, 6272, 7840]


def test_eval_sweep_shards_images_like_the_reference_tool():
    """tools/eval_seg_coco_ddp.py:241: rank r takes images r, r+world, ...; every image exactly once."""
    from dupl_b200.eval_sweep import shard_indices
    for n, world in ((10, 1), (10, 4), (7, 8), (0, 2)):
        parts = [shard_indices(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert all(p == list(range(r, n, world)) for r, p in enumerate(parts))


def test_gemm_plan_fills_the_74_cta_pairs_for_the_training_shapes():
    """Host-side tiling choice (no GPU needed): 148 SMs = 74 CTA pairs; the M = 3140 training shapes must not be left on a
    fraction of them, wgrad shapes (few output tiles, K = 3200) must split K, the big MS-CAM shapes keep 256-wide tiles."""
    import ctypes as C
    from dupl_b200 import _lib as L

    def plan(M, N, K, groups=1, max_ksplit=0):
        bn, ks, items = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(L.lib().dupl_gemm_plan(M, N, K, groups, max_ksplit, C.byref(bn), C.byref(ks), C.byref(items)), "dupl_gemm_plan")
        return bn.value, ks.value, items.value

    assert plan(21976, 3072, 768)[:2] == (256, 1)               # fc1 of the MS-CAM pass: 1032 tiles, 14 waves
    bn, ks, items = plan(3140, 768, 768)                          # proj of the training pass: 39 tiles of 256 would use 53 %
    assert bn < 256 and ks == 1 and items > 39
    bn, ks, items = plan(768, 768, 3200, max_ksplit=8)            # wgrad of proj: 9 output tiles
    assert ks > 1 and 37 <= items <= 148
    assert plan(768, 768, 3200, max_ksplit=0)[1] == 1             # split-K only when a workspace was offered
    bn, ks, items = plan(3072, 768, 3200, max_ksplit=8)           # wgrad of fc1
    assert items >= 70
    assert plan(200, 32, 768)[0] == 64 and plan(260, 128, 256)[0] == 128   # narrow heads keep the narrow instantiations
    with __import__("pytest").raises(RuntimeError):
        plan(16, 16, 60)


def test_mn_major_gemm_argument_rules_without_a_gpu():
    """include/dupl.h: a ragged K is accepted only when BOTH operands are MN-major; an MN-major W needs N >= 128; the leading
    dimensions then count the transposed storage ([K, M] / [K, N])."""
    from dupl_b200 import _lib
    lib = _lib.lib()
    a = _lib.GemmArgs()
    a.groups, a.M, a.N, a.K, a.epilogue = 1, 768, 768, 3140, _lib.EPI_F32
    a.lda, a.ldo, a.ldw = 768, 768, 768
    a.a_mn_major, a.b_mn_major = 1, 0
    assert lib.dupl_gemm_bf16x3(ctypes.byref(a), None) == -1 and b"multiple of 64" in lib.dupl_last_error()
    a.a_mn_major, a.b_mn_major, a.N, a.ldo, a.ldw = 1, 1, 64, 64, 64
    assert lib.dupl_gemm_bf16x3(ctypes.byref(a), None) == -1 and b"N >= 128" in lib.dupl_last_error()
    a.N, a.ldo, a.ldw, a.lda = 768, 768, 768, 760                      # lda < M for the [K, M] storage
    assert lib.dupl_gemm_bf16x3(ctypes.byref(a), None) == -1 and b"lda" in lib.dupl_last_error()
    a.lda = 768                                                          # now well-formed: fails only on the NULL planes
    assert lib.dupl_gemm_bf16x3(ctypes.byref(a), None) == -1 and b"NULL operand plane" in lib.dupl_last_error()


def test_bench_arms_share_one_config():
    """bench.py: `config` of the training-step line is built by one function for this repo's arm and for --impl reference (the
    driver compares the two lines); nothing implementation-specific lives in it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c1 = bench.train_config("voc", 20, "B", 1)
    assert c1 == {"workload": "voc21_dual_student_phaseB_step_448_bs4", "per_gpu_batch": 4, "image": 448, "classes": 21,
                  "parallelism": "single GPU", "l2_policy": c1["l2_policy"]}
    assert "L2" in c1["l2_policy"]
    assert bench.train_config("coco", 80, "C", 8)["parallelism"] == "dp8"
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("train_config(args.dataset, K,") == 2              # both arms call it
