"""GPU parity tests of the dense path (tcgen05 GEMM, LayerNorm, attention, encoder) through the C ABI.
Oracle: torch fp64 on the same inputs for single kernels, oracle/dupl_oracle.py for the encoder."""
import pytest
import torch

from helpers import init_state_dict, mscam_err, rel_err, synth_images

pytestmark = pytest.mark.gpu


def _ops():
    from dupl_b200 import _lib as L, ops
    return L, ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def test_split_bf16_reconstructs_to_2e_minus_16():
    L, ops = _ops()
    x = _rand(1000, 37, seed=1)
    hi, lo = ops.split_bf16(x)
    rec = hi.float() + lo.float()
    assert ((rec - x).abs() <= x.abs() * 2.0 ** -16 + 1e-30).all()


@pytest.mark.parametrize("M,N,K,groups", [(128, 256, 64, 1), (300, 256, 128, 1), (1000, 768, 768, 2), (129, 2304, 768, 1),
                                           (517, 3072, 768, 2), (777, 768, 3072, 1), (260, 128, 256, 1), (200, 32, 768, 2)])
def test_gemm_f32_bias(M, N, K, groups):
    L, ops = _ops()
    gs, refs = [], []
    for g in range(groups):
        a, w, bias = _rand(M, K, seed=10 + g), _rand(N, K, seed=20 + g, scale=0.05), _rand(N, seed=30 + g)
        out = torch.full((M, N), float("nan"), device="cuda")
        gs.append(dict(a=ops.split_bf16(a), w=ops.split_bf16(w), bias=bias, out_f32=out))
        refs.append((a.double() @ w.double().t() + bias.double()))
    ops.gemm_bf16x3(gs, M, N, K, L.EPI_F32)
    torch.cuda.synchronize()
    for g in range(groups):
        assert torch.isfinite(gs[g]["out_f32"]).all()
        assert rel_err(gs[g]["out_f32"], refs[g]) < 3e-5


def test_gemm_epilogues():
    L, ops = _ops()
    M, N, K = 391, 768, 256
    a, w, bias, resid = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.1), _rand(N, seed=3), _rand(M, N, seed=4)
    ref = a.double() @ w.double().t() + bias.double()
    A, W = ops.split_bf16(a), ops.split_bf16(w)
    # SPLIT
    hi, lo = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm_bf16x3([dict(a=A, w=W, bias=bias, out=(hi, lo))], M, N, K, L.EPI_SPLIT)
    assert rel_err(hi.float() + lo.float(), ref) < 5e-5
    # GELU_SPLIT
    ops.gemm_bf16x3([dict(a=A, w=W, bias=bias, out=(hi, lo))], M, N, K, L.EPI_GELU_SPLIT)
    assert rel_err(hi.float() + lo.float(), torch.nn.functional.gelu(ref)) < 5e-5
    # RESID in place
    out = resid.clone()
    ops.gemm_bf16x3([dict(a=A, w=W, bias=bias, resid=out, out_f32=out)], M, N, K, L.EPI_RESID)
    assert rel_err(out, ref + resid.double()) < 3e-5


@pytest.mark.parametrize("M,N,K", [(3140, 768, 768), (3140, 3072, 768), (3140, 768, 3072), (6280, 2304, 768)])
def test_gemm_tile_width_chosen_by_the_cost_model(M, N, K):
    """The training shapes (M = 4 x 785) pick 192-wide tiles on a 148-SM part; results must not depend on the tiling."""
    L, ops = _ops()
    a, w, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05), _rand(N, seed=3)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_bf16x3([dict(a=ops.split_bf16(a), w=ops.split_bf16(w), bias=bias, out_f32=out)], M, N, K, L.EPI_F32)
    ref = a.double() @ w.double().t() + bias.double()
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 3e-5


@pytest.mark.parametrize("M,N,K", [(768, 768, 3200), (3072, 768, 3200), (2304, 768, 3200), (64, 512, 3136), (512, 6912, 3136)])
def test_gemm_split_k_is_exact_and_reproducible(M, N, K):
    """wgrad shapes: few output tiles, long contraction -> split-K through a workspace, fixed summation order."""
    L, ops = _ops()
    a, w = _rand(M, K, seed=5), _rand(N, K, seed=6, scale=0.05)
    A, W = ops.split_bf16(a), ops.split_bf16(w)
    outs = []
    for _ in range(2):
        out = torch.full((M, N), float("nan"), device="cuda")
        ops.gemm_bf16x3([dict(a=A, w=W, out_f32=out)], M, N, K, L.EPI_F32, ksplit=L.MAX_KSPLIT)
        outs.append(out)
    ref = a.double() @ w.double().t()
    assert torch.isfinite(outs[0]).all()
    assert rel_err(outs[0], ref) < 3e-5
    assert torch.equal(outs[0], outs[1])
    plain = torch.empty(M, N, device="cuda")
    ops.gemm_bf16x3([dict(a=A, w=W, out_f32=plain)], M, N, K, L.EPI_F32)
    assert rel_err(outs[0], plain.double()) < 3e-5


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,groups,ksplit",
                         [(768, 3072, 3140, True, True, 2, 8),     # wgrad fc2: dY^T [768, M] x hid [M, 3072], ragged K
                          (2304, 768, 3140, True, True, 1, 8),     # wgrad qkv
                          (512, 6912, 3136, True, True, 2, 8),     # wgrad conv6
                          (200, 128, 100, True, True, 1, 0),       # one 64-column block per CTA, two ragged k-blocks
                          (3140, 3072, 768, False, True, 2, 0),    # dgrad fc2: dY [M, 768] x W [768, 3072] in place
                          (3140, 768, 2304, False, True, 1, 0),    # dgrad qkv
                          (333, 192, 128, False, True, 1, 0),      # N not a multiple of the 128-wide tile
                          (768, 768, 640, True, False, 1, 0)])     # A alone
def test_gemm_mn_major_operands_in_place(M, N, K, a_mn, b_mn, groups, ksplit):
    """a_mn_major / b_mn_major: the operand is read from its transposed storage.  Same products in the same order as the
    K-major launch on transposed copies => bit-identical; and both match fp64."""
    L, ops = _ops()
    gs_mn, gs_k, refs = [], [], []
    Kp = (K + 63) // 64 * 64
    for g in range(groups):
        a, w = _rand(M, K, seed=40 + g), _rand(N, K, seed=50 + g, scale=0.05)
        A, W = ops.split_bf16(a), ops.split_bf16(w)
        pad = lambda pl: tuple(torch.nn.functional.pad(p, (0, Kp - K)).contiguous() for p in pl)   # noqa: E731
        tr = lambda pl: tuple(p.t().contiguous() for p in pl)                                        # noqa: E731
        o1 = torch.full((M, N), float("nan"), device="cuda")
        o2 = torch.full((M, N), float("nan"), device="cuda")
        gs_mn.append(dict(a=tr(A) if a_mn else A, w=tr(W) if b_mn else W, out_f32=o1))
        gs_k.append(dict(a=pad(A), w=pad(W), out_f32=o2))
        refs.append(a.double() @ w.double().t())
    ops.gemm_bf16x3(gs_mn, M, N, K, L.EPI_F32, ksplit=ksplit, a_mn=a_mn, b_mn=b_mn)
    ops.gemm_bf16x3(gs_k, M, N, Kp, L.EPI_F32, ksplit=ksplit)
    for g in range(groups):
        assert torch.isfinite(gs_mn[g]["out_f32"]).all()
        assert rel_err(gs_mn[g]["out_f32"], refs[g]) < 3e-5
        if ksplit == 0:     # every output element is one accumulator over the k-blocks in order, whatever the tile width
            assert torch.equal(gs_mn[g]["out_f32"], gs_k[g]["out_f32"])
        else:               # an MN-major W excludes the 192-wide tile, so the split-K factor may differ: same sums regrouped
            assert rel_err(gs_mn[g]["out_f32"], gs_k[g]["out_f32"].double()) < 2e-5


def test_gemm_gelu_side_output_row_limit():
    L, ops = _ops()
    M, N, K, keep = 700, 768, 256, 300
    a, w, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.1), _rand(N, seed=3)
    hi, lo = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"), torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    pre = torch.full((keep, N), float("nan"), device="cuda")
    guard = torch.full((M - keep, N), 7.0, device="cuda")  # would be overwritten if the limit were ignored
    buf = torch.cat([pre, guard])
    ops.gemm_bf16x3([dict(a=ops.split_bf16(a), w=ops.split_bf16(w), bias=bias, out=(hi, lo), out_f32=buf)], M, N, K,
                    L.EPI_GELU_SPLIT, f32_rows=keep)
    ref = a.double() @ w.double().t() + bias.double()
    assert rel_err(buf[:keep], ref[:keep]) < 3e-5
    assert (buf[keep:] == 7.0).all()
    assert rel_err(hi.float() + lo.float(), torch.nn.functional.gelu(ref)) < 5e-5


def test_gemm_rejects_bad_arguments():
    L, ops = _ops()
    a, w = _rand(16, 60), _rand(16, 60)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.gemm_bf16x3([dict(a=ops.split_bf16(a), w=ops.split_bf16(w), out_f32=torch.empty(16, 16, device="cuda"))],
                        16, 16, 60, L.EPI_F32)


def test_layernorm_split():
    L, ops = _ops()
    x = _rand(333, 768, seed=5) * 3 + 1
    g, b = _rand(768, seed=6), _rand(768, seed=7)
    hi = torch.empty(333, 768, dtype=torch.bfloat16, device="cuda")
    lo = torch.empty_like(hi)
    ops.layernorm_split(x, g, b, hi, lo, eps=1e-6)
    ref = torch.nn.functional.layer_norm(x.double(), (768,), g.double(), b.double(), eps=1e-6)
    assert rel_err(hi.float() + lo.float(), ref) < 3e-5


@pytest.mark.parametrize("shapes", [[(2, 3, 3)], [(1, 14, 14)], [(2, 8, 9), (1, 2, 2)], [(1, 28, 28)], [(2, 5, 7), (1, 14, 14), (3, 1, 1)]])
def test_attention_matches_fp64_softmax_attention(shapes):
    L, ops = _ops()
    segs, M, _ = ops.make_segments(shapes)
    qkv = _rand(M, 2304, seed=8)
    qh, ql = ops.split_bf16(qkv)
    oh = torch.zeros(M, 768, dtype=torch.bfloat16, device="cuda")
    ol = torch.zeros_like(oh)
    ops.attention_fwd(qh, ql, oh, ol, segs, 12, 0.125)
    out = oh.float() + ol.float()
    x = qkv.double()
    for s in segs:
        for i in range(s.batch):
            r0 = s.row_offset + i * s.tokens
            blk = x[r0:r0 + s.tokens].reshape(s.tokens, 3, 12, 64).permute(1, 2, 0, 3)
            att = torch.softmax(blk[0] @ blk[1].transpose(-1, -2) * 0.125, -1)
            ref = (att @ blk[2]).permute(1, 0, 2).reshape(s.tokens, 768)
            assert rel_err(out[r0:r0 + s.tokens], ref) < 1e-4


def _load_model(num_classes=21):
    from dupl_b200.model.model_dupl import siamese_network
    P = init_state_dict(num_classes)
    m = siamese_network("deit_base_patch16_224", num_classes=num_classes, pretrained=False, aux_layer=-3)
    m.load_state_dict(P, strict=True)
    return m.cuda().eval(), P


def test_cam_only_matches_oracle():
    """network.forward(cam_only=True) (model_dupl.py:69-84): CAM / aux-CAM within 1e-3 relative of fp32."""
    from oracle import dupl_oracle as O
    m, P = _load_model()
    x = synth_images(2, 64, 96, seed=3)
    with torch.no_grad():
        ca1, c1, ca2, c2 = m(x.cuda(), cam_only=True)
        oa1, o1 = O.network_cam_only(P, 1, x)
        oa2, o2 = O.network_cam_only(P, 2, x)
    for got, want in ((c1, o1), (ca1, oa1), (c2, o2), (ca2, oa2)):
        assert got.shape == want.shape
        assert rel_err(got, want) < 1e-3
    ca, c = m(x.cuda(), cam_only=True, branch=2)
    assert rel_err(c, o2) < 1e-3 and rel_err(ca, oa2) < 1e-3


def test_multi_scale_cam_matches_oracle():
    """multi_scale_cam2_siamese (cam_helper.py:164-204): normalised CAMs within 1e-3 absolute (range [0,1))."""
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    m, P = _load_model()
    x = synth_images(2, 64, 64, seed=4)
    with torch.no_grad():
        cam, aux = cam_helper.multi_scale_cam2_siamese(m, x.cuda(), (1.0, 0.5, 1.5), branch=1)
        ocam, oaux, osum, oaux_sum = O.multi_scale_cam(P, 1, x, (1.0, 0.5, 1.5), return_sums=True)
    assert cam.shape == ocam.shape == (2, 20, 64, 64)
    # tolerance 1e-3 of the CAM scale; planes the ReLU leaves ~constant amplify any error by 1/(range+1e-5)
    assert mscam_err(cam, ocam, osum) < 1e-3 and mscam_err(aux, oaux, oaux_sum) < 1e-3
    well = O.mscam_condition(osum) < 20
    assert ((cam.cpu() - ocam).abs() * well).max().item() < 1e-3
    (c1, a1), (c2, a2) = cam_helper.multi_scale_cam2_pair(m, x.cuda(), (1.0, 0.5, 1.5))
    assert torch.equal(c1, cam) and torch.equal(a1, aux)
    ocam2, _, osum2, _ = O.multi_scale_cam(P, 2, x, (1.0, 0.5, 1.5), return_sums=True)
    assert mscam_err(c2, ocam2, osum2) < 1e-3


def test_val_forward_matches_oracle():
    """siamese_network.forward(val=True) (model_dupl.py:160-168, 86-98): cls logits, seg logits, feature map and
    aux cls logits within 1e-3 relative of the fp32 oracle; also the single-branch call."""
    from oracle import dupl_oracle as O
    m, P = _load_model()
    x = synth_images(2, 96, 64, seed=5)
    with torch.no_grad():
        res = m(x.cuda(), val=True)
        for br in (1, 2):
            want = O.network_forward(P, br, x)
            got = res[f"branch{br}"]
            assert len(got) == 4
            for g, w in zip(got, want):
                assert g.shape == w.shape
                assert rel_err(g, w) < 1e-3
        one = m(x.cuda(), val=True, branch=2)
        assert torch.equal(one[1], res["branch2"][1])


def test_cam_with_grad_and_par_mask_resize_follow_the_reference():
    """The two interface corners no reference script exercises: siamese_network.forward(cam_with_grad=True)
    (model_dupl.py:100-104, 176-185) returns a fifth tensor computed from `_x4` like the reference; PAR.forward resizes masks of
    another size with align_corners=True (PAR.py:66) before propagating."""
    import torch.nn.functional as F
    from dupl_b200.model.PAR import PAR
    from oracle import dupl_oracle as O
    m, P = _load_model()
    x = synth_images(1, 64, 64, seed=6)
    with torch.no_grad():
        out = m(x.cuda(), cam_with_grad=True, branch=1)
        want = O.network_forward(P, 1, x)
    assert len(out) == 5 and out[4].shape == (1, 20, 4, 4)
    cg = F.conv2d(want[2], P["branch1.classifier.weight"])
    cg = cg + F.adaptive_max_pool2d(-cg, (1, 1))
    cg = cg / F.adaptive_max_pool2d(cg, (1, 1)) + 1e-5
    assert rel_err(out[4], cg) < 2e-3
    both = m(x.cuda(), cam_with_grad=True)
    assert set(both) == {"branch1", "branch2"} and len(both["branch2"]) == 5
    par = PAR(num_iter=3, dilations=[1, 2]).cuda()
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 3, 24, 20, generator=g)
    small = torch.rand(1, 3, 12, 10, generator=g).softmax(1)
    got = par(img.cuda(), small.cuda())
    ref = O.par_forward(img, F.interpolate(small, size=(24, 20), mode="bilinear", align_corners=True), dilations=(1, 2), num_iter=3)
    assert got.shape == (1, 3, 24, 20) and (got.cpu() - ref).abs().max().item() < 1e-5
