import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dupl_b200.model.losses import _PtcLoss, _PtcLossSimt
from dupl_b200.utils import cam_helper
for b in (3,):
    g = torch.Generator().manual_seed(796)
    fmap = torch.randn(b, 768, 28, 28, generator=g).cuda()
    lab = torch.randint(0, 4, (b, 28, 28), generator=g); lab[lab == 3] = 255
    aff = cam_helper.label_to_aff_mask(lab.cuda())
    res = []
    for fn in (_PtcLoss, _PtcLossSimt):
        f = fmap.clone().requires_grad_(True)
        loss = fn.apply(f, aff)
        saved = loss.grad_fn.saved_tensors
        (loss * 0.5).backward()
        res.append((loss.item(), f.grad.clone(), saved[3].clone()))
    G0, G1 = res[0][2], res[1][2]
    xh = torch.nn.functional.normalize(fmap.double().reshape(b, 768, -1), dim=1)
    Gd = torch.einsum("bcp,bcq->bpq", xh, xh)
    idx = ((G0 > 0) != (G1 > 0)).nonzero()
    print("flips", idx.tolist())
    for i, p, q in idx.tolist():
        print("  tc", G0[i, p, q].item(), "simt", G1[i, p, q].item(), "fp64", Gd[i, p, q].item(), "mask", aff[i, p, q].item())
    print("max |G_tc - G64|", (G0.double() - Gd).abs().max().item(), "max |G_simt - G64|", (G1.double() - Gd).abs().max().item())
    offd = ~torch.eye(784, dtype=torch.bool, device="cuda")[None].expand(b, -1, -1)
    print("offdiag max |G_tc - G64|", (G0.double() - Gd)[offd].abs().max().item(), (G1.double() - Gd)[offd].abs().max().item())
    per_img = [((res[0][1][i] - res[1][1][i]).norm() / res[1][1][i].norm()).item() for i in range(b)]
    print("grad rel per image", per_img)
