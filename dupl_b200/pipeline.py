"""The CAM -> PAR -> pseudo-label step of the training loop as ONE call:
what train_final_voc.py:263-343 does between `inputs.to(device)` and the segmentation loss, with the
script-side glue (denormalize_img2, cls_label broadcast, threshold map) included so that callers who
do not keep the reference's inline loop get the same labels from the same inputs.

Nothing here is new arithmetic: it sequences the drop-in functions of utils/cam_helper.py.
"""
import torch

from .model.PAR import PAR
from .utils import cam_helper

IMG_MEAN = (123.675, 116.28, 103.53)
IMG_STD = (58.395, 57.12, 57.375)


def denormalize_img2(imgs):
    """utils/imutils.py:17-31 (script-side, feeds PAR): x*std+mean -> uint8 truncation -> /255.
    Plain torch elementwise ops, exactly as the reference's caller executes them."""
    out = torch.zeros_like(imgs)
    for c in range(3):
        out[:, c] = imgs[:, c] * IMG_STD[c] + IMG_MEAN[c]
    return out.type(torch.uint8) / 255.0


class CamParStep:
    """multi_scale_cam2_siamese for both students + refine_cams_with_dynamic_thres for both students."""

    def __init__(self, model, cam_scales=(1.0, 0.5, 1.5), low_thre=0.25, ignore_index=255,
                 dilations=(1, 2, 4, 8, 12, 24), num_iter=10, fuse_students=True, graph=False, keep_activations=False,
                 refine="dynamic", scalar_high_thre=None):
        self.model = model
        self.scales = tuple(cam_scales)
        self.low_thre = low_thre
        self.ignore_index = ignore_index
        self.par = PAR(num_iter=num_iter, dilations=list(dilations))  # train_final_voc.py:160
        # both students per grouped GEMM launch (one encoder pass over [student][segment] groups): 3 % faster than two passes
        # on B200 (402 vs 376 TFLOP/s over the GEMM launches: twice the tiles per launch fill the 74 CTA pairs better)
        self.fuse_students = fuse_students
        self.graph = graph      # replay the whole step as ONE CUDA graph (the ~220 launches cost the host nothing)
        # keep the encoder activations of the un-flipped scale-1.0 images on each student (network._kept) so that the
        # training forward of the same step does not recompute them (encoder.KeptActivations)
        self.keep_activations = keep_activations
        # refine: "dynamic"    refine_cams_with_dynamic_thres on `cams`   (train_final_voc.py:330-343, coco n_iter > 12000)
        #         "aux_scalar" refine_cams_with_bkg_v2 on `cams_aux` with a scalar high threshold (train_final_coco.py:312-322)
        #         None         MS-CAM only (phase A: no pseudo-labels yet)
        if refine not in ("dynamic", "aux_scalar", None):
            raise ValueError("refine must be 'dynamic', 'aux_scalar' or None")
        self.refine = refine
        self.scalar_high_thre = scalar_high_thre
        self._g = None

    def __call__(self, inputs, cls_label, img_box, high_thres):
        if not self.graph:
            return self._run(inputs, cls_label, img_box, high_thres)
        return self._replay(inputs, cls_label, img_box, high_thres)

    @torch.no_grad()
    def _replay(self, inputs, cls_label, img_box, high_thres):
        dev = inputs.device
        key = (tuple(inputs.shape), tuple(cls_label.shape), dev)
        if self._g is None or self._g["key"] != key:
            st = dict(key=key, x=torch.empty_like(inputs), cls=torch.empty_like(cls_label, dtype=torch.float32),
                      box=torch.empty(inputs.shape[0], 4, dtype=torch.int32, device=dev),
                      thr=torch.empty(inputs.shape[0], dtype=torch.float32, device=dev))
            self._stage(st, inputs, cls_label, img_box, high_thres)
            net = self.model.module if hasattr(self.model, "module") else self.model
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # warm-up outside capture: caches, lazy attributes, allocator pools
                for _ in range(2):
                    self._run(st["x"], st["cls"], st["box"], st["thr"])
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for b in (net.branch1, net.branch2):
                    b.planes().refresh_all()    # replays see the weights of the moment
                st["out"] = self._run(st["x"], st["cls"], st["box"], st["thr"])
            st["graph"] = graph
            # the kept activations this graph writes: several CamParStep graphs can share one model (TrainStep owns three),
            # so every replay must re-point the students at ITS buffers, not at those of the graph captured last
            st["kept"] = (net.branch1._kept, net.branch2._kept)
            self._g = st
        st = self._g
        self._stage(st, inputs, cls_label, img_box, high_thres)
        st["graph"].replay()
        if self.keep_activations:
            net = self.model.module if hasattr(self.model, "module") else self.model
            net.branch1._kept, net.branch2._kept = st["kept"]
        return st["out"]

    @staticmethod
    def _stage(st, inputs, cls_label, img_box, high_thres):
        st["x"].copy_(inputs, non_blocking=True)
        st["cls"].copy_(cls_label, non_blocking=True)
        st["box"].copy_(torch.as_tensor(img_box).to(torch.int32), non_blocking=True)
        if high_thres is not None:
            st["thr"].copy_(high_thres, non_blocking=True)

    @torch.no_grad()
    def _run(self, inputs, cls_label, img_box, high_thres):
        """inputs [b,3,H,W] normalised, cls_label [b,K], img_box [b,4] (CPU int16 as the loader yields it),
        high_thres [b] per-image high threshold (train_final_voc.py:263-275).
        Returns (label_1, label_2, cams_1, cams_2): refined labels float32 [b,H,W] in {0..K,255}."""
        b, _, h, w = inputs.shape
        inputs_denorm = denormalize_img2(inputs.clone())
        net = self.model.module if hasattr(self.model, "module") else self.model
        for n in (net.branch1, net.branch2):
            n._keep_next = bool(self.keep_activations) and 1.0 in self.scales  # scale 1.0 is always segment 0
        if self.fuse_students:
            (cams_1, aux_1), (cams_2, aux_2) = cam_helper.multi_scale_cam2_pair(self.model, inputs, self.scales)
        else:
            cams_1, aux_1 = cam_helper.multi_scale_cam2_siamese(self.model, inputs, self.scales, branch=1)
            cams_2, aux_2 = cam_helper.multi_scale_cam2_siamese(self.model, inputs, self.scales, branch=2)
        # train_final_voc.py:330-335 multiplies the CAMs by the broadcast cls_label before refining.
        # cls_label is one-hot {0,1}: present classes are multiplied by exactly 1.0 and absent classes are
        # never read by the refine kernels, so the 64 MB elementwise pass is skipped without changing a bit.
        for n in (net.branch1, net.branch2):
            n._keep_next = False
        if self.refine is None:
            return None, None, (cams_1, aux_1), (cams_2, aux_2)
        # both students refine over the same images: one PAR affinity for the two calls (cam_helper.refine_cams_shared_affinity)
        if self.refine == "aux_scalar":
            lab_1, lab_2 = cam_helper.refine_cams_shared_affinity(self.par, inputs_denorm, [aux_1, aux_2], cls_label,
                                                                  float(self.scalar_high_thre), self.low_thre, self.ignore_index, img_box)
            return lab_1, lab_2, (cams_1, aux_1), (cams_2, aux_2)
        thr_map = high_thres.to(inputs.device, torch.float32).reshape(b, 1, 1, 1).expand(b, 1, h, w).contiguous()
        lab_1, lab_2 = cam_helper.refine_cams_shared_affinity(self.par, inputs_denorm, [cams_1, cams_2], cls_label, thr_map,
                                                              self.low_thre, self.ignore_index, img_box)
        return lab_1, lab_2, (cams_1, aux_1), (cams_2, aux_2)
