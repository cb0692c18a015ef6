#!/usr/bin/env python
"""Benchmarks of the DuPL hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload train|cam_par|crf_sweep]

BASELINE.json's metric is "train images/sec (dual-student step) at 1/2/4/8 B200; CAM+PAR+CRF ms/img".  The default
workload is therefore `train`: the full phase-B dual-student step (train_final_voc.py:186-472: MS-CAM of both students,
PAR pseudo-labels, forward / backward of both students with all losses, gradient average over the ranks, AdamW) on a
synthetic VOC batch, b = 4 per GPU, 448x448.  Its line carries `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches` and three secondary objects measured in the same run (rank 0, N = 1 only for the last two):
  `cam_par`        BASELINE.json configs[1]: MS-CAM + PAR refine of both students (images/s, ms/img, HBM-roofline fractions
                   of the PAR propagation and MS-CAM post-processing kernels);
  `crf`            DenseCRF mean-field ms/img at 640x480x81, T = 10, device-resident and through the numpy API, next to the
                   1-thread C restatement of pydensecrf;
  `reference_gpu`  the UNMODIFIED reference (baseline/_ref, its own modules on stock PyTorch: fp32, cudnn.deterministic)
                   executing the same loop body on the same GPU and batch — the north_star's ">= 5x" comparator.
Other workloads: `cam_par` alone, `crf_sweep` (configs[4]: COCO eval sweep per image, images strided over ranks),
`--dataset coco`, `--phase C`.

`--impl reference` times the reference's CPU path on the host cores (the unmodified reference from baseline/_ref when it
is there, else the oracle port), one 448x448 image per step, all host threads; rank 0 only under torchrun.  On a box with
a GPU its line also carries `reference_gpu`.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

BATCH = 4
SIZE = 448
SCALES = (1.0, 0.5, 1.5)
UNIT = "images/s"

# Algorithmic work: matmul+conv FLOPs of the reference, both students, flip twins included.
GFLOP_PER_IMAGE_CAM = 2436.1      # 2 * 2 * (157.04 + 35.14 + 416.85)
GFLOP_PER_IMAGE_TRAIN = 3433.6    # + 2 * 166.25 * 3 (fwd + dgrad + wgrad)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def make_inputs(rank, K):
    from helpers import synth_boxes, synth_cls_labels, synth_images
    x = synth_images(BATCH, SIZE, SIZE, seed=rank)
    cls = synth_cls_labels(BATCH, K, seed=rank)
    box = synth_boxes(BATCH, SIZE, SIZE, seed=rank)
    thr = torch.full((BATCH,), 0.65)
    return x, cls, box, thr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            sm = sorted(float(r[0]) for r in rows if len(r) >= 7)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = max(float(r[1]) for r in rows if len(r) >= 7)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].strip() == "Active" for r in rows if len(r) >= 7)]
                out["power_w_max"] = max(float(r[2]) for r in rows if len(r) >= 7)
                out["samples"] = len(sm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


# ======================================================================================= CPU arms / reference arms
def cpu_cam_par(P, x, cls, box, thr, n_images, students=(1, 2)):
    """cam_par on the host cores through the CPU oracle (reference restatement), on `n_images` images.  Seconds."""
    from oracle import dupl_oracle as O
    x, cls, box, thr = x[:n_images], cls[:n_images], box[:n_images], thr[:n_images]
    h, w = x.shape[-2:]
    t0 = time.perf_counter()
    with torch.no_grad():
        den = O.denormalize_img2(x)
        thr_map = thr.reshape(-1, 1, 1, 1).expand(-1, 1, h, w).contiguous()
        for br in students:
            cam, _aux = O.multi_scale_cam(P, br, x, SCALES)
            O.refine_cams(den, cam * cls[:, :, None, None], cls, thr_map, 0.25, 255, box)
    return time.perf_counter() - t0


def reference_available():
    try:
        from baseline import compat
        return compat.available()
    except Exception:
        return False


class CpuTrainArm:
    """The phase-B step on the host cores, ONE 448x448 image per step (b = 1; both students, MS-CAM + PAR labels + losses +
    backward + AdamW): the unmodified reference's own modules from baseline/_ref (kind "reference") when that tree is
    present, else the oracle port (kind "port": losses + backward, no optimizer)."""

    def __init__(self, P, K):
        self.K = K
        self.kind = "reference" if (reference_available() and K == 20) else "port"
        if self.kind == "reference":
            from baseline.ref_step import ReferenceStep
            self.step = ReferenceStep(torch.device("cpu"), state_dict=P, samples_per_gpu=1)
            self.sample = ("1 image per step (b=1, 448x448) through the unmodified reference loop body on CPU: MS-CAM x2 students, PAR "
                           "refine, fwd/bwd of both students, all losses, AdamW (baseline/ref_step.py over baseline/_ref)")
        else:
            self.P = {k: v.clone().requires_grad_("pos_embed" not in k and ".head." not in k) for k, v in P.items()}
            self.sample = ("1 image per step (b=1, 448x448) through the oracle port (oracle/dupl_oracle.py train_losses + backward; "
                           "no optimizer step)")

    def __call__(self, x, cls, box, n_iter):
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.step(x[:1], cls[:1], box[:1], n_iter)
        else:
            from oracle import dupl_oracle as O
            for v in self.P.values():
                v.grad = None
            cfg = O.VOC_CFG if self.K == 20 else O.COCO_CFG
            loss, _, _ = O.train_losses(self.P, x[:1], cls[:1], box[:1], n_iter, cfg)
            loss.backward()
        return time.perf_counter() - t0


def measure_reference_gpu(dev, P, x, cls, box, n_iter, steps, warmup, local_rank=0, ddp=False, strong_aug=False):
    """The unmodified reference on the GPU (stock PyTorch, fp32, cudnn.benchmark=False / deterministic=True as
    train_final_voc.py:95-102), same batch, same loop body, CUDA-event timed.  -> dict or None when baseline/_ref is absent."""
    if not reference_available():
        return None
    from baseline.ref_step import ReferenceStep
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    step = ReferenceStep(dev, state_dict=P, samples_per_gpu=x.shape[0], ddp=ddp, local_rank=local_rank, strong_aug=strong_aug)
    it = n_iter
    for _ in range(warmup):
        step(x, cls, box, it)
        it += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss, parts = step(x, cls, box, it)
        it += 1
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) * 1000.0 / steps
    ms = e0.elapsed_time(e1) / steps
    out = {"value": x.shape[0] / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "wall_ms_per_step": wall, "steps": steps,
           "warmup": warmup, "loss": float(loss), "torch": torch.__version__,
           "what": "UNMODIFIED reference modules (baseline/_ref: model_dupl, cam_helper, PAR, losses, optimizer) on stock PyTorch, "
                   "loop body of train_final_voc.py:186-472 incl. its host syncs, AdamW; fp32 matmul "
                   f"(allow_tf32 matmul={tf32[0]}, cudnn={tf32[1]}: torch defaults, as the script leaves them), cudnn.deterministic; "
                   + ("with" if strong_aug else "without") + " the PIL RandAugment round trip (:191)",
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    del step
    torch.cuda.empty_cache()
    return out


def cpu_crf(img_u8, prob, iters=10):
    from oracle.densecrf_ref import DenseCRF as RefCRF
    t0 = time.perf_counter()
    RefCRF(iters, 1, 1, 4, 121, 5)(img_u8, prob)
    return time.perf_counter() - t0


def synth_coco_image(seed, H=480, W=640):
    import numpy as np
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (H, W, 3)).astype("uint8")
    t = torch.from_numpy(img).permute(2, 0, 1).float()[None]
    return torch.nn.functional.avg_pool2d(t, 5, 1, 2, count_include_pad=False)[0].permute(1, 2, 0).round().to(torch.uint8).contiguous()


def normalise_u8(img_u8):
    from helpers import MEAN, STD
    x = img_u8.permute(2, 0, 1).float()[None]
    for c in range(3):
        x[:, c] = (x[:, c] - MEAN[c]) / STD[c]
    return x


def train_workload(dataset, K, phase):
    return f"{dataset}{K + 1}_dual_student_phase{phase}_step_448_bs4"


def train_config(dataset, K, phase, world):
    """`config` of the training-step line — identical for this repo's arm and the reference arm (the driver compares them);
    what is specific to how THIS implementation runs the workload goes into the line's `notes` object."""
    return {"workload": train_workload(dataset, K, phase), "per_gpu_batch": BATCH, "image": SIZE, "classes": K + 1,
            "parallelism": f"dp{world}" if world > 1 else "single GPU",
            "l2_policy": "per-step working set of several GB >> 126 MB L2; no explicit flush"}


def train_n_iter(K, phase_c):
    from dupl_b200.train_step import Args
    targs = Args if K == 20 else Args.coco()
    if phase_c:
        return targs.gmm_iters + 1000
    return 20000 if K == 80 else (targs.cam_iters + targs.gmm_iters) // 2


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (one process, all threads;
    rank 0 alone under torchrun).  /root/reference does not exist on the GPU box: the unmodified tree travels as
    baseline/_ref (baseline/install_ref.sh); without it the CPU restatement in oracle/ is timed instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from helpers import init_state_dict
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)                       # torchrun exports OMP_NUM_THREADS=1
    K = 80 if (args.dataset == "coco" or args.workload == "crf_sweep") else 20
    cores = torch.get_num_threads()
    times = []
    extra = {}
    kind = "port"
    if args.workload == "cam_par":
        P = init_state_dict(K + 1)
        x, cls, box, thr = make_inputs(0, K)
        sample = "1 of 4 images per step; MS-CAM + PAR refine for both students through the oracle port"
        for i in range(args.warmup + args.steps):
            t = cpu_cam_par(P, x, cls, box, thr, 1)
            if i >= args.warmup:
                times.append(t)
        metric, config, imgs = "cam_par_refine_images_per_sec", {"workload": "voc21_dual_student_cam_par_refine_448_bs4"}, 1.0
    elif args.workload == "train":
        P = init_state_dict(K + 1)
        x, cls, box, thr = make_inputs(0, K)
        if K == 80:
            cls = cls.to(torch.uint8)
        arm = CpuTrainArm(P, K)
        kind, sample = arm.kind, arm.sample
        n_iter = train_n_iter(K, False)
        xs = torch.nn.functional.interpolate(x, size=(64, 64), mode="bilinear", align_corners=False)
        bs = torch.tensor([[0, 64, 0, 64]] * x.shape[0], dtype=torch.int16)
        for i in range(args.warmup):                  # warm-up = thread pools / allocator only: tiny images, untimed
            arm(xs, cls, bs, n_iter)
        for i in range(args.steps):
            times.append(arm(x, cls, box, n_iter + i))
        metric, imgs = "train_images_per_sec", 1.0
        config = train_config(args.dataset, K, "B", int(os.environ.get("WORLD_SIZE", "1")))
        if torch.cuda.is_available() and K == 20 and not args.no_reference_gpu:
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            torch.cuda.set_device(dev)
            try:
                extra["reference_gpu"] = measure_reference_gpu(dev, P, x, cls, box, n_iter, steps=min(args.steps, 10), warmup=3)
            except Exception as exc:
                extra["reference_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    else:
        img = synth_coco_image(0)
        g = torch.Generator().manual_seed(0)
        lg = torch.randn(1, 81, 28, 28, generator=g) * 2.0
        prob = torch.softmax(torch.nn.functional.interpolate(lg, size=img.shape[:2], mode="bilinear", align_corners=False), 1)[0].numpy()
        sample = "DenseCRF stage only (C restatement, 1 thread) of 1 image 640x480x81; the dense passes of the sweep are not included"
        cores = 1
        for i in range(args.warmup + args.steps):
            t = cpu_crf(img.numpy(), prob)
            if i >= args.warmup:
                times.append(t)
        metric, config, imgs = "crf_sweep_images_per_sec", {"workload": "coco81_mscam_mscseg_densecrf_640x480"}, 1.0
    ms = 1000.0 * sum(times) / len(times)
    val = imgs / (ms / 1000.0)
    line = {"impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line))


# ======================================================================================= GPU arm
class Harness:
    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: dupl_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def time_device(self, fn, steps):
        """CUDA events on the launching stream around `steps` calls, barrier + synchronize on both sides."""
        self.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(self.stream)
        for _ in range(steps):
            fn()
        t1.record(self.stream)
        self.barrier()
        return t0.elapsed_time(t1)

    def time_wall(self, fn, steps):
        self.barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return (time.perf_counter() - w0) * 1000.0

    def max_over_ranks(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def finish(self):
        if self.dist is not None:
            import gc
            gc.collect()                  # captured graphs that still reference NCCL work must be gone first
            torch.cuda.synchronize()
            self.dist.destroy_process_group()


class OpTimer:
    """Per-launch CUDA-event timing of selected ops wrappers (dupl_b200.ops.*) on the launching stream: the tcgen05 GEMM
    (dominant kernel, tensor roofline) and the HBM-bound PAR / MS-CAM kernels."""
    NAMES = ("gemm_bf16x3", "attention_fwd", "attention_bwd", "par_propagate", "par_affinity", "mscam_post")

    def __init__(self, stream):
        from dupl_b200 import ops
        self.ops, self.stream = ops, stream
        self.orig = {n: getattr(ops, n) for n in self.NAMES}
        self.events = {n: [] for n in self.NAMES}
        self.meta = {n: [] for n in self.NAMES}

    def _wrap(self, name):
        orig, events, meta, stream = self.orig[name], self.events[name], self.meta[name], self.stream

        def timed(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = orig(*a, **kw)
            e1.record(stream)
            events.append((e0, e1))
            if name == "gemm_bf16x3":
                groups, M, N, K = a[0], a[1], a[2], a[3]
                meta.append(2.0 * M * N * K * len(groups))
            elif name == "par_propagate":
                meta.append((tuple(a[1].shape), kw.get("nactive", a[4] if len(a) > 4 else None), a[3]))
            elif name == "mscam_post":
                meta.append(float(r.numel() * 4))
            else:
                meta.append(None)
            return r
        return timed

    def __enter__(self):
        for n in self.NAMES:   # every module of the package calls these through the `ops` module attribute
            setattr(self.ops, n, self._wrap(n))
        return self

    def __exit__(self, *exc):
        for n in self.NAMES:
            setattr(self.ops, n, self.orig[n])

    def ms(self, name):
        return sum(a.elapsed_time(b) for a, b in self.events[name])

    def summary(self):
        return self.ms("gemm_bf16x3"), sum(self.meta["gemm_bf16x3"]), len(self.events["gemm_bf16x3"])

    def hbm_kernels(self, peaks):
        """Effective-model HBM fractions of the PAR propagation and MS-CAM post-processing launches:
        PAR per call = T * sum_i 4*h*w*(48 + 2*P_i) bytes (affinity once per image and iteration, P_i live mask planes
        read + written); MS-CAM post = bytes of the [b,K,H,W] fp32 tensor it writes."""
        out = {}
        pk = peaks["hbm_gbs"]
        n = len(self.events["par_propagate"])
        if n:
            total = 0.0
            for (shape, nactive, iters) in self.meta["par_propagate"]:
                B, P, hh, ww = shape
                live = nactive.tolist() if nactive is not None else [P] * B
                total += iters * sum(4.0 * hh * ww * (48 + 2 * p) for p in live)
            ms = self.ms("par_propagate")
            out["par_propagate"] = {"calls": n, "ms_per_call": ms / n, "effective_gbs": total / (ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": total / (ms * 1e-3) / 1e9 / pk,
                                    "model": "T x sum_i 4hw(48 + 2 P_i) B per call (affinity streamed once per image and iteration)"}
        n = len(self.events["par_affinity"])
        if n:
            out["par_affinity"] = {"calls": n, "ms_per_call": self.ms("par_affinity") / n}
        n = len(self.events["mscam_post"])
        if n:
            ms, total = self.ms("mscam_post"), sum(self.meta["mscam_post"])
            out["mscam_post"] = {"calls": n, "ms_per_call": ms / n, "effective_gbs": total / (ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": total / (ms * 1e-3) / 1e9 / pk,
                                 "model": "bytes of the normalised [b,K,448,448] fp32 CAM tensor written once (both passes timed)"}
        for k in ("attention_fwd", "attention_bwd"):
            n = len(self.events[k])
            if n:
                out[k] = {"calls": n, "ms_per_call": self.ms(k) / n}
        return out


def gemm_roofline(timer, peaks, peak_kind, step_tflops, timed_in):
    gemm_ms, gemm_flop, n_gemm = timer.summary()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "gemm_bf16x3_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf if peak_tf else None,
            # launches of many shapes are timed here (CAM half M = 21 976 x 2 students, training half M = 3140): there is no
            # single per-launch DRAM figure for the mix; the ncu --set full captures per shape are in profiles/ (r02_summary.md)
            "traffic": None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_kind})",
            "issued_tflops": 3.0 * achieved, "frac_issued": 3.0 * achieved / peak_tf if peak_tf else None,
            "launches_timed": n_gemm, "avg_launch_ms": gemm_ms / max(n_gemm, 1), "timed_in": timed_in,
            "note": "achieved = algorithmic fp32-GEMM FLOPs (2MNK) / CUDA-event time; the kernel issues 3 bf16 MMAs per "
                    "product (split operands: the 1e-3 fp32 parity bar, profiles/r02_precision_table.md), so the "
                    "tensor pipe does 3x this figure (issued_tflops / frac_issued) and frac is capped at 1/3",
            "step_tflops": step_tflops}


def build_model(K, dev, train=False):
    from helpers import init_state_dict
    from dupl_b200.model.model_dupl import siamese_network
    P = init_state_dict(K + 1)
    model = siamese_network("deit_base_patch16_224", num_classes=K + 1, pretrained=False, aux_layer=-3 if K == 20 else 9)
    model.load_state_dict(P, strict=True)
    model = model.to(dev)
    return (model.train() if train else model.eval()), P


def params_checksum(model):
    """One int64 per parameter tensor: the wrap-around sum of its fp32 bit patterns (bit-level, order-independent)."""
    return torch.stack([p.detach().contiguous().view(torch.int32).to(torch.int64).sum() for p in model.parameters()])


def measure_train(h, args, K, steps, warmup, want_roofline):
    """Full training step.  Returns dict(ms, e2e_ms, launches, timer, h2d, d2h, loss, ...)."""
    from dupl_b200 import _lib as L
    from dupl_b200.train_step import PhaseBStep, make_optimizer, Args
    model, P = build_model(K, h.dev, train=True)
    # one process per model: the whole iteration is one CUDA graph; with several ranks the captured step averages the
    # gradients itself (NCCL inside the graph).  DUPL_TRAIN_CAPTURE=0 or DUPL_TRAIN_CAPTURE_DDP=0 fall back to the eager
    # autograd half under DistributedDataParallel (the reducer's hooks need the eager autograd pass)
    capture = os.environ.get("DUPL_TRAIN_CAPTURE", "1") != "0" and (h.world == 1 or os.environ.get("DUPL_TRAIN_CAPTURE_DDP", "1") == "1")
    optim = make_optimizer(model, capturable=capture)
    wrapped = model
    if h.world > 1 and not capture:
        # train_final_voc.py:155 wraps with find_unused_parameters=True (encoder.head.* never receives a gradient).  The
        # fused driver knows the graph is the same every step: static_graph lets the reducer skip the per-step traversal,
        # bucket views avoid the copy back.  DUPL_DDP=reference restores the script's exact construction for A/B runs.
        if os.environ.get("DUPL_DDP", "") == "reference":
            kw = dict(find_unused_parameters=True)
        else:
            kw = dict(find_unused_parameters=True, static_graph=True, gradient_as_bucket_view=True,
                      bucket_cap_mb=int(os.environ.get("DUPL_DDP_BUCKET_MB", "64")))
        wrapped = torch.nn.parallel.DistributedDataParallel(model, device_ids=[h.local_rank], **kw)
    targs = Args if K == 20 else Args.coco()
    step = PhaseBStep(wrapped, optim, args=targs, device=h.dev, capture=capture)
    x, cls, box, _ = make_inputs(h.rank, K)
    if K == 80:
        cls = cls.to(torch.uint8)           # COCO labels are uint8 (datasets/coco.py)
    phase_c = getattr(args, "phase", "B") == "C"
    # phase B: cam_iters <= n_iter < gmm_iters; phase C adds the strongly augmented view (RandAugment on the device, drawn per step
    # like the script does), the GMM noise filter and the consistency term
    it = [train_n_iter(K, phase_c)]
    from helpers import synth_images
    x_aug = None          # phase C: TrainStep augments `inputs` itself (utils.imutils.augment_data_strong on the GPU)
    x_pin, cls_pin = x.pin_memory(), cls.pin_memory()
    x_dev, cls_dev = x.to(h.dev), cls.to(h.dev)
    aug_pin = aug_dev = None
    last = {}

    def device_step():
        last["loss"], _ = step(x_dev, cls_dev, box, it[0], aug_dev)
        it[0] += 1

    def e2e_step():
        xi = x_pin.to(h.dev, non_blocking=True)
        ci = cls_pin.to(h.dev, non_blocking=True)
        ai = None
        loss, _ = step(xi, ci, box, it[0], ai)
        it[0] += 1
        last["loss_host"] = loss.item()   # device -> host read of the step's result

    for _ in range(warmup):
        device_step()
    sampler = ClockSampler(h.local_rank)
    if h.rank == 0:
        sampler.start()
    ms = h.time_device(device_step, steps)
    out = dict(ms=ms / steps, h2d=int(x_pin.numel() * 4 + cls_pin.numel() * cls_pin.element_size()), d2h=4, P=P,
               inputs=(x, cls, box), n_iter=it[0])
    e2e_step()
    out["e2e_ms"] = h.time_wall(e2e_step, steps) / steps
    out["clocks"] = sampler.stop() if h.rank == 0 else None
    timer = None
    if want_roofline:
        eager = PhaseBStep(wrapped, None, args=targs, device=h.dev, graph=False)

        def eager_iter():   # losses + backward launched kernel by kernel; no optimizer step (the parameters stay as they are)
            model.zero_grad(set_to_none=True)
            loss, _, _ = eager.losses(x_dev, cls_dev, box, it[0], aug_dev)
            loss.backward()

        eager_iter()
        launches0 = L.lib().dupl_launch_count()
        with OpTimer(h.stream) as timer:
            h.barrier()
            for _ in range(steps):
                eager_iter()
            h.barrier()
        out["launches"] = int(L.lib().dupl_launch_count() - launches0)
        del eager
    out["timer"] = timer
    out["capture"] = capture
    # replicas must hold bit-identical parameters after the run (same updates on every rank): EVERY parameter tensor
    chk = params_checksum(model)
    lo, hi = chk.clone(), chk.clone()
    if h.dist is not None:
        h.dist.all_reduce(lo, op=h.dist.ReduceOp.MIN)
        h.dist.all_reduce(hi, op=h.dist.ReduceOp.MAX)
    out["ranks_in_sync"] = bool((lo == hi).all().item())
    out["params_checked"] = int(chk.numel())
    out["loss"] = float(last["loss"].item())
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    model.zero_grad(set_to_none=True)
    del step, optim, wrapped, model
    torch.cuda.empty_cache()
    return out


def measure_cam_par(h, args, steps, warmup, want_breakdown=False):
    """BASELINE.json configs[1]: MS-CAM (3 scales + flips) of both students + PAR refine of both students, b = 4."""
    from dupl_b200 import _lib as L, ops
    from dupl_b200.pipeline import CamParStep
    K = 20
    model, P = build_model(K, h.dev)
    step = CamParStep(model, SCALES, fuse_students=args.fuse_students, graph=not args.no_graph)
    eager = CamParStep(model, SCALES, fuse_students=args.fuse_students)
    step.par.to(h.dev)
    x, cls, box, thr = make_inputs(h.rank, K)
    x_pin, cls_pin, thr_pin = x.pin_memory(), cls.pin_memory(), thr.pin_memory()
    x_dev, cls_dev, thr_dev = x.to(h.dev), cls.to(h.dev), thr.to(h.dev)
    out_pin = torch.empty(2, BATCH, SIZE, SIZE, dtype=torch.float32).pin_memory()

    def device_step():
        return step(x_dev, cls_dev, box, thr_dev)

    def e2e_step():
        xi = x_pin.to(h.dev, non_blocking=True)
        ci = cls_pin.to(h.dev, non_blocking=True)
        ti = thr_pin.to(h.dev, non_blocking=True)
        l1, l2, _, _ = step(xi, ci, box, ti)
        out_pin[0].copy_(l1, non_blocking=True)
        out_pin[1].copy_(l2, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(warmup):
        device_step()
    h.barrier()
    sampler = ClockSampler(h.local_rank)
    if h.rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM (the step is replayed as one CUDA graph unless --no-graph)
    ms_total = h.time_device(device_step, steps)
    # ---- same K steps launched kernel by kernel with CUDA events around the GEMM / PAR / MS-CAM launches and the
    #      library's launch counter (kernels per step)
    eager(x_dev, cls_dev, box, thr_dev)
    launches0 = L.lib().dupl_launch_count()
    with OpTimer(h.stream) as timer:
        h.barrier()
        for _ in range(steps):
            eager(x_dev, cls_dev, box, thr_dev)
        h.barrier()
    launches = L.lib().dupl_launch_count() - launches0
    # ---- timed region 2: end to end through the public API with host buffers
    e2e_step()
    e2e_ms_total = h.time_wall(e2e_step, steps)
    clocks = sampler.stop() if h.rank == 0 else None

    breakdown = None
    if want_breakdown and h.rank == 0:
        recs = []
        names = ["gemm_bf16x3", "attention_fwd", "layernorm_split", "patchify", "cls_rows", "cam_contract", "mscam_post",
                 "par_affinity", "par_propagate", "refine_prologue", "refine_epilogue", "split_bf16", "pos_embed_resize"]
        saved = {n: getattr(ops, n) for n in names}

        def wrap(n, fn):
            def f(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(h.stream)
                r = fn(*a, **k)
                e1.record(h.stream)
                recs.append((n, e0, e1))
                return r
            return f
        for n in names:
            setattr(ops, n, wrap(n, saved[n]))
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(h.stream)
        eager(x_dev, cls_dev, box, thr_dev)
        w1.record(h.stream)
        torch.cuda.synchronize()
        for n in names:
            setattr(ops, n, saved[n])
        breakdown = {"step_ms": w0.elapsed_time(w1)}
        for n, e0, e1 in recs:
            breakdown[n] = breakdown.get(n, 0.0) + e0.elapsed_time(e1)
        breakdown = {k: round(v, 3) for k, v in breakdown.items()}
    ms_step, e2e_ms_step = [v / steps for v in h.max_over_ranks(ms_total, e2e_ms_total)]
    res = dict(ms=ms_step, e2e_ms=e2e_ms_step, launches=int(launches), clocks=clocks, timer=timer, breakdown=breakdown, P=P,
               inputs=(x, cls, box, thr), h2d=int(x_pin.numel() * 4 + cls_pin.numel() * 4 + thr_pin.numel() * 4),
               d2h=int(out_pin.numel() * 4))
    del step, eager, model
    torch.cuda.empty_cache()
    return res


def cam_par_config(h, args):
    return {"workload": "voc21_dual_student_cam_par_refine_448_bs4", "per_gpu_batch": BATCH, "image": SIZE,
            "classes": 21, "cam_scales": list(SCALES), "par_iters": 10,
            "parallelism": f"dp{h.world} (independent batches, no collective on this path)",
            "l2_policy": "per-step working set ~1.5 GB per student >> 126 MB L2; no explicit flush",
            "fuse_students": bool(args.fuse_students), "cuda_graph": not args.no_graph}


def run_cam_par(h, args):
    c = measure_cam_par(h, args, args.steps, args.warmup, want_breakdown=args.breakdown)
    imgs = BATCH * h.world
    if h.rank == 0:
        peaks, peak_kind = load_peaks()
        line = {
            "metric": "cam_par_refine_images_per_sec", "value": imgs / (c["ms"] / 1000.0), "unit": UNIT, "n_gpus": h.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": c["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)", "data": "synthetic",
            "config": cam_par_config(h, args),
            "e2e": {"value": imgs / (c["e2e_ms"] / 1000.0), "unit": UNIT, "h2d_bytes_per_step": c["h2d"],
                    "d2h_bytes_per_step": c["d2h"], "ms_per_step": c["e2e_ms"]},
            "gpu_launches": c["launches"], "clocks": c["clocks"],
            "roofline": gemm_roofline(c["timer"], peaks, peak_kind, GFLOP_PER_IMAGE_CAM * 1e9 * BATCH / (c["ms"] * 1e-3) / 1e12,
                                      "the same K steps launched eagerly right after the (graph-replayed) timed region"),
            "hbm_kernels": c["timer"].hbm_kernels(peaks),
        }
        if c["breakdown"] is not None:
            line["breakdown_ms"] = c["breakdown"]
        if h.world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            x, cls, box, thr = c["inputs"]
            secs = cpu_cam_par(c["P"], x, cls, box, thr, 1)
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "1 of 4 images; MS-CAM + PAR refine for both students through the oracle port, timed once"}
        print(json.dumps(line))


def measure_crf(h, steps):
    """DenseCRF mean-field at COCO size (640x480x81, T = 10, tools/eval_seg_coco_ddp.py:156-188 parameters): device-resident
    ms/img, the same through the reference's numpy interface (utils/dcrf.py:42-69: numpy in, numpy out), and the 1-thread C
    restatement of pydensecrf on the host."""
    from dupl_b200 import ops
    from dupl_b200.utils.dcrf import DenseCRF
    H, W, Cn = 480, 640, 81
    imgs = [synth_coco_image(i) for i in range(4)]
    g = torch.Generator().manual_seed(0)
    probs = []
    for _ in imgs:
        lg = torch.randn(1, Cn, 28, 28, generator=g) * 2.0
        probs.append(torch.softmax(torch.nn.functional.interpolate(lg, size=(H, W), mode="bilinear", align_corners=False), 1)[0].contiguous())
    crf = DenseCRF(10, 1, 1, 4, 121, 5)
    dev_in = [(a.to(h.dev), p.to(h.dev)) for a, p in zip(imgs, probs)]
    np_in = [(a.numpy(), p.numpy()) for a, p in zip(imgs, probs)]
    k = [0]

    def device_step():
        a, p = dev_in[k[0] % len(dev_in)]
        k[0] += 1
        return crf(a, p)

    def numpy_step():
        a, p = np_in[k[0] % len(np_in)]
        k[0] += 1
        return crf(a, p)

    for _ in range(3):
        device_step()
    ms = h.time_device(device_step, steps) / steps
    numpy_step()
    e2e = h.time_wall(numpy_step, steps) / steps
    m2, m5 = ops.last_crf_lattice_sizes()
    N, T = H * W, 10
    per_iter = 0
    for d, M in ((2, m2), (5, m5)):   # splat + (d+1) blurs + slice
        per_iter += (N * Cn * 4 + N * (d + 1) * 8 + M * Cn * 4) + (d + 1) * (2 * M * Cn * 4 + M * 8) + (N * (d + 1) * 8 + M * Cn * 4 + N * Cn * 4)
    per_iter += 2 * N * Cn * 4
    peaks, _ = load_peaks()
    out = {"ms_per_image": ms, "e2e_numpy_ms_per_image": e2e, "size": "640x480x81, T=10, (pos_w 1, sxy 1), (bi_w 4, sxy 121, srgb 5)",
           "lattice_vertices": {"gaussian_d2": m2, "bilateral_d5": m5},
           "effective_gbs": per_iter * T / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": per_iter * T / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
           "e2e_bytes": {"h2d": int(H * W * 3 + Cn * H * W * 4), "d2h": int(Cn * H * W * 4)}}
    return out, np_in[0]


def run_train(h, args):
    K = 80 if args.dataset == "coco" else 20
    t = measure_train(h, args, K, args.steps, args.warmup, want_roofline=not args.no_roofline)
    ms_step, e2e_ms = h.max_over_ranks(t["ms"], t["e2e_ms"])
    imgs = BATCH * h.world
    peaks, peak_kind = load_peaks()
    secondary = {}
    if not args.no_secondary:
        # ---- CAM + PAR half alone (configs[1]) — every rank takes part (max over ranks inside)
        try:
            c = measure_cam_par(h, args, max(5, args.steps // 2), 3)
            secondary["cam_par"] = {"metric": "cam_par_refine_images_per_sec", "value": imgs / (c["ms"] / 1000.0), "unit": UNIT,
                                    "ms_per_step": c["ms"], "ms_per_image": c["ms"] / BATCH,
                                    "e2e": {"value": imgs / (c["e2e_ms"] / 1000.0), "unit": UNIT, "ms_per_step": c["e2e_ms"]},
                                    "workload": "voc21_dual_student_cam_par_refine_448_bs4 (BASELINE.json configs[1])",
                                    "hbm_kernels": c["timer"].hbm_kernels(peaks)}
        except Exception as exc:      # a secondary figure must never take the headline line down with it
            secondary["cam_par"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if h.world == 1:
            try:
                crf, crf_np = measure_crf(h, max(5, args.steps // 2))
                if not args.no_cpu_baseline:
                    secs = cpu_crf(*crf_np)
                    crf["cpu_1_thread_ms_per_image"] = secs * 1000.0
                    crf["speedup_vs_cpu_1_thread"] = secs * 1000.0 / crf["ms_per_image"]
                    crf["cpu_kind"] = "C restatement of the permutohedral mean-field (oracle/densecrf_ref.c), not pydensecrf"
                secondary["crf"] = crf
            except Exception as exc:
                secondary["crf"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            if K == 20 and not args.no_reference_gpu and args.phase == "B":
                try:
                    x, cls, box = t["inputs"]
                    r = measure_reference_gpu(h.dev, t["P"], x, cls, box, train_n_iter(K, False), steps=min(args.steps, 10), warmup=3)
                    if r is not None:
                        r["speedup"] = (imgs / (ms_step / 1000.0)) / r["value"]
                        r["speedup_e2e"] = (imgs / (e2e_ms / 1000.0)) / r["value"]
                        r["target"] = ">= 5x (BASELINE.json north_star)"
                    else:
                        r = {"unavailable": "baseline/_ref absent (run baseline/install_ref.sh where /root/reference is mounted)"}
                    secondary["reference_gpu"] = r
                except Exception as exc:
                    secondary["reference_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if h.rank == 0:
        line = {
            "metric": "train_images_per_sec", "value": imgs / (ms_step / 1000.0), "unit": UNIT, "n_gpus": h.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)", "data": "synthetic",
            "config": train_config(args.dataset, K, args.phase, h.world),
            "notes": {"n_iter": t["n_iter"],
                      "gradient_average": ("none (single GPU)" if h.world == 1 else "NCCL all-reduce of the gradient arenas inside the captured graph"
                                           if t.get("capture") else "DistributedDataParallel reducer"),
                      "forward_reuse": "training forward starts from the MS-CAM pass activations of the un-flipped scale-1.0 images",
                      "cuda_graph": "whole iteration" if t.get("capture") else "CAM + PAR half only (DDP reducer needs eager autograd)"},
            "e2e": {"value": imgs / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": t["h2d"], "d2h_bytes_per_step": t["d2h"],
                    "ms_per_step": e2e_ms},
            "gpu_launches": t.get("launches"), "clocks": t["clocks"], "loss": t["loss"], "peak_mem_gb": round(t["peak_mem_gb"], 2),
            "ranks_in_sync": t["ranks_in_sync"], "params_checked": t["params_checked"],
            "roofline": None if t["timer"] is None else gemm_roofline(
                t["timer"], peaks, peak_kind, GFLOP_PER_IMAGE_TRAIN * 1e9 * BATCH / (ms_step * 1e-3) / 1e12,
                "the same K steps launched eagerly (no graph, no optimizer step) right after the timed region"),
            "hbm_kernels": None if t["timer"] is None else t["timer"].hbm_kernels(peaks),
        }
        line.update(secondary)
        if h.world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            x, cls, box = t["inputs"]
            arm = CpuTrainArm(t["P"], K)
            secs = arm(x, cls, box, train_n_iter(K, False))
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": arm.kind,
                                    "sample": arm.sample + "; timed once"}
        print(json.dumps(line))


def run_crf_sweep(h, args):
    """configs[4]: COCO eval sweep; `--steps` images per rank (strided shard of a synthetic list), batch 1 like the tools."""
    from dupl_b200 import _lib as L, ops
    from dupl_b200.eval_sweep import SegCrfSweep, shard_indices
    from helpers import synth_cls_labels
    K = 80
    model, _ = build_model(K, h.dev)
    sweep = SegCrfSweep(model, flavour="coco")
    n_total = (args.warmup + args.steps) * h.world
    mine = shard_indices(n_total, h.rank, h.world)
    imgs_u8 = [synth_coco_image(i) for i in mine]
    inputs = [torch.nn.functional.interpolate(normalise_u8(im), size=(SIZE, SIZE), mode="bilinear", align_corners=False) for im in imgs_u8]
    cls = [synth_cls_labels(1, K, seed=i) for i in mine]
    pin = [(a.pin_memory(), b.pin_memory(), c.pin_memory()) for a, b, c in zip(imgs_u8, inputs, cls)]
    dev_in = [(a.to(h.dev), b.to(h.dev), c.to(h.dev)) for a, b, c in zip(imgs_u8, inputs, cls)]
    out_pin = torch.empty(480, 640, dtype=torch.int64).pin_memory()
    k = [0]

    def device_step():
        a, b, c = dev_in[k[0] % len(dev_in)]
        k[0] += 1
        return sweep([a], b, c, branch=1)

    def e2e_step():
        a, b, c = pin[k[0] % len(pin)]
        k[0] += 1
        out = sweep([a.to(h.dev, non_blocking=True)], b.to(h.dev, non_blocking=True), c.to(h.dev, non_blocking=True), branch=1)
        out_pin.copy_(out["crf_pred"][0], non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(h.local_rank)
    if h.rank == 0:
        sampler.start()
    launches0 = L.lib().dupl_launch_count()
    ms_total = h.time_device(device_step, args.steps)
    launches = L.lib().dupl_launch_count() - launches0
    # CRF stage alone (its own events) for the roofline of the mean-field kernels
    a, b, c = dev_in[0]
    seg1, _ = sweep.msc_seg(b)
    crf_ms = h.time_device(lambda: sweep.crf_prob(a, seg1[:1]), args.steps) / args.steps
    e2e_step()
    e2e_total = h.time_wall(e2e_step, args.steps)
    clocks = sampler.stop() if h.rank == 0 else None
    ms_step, e2e_ms, crf_ms = h.max_over_ranks(ms_total / args.steps, e2e_total / args.steps, crf_ms)
    if h.rank == 0:
        peaks, peak_kind = load_peaks()
        N, Cn, T = 480 * 640, 81, 10
        m2, m5 = ops.last_crf_lattice_sizes()
        per_iter = 0
        for d, M in ((2, m2), (5, m5)):   # splat + (d+1) blurs + slice
            per_iter += (N * Cn * 4 + N * (d + 1) * 8 + M * Cn * 4) + (d + 1) * (2 * M * Cn * 4 + M * 8) + (N * (d + 1) * 8 + M * Cn * 4 + N * Cn * 4)
        per_iter += 2 * N * Cn * 4
        achieved = per_iter * T / (crf_ms * 1e-3) / 1e9
        line = {
            "metric": "crf_sweep_images_per_sec", "value": h.world / (ms_step / 1000.0), "unit": UNIT, "n_gpus": h.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (CRF) / bf16x3 (dense)", "data": "synthetic",
            "config": {"workload": "coco81_mscam_mscseg_densecrf_640x480", "images_per_step": 1, "seg_scales": list(sweep.scales),
                       "cam_scales": list(sweep.cam_scales), "crf": "T=10, (pos_w 1, sxy 1), (bi_w 4, sxy 121, srgb 5)",
                       "parallelism": f"dp{h.world} (images strided over ranks, no collective)",
                       "l2_policy": "a different image every step; CRF working set ~0.5 GB >> 126 MB L2"},
            "e2e": {"value": h.world / (e2e_ms / 1000.0), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(480 * 640 * 3 + 3 * SIZE * SIZE * 4 + K * 4), "d2h_bytes_per_step": int(480 * 640 * 8)},
            "gpu_launches": int(launches), "clocks": clocks, "crf_ms_per_image": crf_ms,
            "lattice_vertices": {"gaussian_d2": m2, "bilateral_d5": m5},
            "roofline": {"bound": "hbm", "kernel": "crf_* (splat / blur / slice of the two lattices, 10 iterations)", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                         "note": "achieved = SURVEY §8(d) streaming bytes of one image's mean-field / CUDA-event time of the CRF stage"},
        }
        if h.world == 1 and not args.no_cpu_baseline:
            lg = torch.nn.functional.interpolate(seg1[:1].cpu(), size=(480, 640), mode="bilinear", align_corners=False)
            secs = cpu_crf(imgs_u8[0].numpy(), torch.softmax(lg, 1)[0].numpy())
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "DenseCRF stage only of 1 image (C restatement of the lattice mean-field, not pydensecrf)",
                                    "crf_speedup_vs_1_thread": secs * 1000.0 / crf_ms}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dupl_b200", choices=["dupl_b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "cam_par", "crf_sweep"])
    ap.add_argument("--dataset", default="voc", choices=["voc", "coco"], help="train workload: class count / loss weights")
    ap.add_argument("--phase", default="B", choices=["B", "C"], help="train workload: B = CAM+PAR+seg (default), C = + aug view, GMM filter, consistency")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="train: skip the secondary cam_par / crf / reference_gpu measurements")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the unmodified reference on the GPU")
    ap.add_argument("--no-roofline", action="store_true", help="train: skip the eager per-kernel timing pass (quick experiments)")
    ap.add_argument("--no-fuse-students", dest="fuse_students", action="store_false",
                    help="one encoder pass per student instead of both students per grouped GEMM launch")
    ap.add_argument("--breakdown", action="store_true", help="extra untimed step with CUDA events around every op")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    h = Harness(args)
    {"cam_par": run_cam_par, "train": run_train, "crf_sweep": run_crf_sweep}[args.workload](h, args)
    h.finish()


if __name__ == "__main__":
    main()
