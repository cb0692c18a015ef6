import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images
from dupl_b200.model.model_dupl import siamese_network
from dupl_b200.train_step import TrainStep, make_optimizer, Args
m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
m.load_state_dict(init_state_dict(21), strict=True); m = m.cuda().train()
opt = make_optimizer(m)
step = TrainStep(m, opt, args=Args)
x = synth_images(4, 448, 448, seed=0).cuda(); xa = synth_images(4, 448, 448, seed=100).cuda()
cls = synth_cls_labels(4, 20, seed=0).cuda(); box = synth_boxes(4, 448, 448, seed=0)
for i in range(3): step(x, cls, box, 9000 + i, xa)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(x, cls, box, 9005, xa); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
