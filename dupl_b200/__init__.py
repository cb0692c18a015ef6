"""dupl_b200 — B200-native (sm_100a) implementation of the DuPL hot path.

Drop-in module surface (same names / signatures as the reference):
    dupl_b200.model.model_dupl   siamese_network, network
    dupl_b200.model.PAR          PAR
    dupl_b200.model.losses       get_masked_ptc_loss, get_seg_loss, get_seg_loss_conflict_v2
    dupl_b200.utils.cam_helper   multi_scale_cam2_siamese, cam_to_label, refine_cams_with_*, ...
    dupl_b200.utils.camutils     the subset the reference keeps in utils/camutils.py
All arithmetic runs in libdupl.so (hand-written CUDA); see include/dupl.h and INTEGRATION.md.
"""
__version__ = "0.1.0"
