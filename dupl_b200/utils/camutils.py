"""The subset of helpers the reference duplicates in utils/camutils.py:5-194 (same semantics as
utils/cam_helper.py; north_star names this module, so both are exported)."""
from .cam_helper import (cam_to_label, get_valid_cam, label_to_aff_mask, multi_scale_cam2,  # noqa: F401
                         refine_cams_with_bkg_v2, _refine_cams)
