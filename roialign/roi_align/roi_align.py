"""Import-path shadow of the reference's roialign/roi_align/roi_align.py."""
from sln_amodal_b200.crop_and_resize import CropAndResize, CropAndResizeFunction, RoIAlign  # noqa: F401
