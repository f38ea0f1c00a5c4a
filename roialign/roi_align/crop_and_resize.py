"""Import-path shadow of the reference's roialign/roi_align/crop_and_resize.py: the reference's
`from roialign.roi_align.crop_and_resize import CropAndResizeFunction` (modal/modals.py:6,
modal/Functions.py:7) resolves here when this repo precedes it on sys.path."""
from sln_amodal_b200.crop_and_resize import CropAndResize, CropAndResizeFunction  # noqa: F401
