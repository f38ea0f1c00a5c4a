"""Import-path shadow of the reference's nms/nms_wrapper.py (`from nms.nms_wrapper import nms`,
modal/Functions.py:5)."""
from sln_amodal_b200.nms import nms, pth_nms  # noqa: F401
