"""Import-path shadow of the reference's nms/pth_nms.py."""
from sln_amodal_b200.nms import pth_nms  # noqa: F401
