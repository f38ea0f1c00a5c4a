"""Generates tests/golden/rpn_pack.npz by running the REFERENCE's own RPN module (modal/modals.py:369-412, unmodified,
imported from /root/reference) on CPU over three small pyramid levels and concatenating its outputs the way
MaskRCNN.predict does (model.py:553-563).  Stored: the two conv outputs of every level (captured with forward hooks --
the inputs of sln_rpn_pack) and the three concatenated tensors.  Run in the build container only:

    python tests/golden/make_golden_rpn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import setup_reference_imports  # noqa: E402


def main():
    setup_reference_imports()
    import modal.modals as M
    torch.manual_seed(7)
    rpn = M.RPN(3, 1, 16)                      # anchors per location, anchor stride, depth
    rpn.eval()
    caught = {"cls": [], "box": []}
    rpn.conv_class.register_forward_hook(lambda m, i, o: caught["cls"].append(o.detach().clone()))
    rpn.conv_bbox.register_forward_hook(lambda m, i, o: caught["box"].append(o.detach().clone()))
    feats = [torch.randn(2, 16, s, w) * 3 for s, w in ((12, 10), (6, 5), (3, 3))]
    with torch.no_grad():
        layer_outputs = [rpn(p) for p in feats]
        outputs = list(zip(*layer_outputs))
        outputs = [torch.cat(list(o), dim=1) for o in outputs]
    out = {"rpn_class_logits": outputs[0].numpy(), "rpn_class": outputs[1].numpy(), "rpn_bbox": outputs[2].numpy()}
    for l in range(3):
        out["cls_%d" % l] = caught["cls"][l].numpy()
        out["box_%d" % l] = caught["box"][l].numpy()
    np.savez_compressed(os.path.join(HERE, "rpn_pack.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
