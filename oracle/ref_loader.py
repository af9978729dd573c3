"""Import the UNMODIFIED reference modules when they are on this machine (TEST ORACLE helper).

Only the build container has ``/root/reference``; the GPU box does not, so nothing under
``-m gpu`` may depend on this.  Four of the five modules import (``Contour_based_metrics`` needs
scikit-image, absent from the image); ``load()`` returns a namespace exposing them under the
names ``metrics_oracle`` uses, or None when the reference is not present.
"""
from __future__ import annotations

import importlib.util
import os
import types

# the installed copy first (it travels to the GPU box; bench.py's reference arm uses it), the checkout for the CPU tests
SEARCH = (os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref", "Metrics"), "/root/reference/Metrics")


def _import(path, name):
    spec = importlib.util.spec_from_file_location("_octref_" + name, os.path.join(path, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    for root in SEARCH:
        if os.path.isfile(os.path.join(root, "Region_based_metrics.py")):
            break
    else:
        return None
    cmm = _import(root, "ConfusionMatrix_based_metrics")
    reg = _import(root, "Region_based_metrics")
    pix = _import(root, "PixelError_based_metrics")
    bio = _import(root, "Biomarker_based_metrics")
    ns = types.SimpleNamespace(
        root=root,
        accuracy=cmm.accuracy, sensitivity=cmm.sensitivity, cm_precision=cmm.precision,
        specificity=cmm.specificity, auc_score=cmm.auc_score,
        dice_coefficient=reg.dice_coefficient, iou_score=reg.iou_score,
        region_precision=reg.precision, recall=reg.recall,
        mean_squared_error=pix.mean_squared_error, root_mean_squared_error=pix.root_mean_squared_error,
        thickness_difference=bio.thickness_difference, vascularity_index=bio.vascularity_index,
    )
    # Contour_based_metrics.py cannot be imported without scikit-image; its `mad` is plain numpy
    # (lines 58-73) and is restated in metrics_oracle.mad.
    from . import metrics_oracle as mo
    ns.mad = mo.mad
    # hausdorff_distance / hausdorff_distance_95 / assd live in the same un-importable module: the reference's
    # expressions (:19-22, :36-39, :53-56) on the restated find_contours
    ns.hausdorff_distance, ns.hausdorff_distance_95, ns.assd = mo.hausdorff_distance, mo.hausdorff_distance_95, mo.assd
    return ns
