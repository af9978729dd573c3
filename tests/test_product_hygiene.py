"""CPU tier: the product never routes through the oracle and fails loudly without CUDA."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "retinal_oct_image_segmentation_via_deep_learning_b200")


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_gpu_tests_and_bench_do_not_read_the_reference_checkout():
    assert "/root/reference" not in open(os.path.join(ROOT, "bench.py")).read()
    # __graft_entry__.build() installs the reference's Metrics/ into baseline/_ref in the build container (the one
    # place that may name the checkout); smoke() runs on the GPU box and must not
    import inspect
    import importlib.util
    spec = importlib.util.spec_from_file_location("graft_entry_check", os.path.join(ROOT, "__graft_entry__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert "/root/reference" not in inspect.getsource(mod.smoke)
    for f in os.listdir(os.path.join(ROOT, "tests")):
        if f.startswith("test_gpu"):
            assert "/root/reference" not in open(os.path.join(ROOT, "tests", f)).read()


def test_dropin_modules_have_the_reference_function_names():
    import importlib.util
    want = {
        "ConfusionMatrix_based_metrics": ["accuracy", "sensitivity", "precision", "specificity", "auc_score"],
        "Region_based_metrics": ["dice_coefficient", "iou_score", "precision", "recall"],
        "Contour_based_metrics": ["hausdorff_distance", "hausdorff_distance_95", "assd", "mad"],
        "PixelError_based_metrics": ["mean_squared_error", "root_mean_squared_error"],
        "Biomarker_based_metrics": ["thickness_difference", "vascularity_index"],
    }
    import inspect
    for mod, names in want.items():
        spec = importlib.util.spec_from_file_location(mod, os.path.join(PKG, "Metrics", mod + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        for n in names:
            assert list(inspect.signature(getattr(m, n)).parameters) == ["y_true", "y_pred"], (mod, n)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import importlib.util
    spec = importlib.util.spec_from_file_location("R", os.path.join(PKG, "Metrics", "Region_based_metrics.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.dice_coefficient(np.zeros((4, 4), np.uint8), np.zeros((4, 4), np.uint8))
    from retinal_oct_image_segmentation_via_deep_learning_b200 import suite
    with pytest.raises(TypeError):
        suite.evaluate(torch.zeros((1, 4, 4), dtype=torch.uint8), torch.zeros((1, 4, 4), dtype=torch.uint8), 2)
