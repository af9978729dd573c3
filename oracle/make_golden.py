"""Generate ``tests/golden/*.npz`` by EXECUTING the reference (run in the build container only).

    python -m oracle.make_golden

* ``counts_golden.npz``   -- per-class outputs of the 12 count-derived functions and
  ``thickness_difference`` produced by the unmodified reference modules under
  ``/root/reference/Metrics`` on seeded label maps (inputs stored alongside).  This is what pins
  ``metrics_oracle``.
* ``contours_golden.npz`` -- contour ``[0]`` vertices, squared distances and the three contour
  metrics from the RESTATED ``find_contours`` + the reference's distance expressions
  (scikit-image is absent, so this part is self-generated: "parity unpinned"); it pins the CUDA
  path to the oracle across refactors, and the squared distances are cross-checked against
  ``scipy.ndimage.distance_transform_edt`` in the tests.
* ``auc_golden.npz``      -- ``auc_score`` of the unmodified reference (scikit-learn of this image) on
  smooth, heavily tied and degenerate score maps.
* ``suite_golden.npz``    -- one reduced-size and one full-size B-scan per BASELINE config with all
  integer intermediates (confusion, thickness, boundaries).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import labelmap_oracle as lo            # noqa: E402
from oracle import ref_loader                       # noqa: E402
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF_FUNCS = lo.COUNT_METRICS + ("thickness_difference",)


def count_cases():
    """(name, y_true, y_pred, K) small label-map pairs incl. the edge cases of SURVEY.md 8a."""
    cases = []
    yt, yp = synth.layered_pair(3, 48, 64, 5, seed=11, noise=0.02)
    for i in range(3):
        cases.append((f"layered48x64_{i}", yt[i], yp[i], 5))
    yt, yp = synth.random_pair(2, 31, 37, 8, seed=12)
    for i in range(2):
        cases.append((f"random31x37_{i}", yt[i], yp[i], 8))
    yt, yp = synth.layered_pair(1, 496, 512, 8, seed=13, noise=0.01)
    cases.append(("layered496x512", yt[0], yp[0], 8))
    z = np.zeros((16, 20), np.uint8)
    o = np.ones((16, 20), np.uint8)
    cases.append(("both_empty", z, z, 2))           # class 1 absent everywhere
    cases.append(("true_empty", z, o, 2))
    cases.append(("both_full", o, o, 2))
    yt, yp = synth.lesion_pair(1, 64, 64, 4, seed=14)
    cases.append(("lesion64", yt[0], yp[0], 4))
    return cases


def make_counts(ref):
    blob = {}
    names = []
    for name, yt, yp, k in count_cases():
        names.append(name)
        blob[name + "/y_true"], blob[name + "/y_pred"] = yt, yp
        blob[name + "/K"] = np.int64(k)
        for fn in REF_FUNCS:
            vals = np.empty(k, np.float64)
            for c in range(k):
                vals[c] = getattr(ref, fn)((yt == c).astype(np.int64), (yp == c).astype(np.int64))
            blob[f"{name}/{fn}"] = vals
        # the same functions on bool masks (appendix B: identical to the int64 path)
        blob[name + "/dice_bool"] = np.array(
            [ref.dice_coefficient(yt == c, yp == c) for c in range(k)], np.float64)
    blob["names"] = np.array(names)
    blob["source"] = np.array("executed reference: " + ref.root)
    np.savez_compressed(os.path.join(OUT, "counts_golden.npz"), **blob)
    print("counts_golden.npz:", len(names), "cases")


def contour_cases():
    cases = []
    rng = np.random.default_rng(21)
    for i in range(6):                                   # random blobs, all topologies incl. saddles
        h, w = rng.integers(6, 20, size=2)
        a = (rng.random((h, w)) < 0.5).astype(np.uint8)
        b = (rng.random((h, w)) < 0.5).astype(np.uint8)
        cases.append((f"rand{i}", a, b))
    yt, yp = synth.lesion_pair(2, 96, 96, 4, seed=22)
    cases.append(("lesion_closed", (yt[0] == 1).astype(np.uint8), (yp[0] == 1).astype(np.uint8)))
    yt, yp = synth.lesion_pair(2, 96, 96, 4, seed=23, single_blob_interior=False)
    cases.append(("lesion_multi", (yt[1] == 2).astype(np.uint8), (yp[1] == 2).astype(np.uint8)))
    yt, yp = synth.layered_pair(1, 120, 160, 6, seed=24)
    cases.append(("layer_open", (yt[0] == 2).astype(np.uint8), (yp[0] == 2).astype(np.uint8)))
    yt, yp = synth.layered_pair(1, 496, 512, 8, seed=25)
    cases.append(("layer496x512_c3", (yt[0] == 3).astype(np.uint8), (yp[0] == 3).astype(np.uint8)))
    return cases


def make_contours():
    from oracle import metrics_oracle as mo
    blob, names = {}, []
    for name, a, b in contour_cases():
        names.append(name)
        im = lo.contour_intermediates(a, b)
        blob[name + "/mask_true"], blob[name + "/mask_pred"] = a, b
        for k, v in im.items():
            blob[f"{name}/{k}"] = v
        blob[name + "/metrics"] = np.array(
            [mo.hausdorff_distance(a, b), mo.hausdorff_distance_95(a, b), mo.assd(a, b)], np.float64)
    blob["names"] = np.array(names)
    blob["source"] = np.array("oracle restatement of find_contours (scikit-image absent): parity unpinned")
    np.savez_compressed(os.path.join(OUT, "contours_golden.npz"), **blob)
    print("contours_golden.npz:", len(names), "cases")


def make_suite():
    blob, names = {}, []
    plan = [
        ("cfg1_small", synth.layered_pair(2, 62, 96, 8, seed=1001, noise=0.01), 8),
        ("cfg1_full", synth.layered_pair(1, 496, 768, 8, seed=1001, noise=0.01), 8),
        ("cfg2_small", synth.layered_pair(2, 62, 128, 10, seed=2002), 10),
        ("cfg2_full", synth.layered_pair(1, 496, 1024, 10, seed=2002), 10),
        ("cfg3_small", synth.lesion_pair(2, 64, 64, 4, seed=3003), 4),
        ("cfg4_full", synth.layered_pair(1, 496, 512, 8, seed=4004), 8),
        ("cfg4_random", synth.random_pair(1, 496, 512, 8, seed=4005), 8),
        ("ragged_33x50", synth.random_pair(3, 33, 50, 16, seed=4006), 16),
    ]
    for name, (yt, yp), k in plan:
        names.append(name)
        blob[name + "/y_true"], blob[name + "/y_pred"], blob[name + "/K"] = yt, yp, np.int64(k)
        per = [lo.score_bscan_fast(yt[i], yp[i], k) for i in range(len(yt))]
        for key in per[0]:
            blob[f"{name}/{key}"] = np.stack([p[key] for p in per])
    blob["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "suite_golden.npz"), **blob)
    print("suite_golden.npz:", len(names), "cases")


def auc_cases():
    """(name, y_true mask, score map): smooth probabilities, heavy ties, float32/float64, edge cases."""
    rng = np.random.default_rng(31)
    cases = []
    yt, _ = synth.lesion_pair(1, 64, 64, 4, seed=32)
    m = (yt[0] > 0).astype(np.uint8)
    logit = 3.0 * (m.astype(np.float64) - 0.5) + rng.normal(0, 1.5, m.shape)
    cases.append(("lesion64_f64", m, 1.0 / (1.0 + np.exp(-logit))))
    cases.append(("lesion64_f32", m, (1.0 / (1.0 + np.exp(-logit))).astype(np.float32)))
    cases.append(("ties_quantised", m, np.round(1.0 / (1.0 + np.exp(-logit)), 1)))          # 11 distinct scores
    cases.append(("binary_scores", m, (logit > 0).astype(np.float64)))                       # 2 distinct scores
    cases.append(("constant_scores", m, np.full(m.shape, 0.5)))                              # one tie run -> 0.5
    cases.append(("perfect", m, m.astype(np.float64) * 0.8 + 0.1))
    cases.append(("inverted", m, 0.9 - m.astype(np.float64) * 0.8))
    cases.append(("negative_and_zero", m, np.where(rng.random(m.shape) < 0.3, -0.0, logit)))  # -0.0 == +0.0 ties
    yt2, yp2 = synth.layered_pair(1, 496, 512, 8, seed=33, noise=0.02)
    m2 = (yt2[0] == 3).astype(np.uint8)
    p2 = np.clip((yp2[0] == 3) * 0.7 + rng.random(m2.shape) * 0.3, 0, 1).astype(np.float32)
    cases.append(("layer496x512_f32", m2, p2))
    cases.append(("single_class", np.zeros((8, 8), np.uint8), rng.random((8, 8))))
    cases.append(("three_labels", (rng.integers(0, 3, (8, 8))).astype(np.uint8), rng.random((8, 8))))
    bad = rng.random((8, 8)); bad[2, 3] = np.nan
    cases.append(("nan_score", (rng.random((8, 8)) < 0.5).astype(np.uint8), bad))
    return cases


def make_auc(ref):
    import warnings
    blob, names = {}, []
    for name, m, p in auc_cases():
        names.append(name)
        blob[name + "/y_true"], blob[name + "/scores"] = m, p
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            blob[name + "/auc"] = np.float64(ref.auc_score(m, p))
    import sklearn
    blob["names"] = np.array(names)
    blob["source"] = np.array(f"executed reference: {ref.root} with scikit-learn {sklearn.__version__}")
    np.savez_compressed(os.path.join(OUT, "auc_golden.npz"), **blob)
    print("auc_golden.npz:", len(names), "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    if ref is None:
        raise SystemExit("reference not found: run this in the build container")
    which = sys.argv[1:] or ["counts", "contours", "suite", "auc"]
    if "counts" in which:
        make_counts(ref)
    if "contours" in which:
        make_contours()
    if "suite" in which:
        make_suite()
    if "auc" in which:
        make_auc(ref)


if __name__ == "__main__":
    main()
