"""CPU model of distance_search_kernel's pruning: counts box scans per 32-query chunk for a few strategies.
Tuning aid only (uses the oracle's contours); not part of the product path."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import labelmap_oracle as lo
from retinal_oct_image_segmentation_via_deep_learning_b200 import synth

KBOX = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ORDER = sys.argv[2] if len(sys.argv) > 2 else "polyline"


def boxes_of(src, kbox):
    nb = (len(src) + kbox - 1) // kbox
    bx = np.zeros((nb, 4), np.int64)
    for b in range(nb):
        s = src[b * kbox:(b + 1) * kbox]
        bx[b] = (s[:, 0].min(), s[:, 0].max(), s[:, 1].min(), s[:, 1].max())
    return bx


def lb_box_box(bx, ymin, ymax, xmin, xmax):
    dy = np.maximum(np.maximum(bx[:, 0] - ymax, ymin - bx[:, 1]), 0)
    dx = np.maximum(np.maximum(bx[:, 2] - xmax, xmin - bx[:, 3]), 0)
    return dy * dy + dx * dx


def lb_box_pts(b, q):
    dy = np.maximum(np.maximum(b[0] - q[:, 0], q[:, 0] - b[1]), 0)
    dx = np.maximum(np.maximum(b[2] - q[:, 1], q[:, 1] - b[3]), 0)
    return dy * dy + dx * dx


def scan(src, b, q, best, kbox):
    s = src[b * kbox:(b + 1) * kbox]
    d = ((q[:, None, :] - s[None, :, :]) ** 2).sum(-1).min(1)
    return np.minimum(best, d)


def run(src, qry, kbox, strategy):
    bx = boxes_of(src, kbox)
    nb = len(bx)
    ratio = len(src) / len(qry)
    scans = tests = 0
    out = np.zeros(len(qry), np.int64)
    for c in range((len(qry) + 31) // 32):
        q = qry[c * 32:(c + 1) * 32]
        best = np.full(len(q), 1 << 40)
        bg = min(min(len(src) - 1, int((c * 32 + 16) * ratio)) // kbox, nb - 1)
        primed = [bg]
        if strategy == "prime3":
            primed = [b for b in (bg - 1, bg, bg + 1) if 0 <= b < nb]
        if strategy == "prime2":       # 32 queries span about 2 boxes of 16
            b0 = min(min(len(src) - 1, int((c * 32 + 8) * ratio)) // kbox, nb - 1)
            b1 = min(min(len(src) - 1, int((c * 32 + 24) * ratio)) // kbox, nb - 1)
            primed = sorted({b0, b1})
        for b in primed:
            best = scan(src, b, q, best, kbox); scans += 1
        lbq = lb_box_box(bx, q[:, 0].min(), q[:, 0].max(), q[:, 1].min(), q[:, 1].max())
        if strategy == "sorted":
            order = np.argsort(lbq, kind="stable")
            for b in order:
                if b in primed or lbq[b] >= best.max():
                    continue
                tests += 1
                if (lb_box_pts(bx[b], q) < best).any():
                    best = scan(src, b, q, best, kbox); scans += 1
        else:
            for b0 in range(0, nb, 32):
                bmax = best.max()
                cand = [b for b in range(b0, min(nb, b0 + 32)) if b not in primed and lbq[b] < bmax]
                if strategy == "outward":
                    cand.sort(key=lambda b: abs(b - bg))
                for b in cand:
                    tests += 1
                    if (lb_box_pts(bx[b], q) < best).any():
                        best = scan(src, b, q, best, kbox); scans += 1
        out[c * 32:(c + 1) * 32] = best
    return out, scans, tests


yt, yp = synth.layered_pair(4, 496, 512, 8, seed=4004)
tot = {}
nchunks = 0
for i in range(2):
    for c in range(8):
        im = lo.contour_intermediates(yt[i] == c, yp[i] == c)
        if im is None:
            continue
        a, b = im["verts_true"].astype(np.int64), im["verts_pred"].astype(np.int64)
        if ORDER == "gpu":     # the walk starts at the raster-first vertex: forward run, then the backward run
            def gpu_order(p):
                if (p[0] == p[-1]).all():
                    return p
                s0 = np.lexsort((p[:, 1], p[:, 0]))[0]
                return np.concatenate([p[s0:], p[:s0][::-1]])
            a, b = gpu_order(a), gpu_order(b)
        for src, qry, ref in ((a, b, im["sq_pred_to_true"]), (b, a, im["sq_true_to_pred"])):
            nchunks += (len(qry) + 31) // 32
            for st in ("current", "outward", "sorted", "prime3", "prime2"):
                out, scans, tests = run(src, qry, KBOX, st)
                assert (np.sort(out) == np.sort(ref)).all(), st
                t = tot.setdefault(st, [0, 0])
                t[0] += scans; t[1] += tests
print("kbox", KBOX, "chunks", nchunks)
for st, (s, t) in tot.items():
    print(f"{st:8s} scans/chunk {s / nchunks:.2f}  vertex evals/chunk {s * KBOX / nchunks:.0f}  lane tests/chunk {t / nchunks:.2f}")


def run_columns(src, qry, phases):
    """Sources sorted by column; a chunk scans its own column span, then flanks of r = ceil(sqrt(max best)) columns."""
    import math
    order = np.argsort(src[:, 1], kind="stable")
    s = src[order]
    xs = s[:, 1]
    evals = 0
    out = np.zeros(len(qry), np.int64)
    for c in range((len(qry) + 31) // 32):
        q = qry[c * 32:(c + 1) * 32]
        x0, x1 = q[:, 1].min(), q[:, 1].max()
        lo, hi = np.searchsorted(xs, x0, "left"), np.searchsorted(xs, x1, "right")
        best = np.full(len(q), 1 << 40)
        if hi > lo:
            best = np.minimum(best, ((q[:, None, :] - s[None, lo:hi, :]) ** 2).sum(-1).min(1))
        evals += hi - lo
        done_l, done_r = lo, hi
        for ph in range(phases):
            bmax = best.max()
            r = len(xs) * 4 if bmax >= (1 << 40) else int(math.isqrt(int(bmax) - 1)) if bmax > 0 else 0
            if ph < phases - 1:
                r = min(r, 4 << ph)
            nl, nr = np.searchsorted(xs, x0 - r, "left"), np.searchsorted(xs, x1 + r, "right")
            for a, b in ((nl, done_l), (done_r, nr)):
                if b > a:
                    best = np.minimum(best, ((q[:, None, :] - s[None, a:b, :]) ** 2).sum(-1).min(1))
                    evals += b - a
            done_l, done_r = min(nl, done_l), max(nr, done_r)
        out[c * 32:(c + 1) * 32] = best
    return out, evals


tot = {1: 0, 2: 0, 3: 0}
for i in range(2):
    for c in range(8):
        im = lo.contour_intermediates(yt[i] == c, yp[i] == c)
        if im is None:
            continue
        a, b = im["verts_true"].astype(np.int64), im["verts_pred"].astype(np.int64)
        for src, qry, ref in ((a, b, im["sq_pred_to_true"]), (b, a, im["sq_true_to_pred"])):
            for ph in tot:
                out, ev = run_columns(src, qry, ph)
                assert (np.sort(out) == np.sort(ref)).all(), ph
                tot[ph] += ev
for ph, ev in tot.items():
    print(f"columns, {ph} flank phase(s): vertex evals/chunk {ev / nchunks:.0f}")
