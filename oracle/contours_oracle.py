"""Restatement of ``skimage.measure.find_contours`` for the contour metrics (TEST ORACLE).

PARITY UNPINNED: the reference calls scikit-image (version not pinned by the
reference, not installed in this image) at ``Contour_based_metrics.py:15,16,
33,34,50,51``.  This file restates the published marching-squares algorithm of
``skimage/measure/_find_contours.py`` (``find_contours``, ``_assemble_contours``)
and ``_find_contours_cy.pyx`` (``_get_contour_segments``, ``_get_fraction``) with
the defaults the reference uses: ``fully_connected='low'``,
``positive_orientation='low'``, no mask.  It cannot be machine-checked against
real scikit-image offline; its invariants are tested in
``tests/test_oracle_contours.py``.
"""
from __future__ import annotations

from collections import deque

import numpy as np

# Which two (or four) square edges each marching-squares case joins, as (from, to) edge
# names; orientation keeps the low side on the left.  Saddles 6 and 9 use the
# "fully_connected='low'" pairing (the two high pixels are separated).
_T, _B, _L, _R = "top", "bottom", "left", "right"
_CASE_SEGMENTS = {
    1: ((_T, _L),), 2: ((_R, _T),), 3: ((_R, _L),), 4: ((_L, _B),), 5: ((_T, _B),),
    6: ((_R, _T), (_L, _B)), 7: ((_R, _B),), 8: ((_B, _R),), 9: ((_T, _L), (_B, _R)),
    10: ((_B, _T),), 11: ((_B, _L),), 12: ((_L, _R),), 13: ((_T, _R),), 14: ((_L, _T),),
}


def _fraction(a, b, level):
    """Linear-interpolation position of ``level`` between samples a and b (0 when a == b)."""
    if b == a:
        return 0.0
    return (level - a) / (b - a)


def contour_segments(image, level):
    """All iso-contour segments, in the raster order the squares are visited (r-major).

    Same output as ``contour_segments_literal``; the square classification is vectorised and only
    the mixed squares are visited in Python (upstream does this scan in Cython, so the literal
    double loop would misrepresent its cost in the CPU baseline)."""
    img = np.asarray(image, dtype=np.float64)
    hi = img > level
    case = (hi[:-1, :-1] * 1 + hi[:-1, 1:] * 2 + hi[1:, :-1] * 4 + hi[1:, 1:] * 8).astype(np.uint8)
    nan = np.isnan(img)
    if nan.any():
        bad = nan[:-1, :-1] | nan[:-1, 1:] | nan[1:, :-1] | nan[1:, 1:]
        case[bad] = 0
    rr, cc = np.nonzero((case != 0) & (case != 15))
    out = []
    for r0, c0, k in zip(rr.tolist(), cc.tolist(), case[rr, cc].tolist()):
        r1, c1 = r0 + 1, c0 + 1
        ul, ur, ll, lr = img[r0, c0], img[r0, c1], img[r1, c0], img[r1, c1]
        edge = {
            _T: (float(r0), c0 + _fraction(ul, ur, level)),
            _B: (float(r1), c0 + _fraction(ll, lr, level)),
            _L: (r0 + _fraction(ul, ll, level), float(c0)),
            _R: (r0 + _fraction(ur, lr, level), float(c1)),
        }
        for a, b in _CASE_SEGMENTS[k]:
            out.append((edge[a], edge[b]))
    return out


def contour_segments_literal(image, level):
    """Square-by-square form of the scan (slow; the definition ``contour_segments`` must match).

    Returns a list of ((r, c), (r, c)) float tuples.  Each 2x2 square is classified by which
    corners exceed ``level`` (ul=1, ur=2, ll=4, lr=8)."""
    img = np.asarray(image, dtype=np.float64)
    rows, cols = img.shape
    out = []
    for r0 in range(rows - 1):
        r1 = r0 + 1
        for c0 in range(cols - 1):
            c1 = c0 + 1
            ul, ur, ll, lr = img[r0, c0], img[r0, c1], img[r1, c0], img[r1, c1]
            if np.isnan(ul) or np.isnan(ur) or np.isnan(ll) or np.isnan(lr):
                continue
            case = (1 if ul > level else 0) | (2 if ur > level else 0) \
                | (4 if ll > level else 0) | (8 if lr > level else 0)
            if case == 0 or case == 15:
                continue
            edge = {
                _T: (float(r0), c0 + _fraction(ul, ur, level)),
                _B: (float(r1), c0 + _fraction(ll, lr, level)),
                _L: (r0 + _fraction(ul, ll, level), float(c0)),
                _R: (r0 + _fraction(ur, lr, level), float(c1)),
            }
            for a, b in _CASE_SEGMENTS[case]:
                out.append((edge[a], edge[b]))
    return out


def assemble_contours(segments):
    """Stitch oriented segments into polylines.

    Polylines are keyed by creation order; when two merge the older key survives; a polyline
    whose two ends meet gets its first vertex repeated at the end."""
    next_key = 0
    lines = {}
    by_start = {}
    by_end = {}
    for p_from, p_to in segments:
        if p_from == p_to:
            continue
        after, after_key = by_start.pop(p_to, (None, None))      # polyline beginning at p_to
        before, before_key = by_end.pop(p_from, (None, None))    # polyline ending at p_from
        if after is not None and before is not None:
            if after is before:
                before.append(p_to)                              # loop closed
            elif after_key > before_key:
                before.extend(after)
                lines.pop(after_key, None)
                by_start[before[0]] = (before, before_key)
                by_end[before[-1]] = (before, before_key)
            else:
                after.extendleft(reversed(before))
                by_start.pop(before[0], None)
                lines.pop(before_key, None)
                by_start[after[0]] = (after, after_key)
                by_end[after[-1]] = (after, after_key)
        elif after is None and before is None:
            line = deque((p_from, p_to))
            lines[next_key] = line
            by_start[p_from] = (line, next_key)
            by_end[p_to] = (line, next_key)
            next_key += 1
        elif before is None:
            after.appendleft(p_from)
            by_start[p_from] = (after, after_key)
        else:
            before.append(p_to)
            by_end[p_to] = (before, before_key)
    return [np.array(lines[k]) for k in sorted(lines)]


def find_contours(image, level=0.5):
    """``skimage.measure.find_contours(image, level)`` with default options."""
    img = np.asarray(image)
    if img.ndim != 2:
        raise ValueError("Only 2D arrays are supported.")
    if img.shape[0] < 2 or img.shape[1] < 2:
        raise ValueError("Input array must be at least 2x2.")
    return assemble_contours(contour_segments(img.astype(np.float64), float(level)))


# ------------------------------------------------------------------ integer-lattice helpers
def to_lattice(contour):
    """(n,2) float (row, col) vertices at edge midpoints -> (n,2) int64 on the doubled lattice."""
    c2 = np.asarray(contour) * 2.0
    ci = np.rint(c2).astype(np.int64)
    assert np.array_equal(ci.astype(np.float64), c2), "vertex not on the half-pixel lattice"
    return ci


def first_contour_lattice(mask):
    """Doubled-lattice vertices of ``find_contours(mask, .5)[0]`` (with the closing repeat)."""
    return to_lattice(find_contours(np.asarray(mask), 0.5)[0])


def directed_sq_distances(src, qry):
    """For each vertex of ``qry`` the minimum squared lattice distance to ``src`` (exact int64)."""
    src = np.asarray(src, dtype=np.int64)
    qry = np.asarray(qry, dtype=np.int64)
    out = np.empty(len(qry), dtype=np.int64)
    step = max(1, (1 << 22) // max(1, len(src)))
    for s in range(0, len(qry), step):
        q = qry[s:s + step]
        d = (q[:, None, 0] - src[None, :, 0]) ** 2 + (q[:, None, 1] - src[None, :, 1]) ** 2
        out[s:s + step] = d.min(axis=1)
    return out


def all_crack_vertices(mask):
    """Doubled-lattice midpoints of every crack between 4-adjacent unequal pixels (sorted rows)."""
    m = np.asarray(mask) > 0.5
    r, c = np.nonzero(m[:-1, :] != m[1:, :])
    v = np.stack([2 * r + 1, 2 * c], axis=1)
    r, c = np.nonzero(m[:, :-1] != m[:, 1:])
    h = np.stack([2 * r, 2 * c + 1], axis=1)
    return np.concatenate([v, h], axis=0).astype(np.int64)
