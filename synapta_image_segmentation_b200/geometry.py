"""Box filtering / merging rules of the reference, applied to raster component boxes.

Pure float64 Python arithmetic in the reference's own operation order, so decisions and the
`confidence` floats are bit-identical (0.3+0.2+0.2-0.2-0.3 must come out as
0.19999999999999996).  Every function cites the reference method it mirrors; golden vectors produced
by the imported reference are in tests/golden/reference_geometry.json.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .datamodel import BoundingBox

# constants (points unless noted)
MIN_AREA = 3000            # pdf_image_segmentation.py:2944
GOOD_AREA = 10000          # :2946
MIN_DIM_PX = 50            # :2954
SUBSTANTIAL_DIM_PX = 200   # :2957
KEEP_SCORE = 0.5           # :2885
CONFLICT_OVERLAP = 0.4     # :3025
DUPLICATE_OVERLAP = 0.5    # :3633
CLUSTER_GAP = 100          # :3560
CLUSTER_MIN_MEMBERS = 3    # :3591
REGION_PADDING = 10        # :3535
DRAWING_MIN_AREA = 5000    # :3549
DRAWING_MAX_PAGE_FRACTION = 0.8


def calculate_overlap_ratio(b1: BoundingBox, b2: BoundingBox) -> float:
    """intersection / smaller area (_calculate_overlap_ratio, pdf_image_segmentation.py:3029-3039)."""
    xo = max(0, min(b1.x1, b2.x1) - max(b1.x0, b2.x0))
    yo = max(0, min(b1.y1, b2.y1) - max(b1.y0, b2.y0))
    inter = xo * yo
    smaller = min(b1.area(), b2.area())
    return inter / smaller if smaller > 0 else 0


def find_conflicting(candidate_bbox: BoundingBox, existing: Sequence, key=lambda s: s.bbox):
    """First existing item whose overlap ratio with the candidate exceeds 0.4 (_find_conflicting_segment, :3020-3027)."""
    for item in existing:
        if calculate_overlap_ratio(candidate_bbox, key(item)) > CONFLICT_OVERLAP:
            return item
    return None


def overlaps_with_existing(bbox: BoundingBox, existing_regions: Sequence[Dict]) -> bool:
    """intersection > 0.5 * area(candidate) against any existing region (_overlaps_with_existing, :3620-3636)."""
    for region in existing_regions:
        e = region["bbox"]
        xo = max(0, min(bbox.x1, e.x1) - max(bbox.x0, e.x0))
        yo = max(0, min(bbox.y1, e.y1) - max(bbox.y0, e.y0))
        if xo * yo > bbox.area() * DUPLICATE_OVERLAP:
            return True
    return False


def drawing_distance(r1: Sequence[float], r2: Sequence[float]) -> float:
    """Gap between two rects, 0 when they touch or overlap (_drawing_distance, :3596-3618)."""
    if r1[0] <= r2[2] and r1[2] >= r2[0] and r1[1] <= r2[3] and r1[3] >= r2[1]:
        return 0
    dx = max(0, max(r1[0] - r2[2], r2[0] - r1[2]))
    dy = max(0, max(r1[1] - r2[3], r2[1] - r1[3]))
    return (dx ** 2 + dy ** 2) ** 0.5


def cluster_rects(rects: Sequence[Sequence[float]], distance_threshold: float = CLUSTER_GAP,
                  min_members: int = CLUSTER_MIN_MEMBERS) -> List[List[int]]:
    """Greedy single-pass clustering (_cluster_drawings, :3559-3594): seed i absorbs every unused j closer
    than the threshold to the SEED (no transitive closure); clusters below `min_members` are dropped but
    their members stay consumed.  Returns index lists in the reference's order.

    The O(n) inner loop of a seed is vectorised (numpy f64, the reference's own operation order for the gaps);
    `dx*dx + dy*dy < t*t` decides every pair except those within 1e-9 (relative) of the threshold, which go
    through the scalar `drawing_distance` -- `(dx**2 + dy**2)**0.5 < t` as the reference writes it -- so the
    result is the reference's for every input."""
    n = len(rects)
    if n == 0:
        return []
    if n < 16:
        return _cluster_rects_scalar(rects, distance_threshold, min_members)
    import numpy as np
    r = np.asarray(rects, dtype=np.float64).reshape(n, 4)
    x0, y0, x1, y1 = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    used = np.zeros(n, dtype=bool)
    t2 = float(distance_threshold) * float(distance_threshold)
    clusters: List[List[int]] = []
    for i in range(n):
        if used[i]:
            continue
        used[i] = True
        a0, b0, a1, b1 = x0[i], y0[i], x1[i], y1[i]
        touch = (a0 <= x1) & (a1 >= x0) & (b0 <= y1) & (b1 >= y0)
        dx = np.maximum(0.0, np.maximum(a0 - x1, x0 - a1))
        dy = np.maximum(0.0, np.maximum(b0 - y1, y0 - b1))
        v = dx * dx + dy * dy
        near = (touch & (distance_threshold > 0)) | (~touch & (v < t2))
        doubt = ~touch & (np.abs(v - t2) <= 1e-9 * t2)
        for j in np.nonzero(doubt)[0]:
            near[j] = drawing_distance(rects[i], rects[j]) < distance_threshold
        near &= ~used
        members = [i] + np.nonzero(near)[0].tolist()
        used |= near
        if len(members) >= min_members:
            clusters.append(members)
    return clusters


def _cluster_rects_scalar(rects, distance_threshold, min_members) -> List[List[int]]:
    """The reference's loops verbatim in structure (:3571-3592); also the checker of the vectorised form."""
    clusters: List[List[int]] = []
    used = set()
    for i, r1 in enumerate(rects):
        if i in used:
            continue
        members = [i]
        used.add(i)
        for j, r2 in enumerate(rects):
            if j in used or j == i:
                continue
            if drawing_distance(r1, r2) < distance_threshold:
                members.append(j)
                used.add(j)
        if len(members) >= min_members:
            clusters.append(members)
    return clusters


def regions_from_rects(rects: Sequence[Sequence[float]], page_width: float, page_height: float,
                       method: str = "drawing_based", noun: str = "drawing commands") -> List[Dict]:
    """Cluster -> padded bbox -> area filter -> region dicts (_detect_by_drawings, :3511-3557)."""
    regions = []
    for members in cluster_rects(rects):
        x0 = min(rects[i][0] for i in members); y0 = min(rects[i][1] for i in members)
        x1 = max(rects[i][2] for i in members); y1 = max(rects[i][3] for i in members)
        x0 = max(0, x0 - REGION_PADDING); y0 = max(0, y0 - REGION_PADDING)
        x1 = min(page_width, x1 + REGION_PADDING); y1 = min(page_height, y1 + REGION_PADDING)
        bbox = BoundingBox(x0=x0, y0=y0, x1=x1, y1=y1, page_width=page_width, page_height=page_height)
        area = bbox.area()
        if DRAWING_MIN_AREA < area < page_width * page_height * DRAWING_MAX_PAGE_FRACTION:
            regions.append({"bbox": bbox, "caption": None, "detection_method": method,
                            "notes": f"Detected from {len(members)} {noun}"})
    return regions


def validate_region(bbox: BoundingBox, crop_width_px: int, crop_height_px: int, variance: float,
                    page_height: float, has_caption: bool = False) -> Tuple[float, str]:
    """Score a candidate exactly like _validate_embedded_image (:2933-2998); the grey variance comes from
    the GPU's exact integer moments instead of np.var(PIL 'L')."""
    score = 0.0
    notes = []
    area = bbox.area()
    if area < MIN_AREA:
        return 0.0, "too_small"
    elif area > GOOD_AREA:
        score += 0.3
        notes.append("good_size")
    else:
        score += 0.1
        notes.append("moderate_size")
    if crop_width_px < MIN_DIM_PX or crop_height_px < MIN_DIM_PX:
        return 0.0, "tiny_dimensions"
    if crop_width_px > SUBSTANTIAL_DIM_PX and crop_height_px > SUBSTANTIAL_DIM_PX:
        score += 0.2
        notes.append("substantial_dimensions")
    aspect = crop_width_px / crop_height_px if crop_height_px > 0 else 1.0
    if 0.2 < aspect < 5.0:
        score += 0.2
        notes.append("good_aspect_ratio")
    else:
        score -= 0.1
        notes.append("unusual_aspect_ratio")
    y_position = bbox.y0 / page_height
    if y_position < 0.1 or y_position > 0.9:
        score -= 0.2
        notes.append("likely_header_footer")
    else:
        score += 0.1
        notes.append("good_position")
    if has_caption:
        score += 0.4
        notes.append("has_caption")
    if variance < 10:
        score -= 0.3
        notes.append("low_variance")
    elif variance > 100:
        score += 0.2
        notes.append("good_content_variance")
    return min(score, 1.0), ", ".join(notes)


def resolve_conflict(embedded_bbox: BoundingBox, embedded_confidence: float, embedded_variance: Optional[float],
                     caption_bbox: BoundingBox, caption_text: Optional[str], drawings_in_caption_region: int = 0,
                     embedded_is_raster: bool = True) -> Tuple[str, str]:
    """The five-factor vote of _resolve_conflict (:3041-3103).  Factor 3 takes the variance of the embedded
    crop (GPU moments); factor 4 takes a count supplied by the caller (0 when no vector layer is available)."""
    reasons = []
    emb, cap = 0, 0
    if caption_text:
        cap += 3
        reasons.append("caption_based has caption")
    ea, ca = embedded_bbox.area(), caption_bbox.area()
    if ca > ea * 1.2:
        cap += 2
        reasons.append("caption_based includes more context")
    elif ea > ca * 1.2:
        emb += 1
        reasons.append("embedded is larger")
    if embedded_is_raster and embedded_variance is not None and embedded_variance > 1000:
        emb += 2
        reasons.append("embedded is photo-like (raster)")
    if drawings_in_caption_region > 10:
        cap += 2
        reasons.append("many vector drawings (chart/diagram)")
    if embedded_confidence > 0.7:
        emb += 1
        reasons.append(f"embedded has high validation ({embedded_confidence:.2f})")
    return ("keep_caption" if cap > emb else "keep_embedded"), "; ".join(reasons)


def caption_near_region(draw_bbox: BoundingBox, caption_bbox: Sequence[float]) -> bool:
    """Caption top-left inside x-range and within +-50 pt of the region's bottom (_detect_visual_regions, :3136-3142)."""
    return draw_bbox.x0 <= caption_bbox[0] <= draw_bbox.x1 and draw_bbox.y1 - 50 <= caption_bbox[1] <= draw_bbox.y1 + 50


def merge_visual_regions(primary: List[Dict], secondary: List[Dict]) -> List[Dict]:
    """_detect_visual_regions' merge (:3122-3144): keep every primary region, add a secondary region unless it
    duplicates one already kept (>50 % of its own area) or sits right above a primary region's caption."""
    out = list(primary)
    for reg in secondary:
        if overlaps_with_existing(reg["bbox"], out):
            continue
        if any("caption_bbox" in p and caption_near_region(reg["bbox"], p["caption_bbox"]) for p in primary):
            continue
        out.append(reg)
    return out


def drawings_in_region(bbox: BoundingBox, drawing_rects: Optional[Sequence[Sequence[float]]]) -> int:
    """Factor 4's count (:3080-3089): drawing rects whose TOP-LEFT corner lies inside the caption-based box."""
    if not drawing_rects:
        return 0
    return sum(1 for d in drawing_rects if bbox.x0 <= d[0] <= bbox.x1 and bbox.y0 <= d[1] <= bbox.y1)


def resolve_page_conflicts(caption_regions: Sequence[Dict], candidates: Sequence[Dict],
                           drawing_rects: Optional[Sequence[Sequence[float]]] = None) -> List[Dict]:
    """Pass 2 of _extract_images_from_page (:2822-2847) on region dicts.

    caption_regions: the page's caption-based regions (pass 1; kept with confidence 0.9, :2787-2799).
    candidates     : validated raster regions in detection order -- the role `_extract_embedded_images_validated`'s
                     embedded images play (dicts with 'bbox', 'confidence', 'variance').
    Every candidate is compared with the segments kept SO FAR (caption-based ones and candidates added before it):
    the first one overlapping it by more than 0.4 of the smaller area is the conflict (`find_conflicting`); the
    five-factor vote (`resolve_conflict`) then either replaces that segment by the candidate or discards the candidate.
    Returns the final list in the reference's order (survivors of pass 1, then added candidates)."""
    segments = list(caption_regions)
    for cand in candidates:
        conflict = find_conflicting(cand["bbox"], segments, key=lambda r: r["bbox"])
        if conflict is None:
            segments.append(cand)
            continue
        decision, reason = resolve_conflict(cand["bbox"], cand["confidence"], cand.get("variance"), conflict["bbox"], conflict.get("caption"),
                                            drawings_in_region(conflict["bbox"], drawing_rects))
        if decision == "keep_embedded":
            segments.remove(conflict)
            segments.append(dict(cand, conflict_resolution=reason))
    return segments
