"""Detection data model -- the reference's output contract for this stage.

Field names, defaults and the ``to_dict`` key layout follow the reference so that records produced
here are interchangeable with the reference's (pdf_image_segmentation.py:35-295; the one shipped
output, extracted_visuals_excelSS/*.json, pins the JSON shape).  Only the fields the detection and
hint stages fill are given behaviour; LLM / OCR / concept-linking fields exist as plain storage.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class VisualType(str, Enum):
    """pdf_image_segmentation.py:35-42"""
    FIGURE = "figure"
    CHART = "chart"
    DIAGRAM = "diagram"
    FLOWCHART = "flowchart"
    IMAGE = "image"
    UNKNOWN = "unknown"


@dataclass
class ChartSpecificData:
    """pdf_image_segmentation.py:44-55; grid_detected / color_scheme / estimated_data_points are hint outputs."""
    chart_subtype: Optional[str] = None
    axes_info: Dict[str, Any] = field(default_factory=dict)
    value_ranges: Dict[str, Tuple[float, float]] = field(default_factory=dict)
    legend_items: List[str] = field(default_factory=list)
    series_count: int = 0
    grid_detected: bool = False
    color_scheme: List[str] = field(default_factory=list)
    estimated_data_points: int = 0
    tick_labels: Dict[str, List[str]] = field(default_factory=dict)


@dataclass
class DiagramSpecificData:
    """pdf_image_segmentation.py:58-69"""
    diagram_subtype: Optional[str] = None
    node_count: int = 0
    nodes: List[Dict[str, Any]] = field(default_factory=list)
    connections: List[Dict[str, Any]] = field(default_factory=list)
    arrow_count: int = 0
    hierarchy_detected: bool = False
    layout_type: Optional[str] = None
    shapes_detected: Dict[str, int] = field(default_factory=dict)
    has_decision_points: bool = False


@dataclass
class ImageSpecificData:
    """pdf_image_segmentation.py:72-89"""
    image_subtype: Optional[str] = None
    contains_text: bool = False
    text_density: str = "none"
    is_embedded_table: bool = False
    dominant_colors: List[str] = field(default_factory=list)
    estimated_content_type: Optional[str] = None
    definitions: List[Dict[str, str]] = field(default_factory=list)
    formulas: List[Dict[str, str]] = field(default_factory=list)
    variables: List[Dict[str, str]] = field(default_factory=list)
    tables: List[Dict[str, Any]] = field(default_factory=list)
    input_variables: List[Dict[str, Any]] = field(default_factory=list)
    output_values: List[Dict[str, Any]] = field(default_factory=list)
    calculation_verification: Optional[Dict[str, Any]] = None


@dataclass
class FigureSpecificData:
    """pdf_image_segmentation.py:92-99"""
    is_composite: bool = False
    sub_figure_count: int = 0
    contains_chart: bool = False
    contains_diagram: bool = False
    contains_image: bool = False


@dataclass
class BoundingBox:
    """Box in PDF points, top-left origin, plus the page size (pdf_image_segmentation.py:101-122)."""
    x0: float
    y0: float
    x1: float
    y1: float
    page_width: float
    page_height: float

    def to_dict(self) -> Dict[str, float]:
        return {"x0": self.x0, "y0": self.y0, "x1": self.x1, "y1": self.y1,
                "width": self.x1 - self.x0, "height": self.y1 - self.y0,
                "page_width": self.page_width, "page_height": self.page_height}

    def area(self) -> float:
        return (self.x1 - self.x0) * (self.y1 - self.y0)

    def to_pixels(self, dpi: float) -> Tuple[int, int, int, int]:
        """(x, y, w, h) of the crop in a page raster at `dpi` (px = pt * dpi / 72, pdf_image_segmentation.py:3649)."""
        s = dpi / 72.0
        x0, y0 = int(round(self.x0 * s)), int(round(self.y0 * s))
        x1, y1 = int(round(self.x1 * s)), int(round(self.y1 * s))
        return x0, y0, max(1, x1 - x0), max(1, y1 - y0)


@dataclass
class OCRResult:
    """pdf_image_segmentation.py:125-139 (OCR itself is out of scope; detected_arrows is a hint output)."""
    raw_text: str = ""
    blocks: List[Dict[str, Any]] = field(default_factory=list)
    confidence: float = 0.0
    axis_labels: Dict[str, str] = field(default_factory=dict)
    legend_items: List[str] = field(default_factory=list)
    tick_labels: Dict[str, List[str]] = field(default_factory=dict)
    node_texts: List[str] = field(default_factory=list)
    detected_arrows: int = 0


@dataclass
class VisualSegment:
    """The per-region record (pdf_image_segmentation.py:151-205)."""
    segment_id: str
    segment_type: VisualType
    book_id: str
    page_no: int
    bbox: BoundingBox
    image_path: Optional[str] = None
    image_bytes: Optional[bytes] = None
    caption_text: Optional[str] = None
    figure_number: Optional[str] = None
    reference_keys: List[str] = field(default_factory=list)
    ocr_result: Optional[OCRResult] = None
    mermaid_repr: Optional[Any] = None
    chart_data: Optional[ChartSpecificData] = None
    diagram_data: Optional[DiagramSpecificData] = None
    image_data: Optional[ImageSpecificData] = None
    figure_data: Optional[FigureSpecificData] = None
    extracted_text_structured: Dict[str, List[str]] = field(default_factory=dict)
    classification_confidence: float = 0.0
    classification_method: str = "heuristic"
    summary: Optional[str] = None
    summary_confidence: float = 0.0
    linked_concept_ids: List[Dict[str, Any]] = field(default_factory=list)
    heading_path: List[str] = field(default_factory=list)
    linked_segment_ids: List[str] = field(default_factory=list)
    nearby_text: Optional[str] = None
    extraction_method: str = "native"
    confidence: float = 1.0
    notes: str = ""

    @staticmethod
    def _plain(obj):
        """numpy scalars / arrays -> Python natives, recursively (pdf_image_segmentation.py:207-225)."""
        if isinstance(obj, np.bool_):
            return bool(obj)
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, dict):
            return {k: VisualSegment._plain(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return [VisualSegment._plain(v) for v in obj]
        return obj

    def to_dict(self) -> Dict[str, Any]:
        """JSON record with the reference's key layout (pdf_image_segmentation.py:227-295)."""
        d = asdict(self)
        d["segment_type"] = self.segment_type.value
        d["bbox"] = self.bbox.to_dict() if self.bbox else None
        d.pop("image_bytes", None)
        c, g, im, f = self.chart_data, self.diagram_data, self.image_data, self.figure_data
        if c:
            d["chart_details"] = {"subtype": c.chart_subtype, "axes": c.axes_info, "legend": c.legend_items,
                                  "series_count": c.series_count, "data_points": c.estimated_data_points,
                                  "has_grid": c.grid_detected, "colors": c.color_scheme,
                                  "value_ranges": c.value_ranges, "tick_labels": c.tick_labels}
        if g:
            d["diagram_details"] = {"subtype": g.diagram_subtype, "node_count": g.node_count, "nodes": g.nodes[:15],
                                    "connection_count": len(g.connections), "arrow_count": g.arrow_count,
                                    "layout_type": g.layout_type, "has_hierarchy": g.hierarchy_detected,
                                    "has_decision_points": g.has_decision_points, "shapes": g.shapes_detected}
        if im:
            d["image_details"] = {"subtype": im.image_subtype, "contains_text": im.contains_text,
                                  "text_density": im.text_density, "is_embedded_table": im.is_embedded_table,
                                  "content_type": im.estimated_content_type, "dominant_colors": im.dominant_colors[:5],
                                  "definitions": im.definitions, "formulas": im.formulas, "variables": im.variables,
                                  "tables": im.tables, "input_variables": im.input_variables,
                                  "output_values": im.output_values,
                                  "calculation_verification": im.calculation_verification}
        if f:
            d["figure_details"] = {"is_composite": f.is_composite, "sub_figure_count": f.sub_figure_count,
                                   "contains_chart": f.contains_chart, "contains_diagram": f.contains_diagram,
                                   "contains_image": f.contains_image}
        if self.extracted_text_structured:
            d["extracted_text_structured"] = self.extracted_text_structured
        return self._plain(d)
