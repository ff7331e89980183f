"""Generates tests/golden/reference_writers.json by importing the reference in THIS container.

    python tests/golden/make_schema_golden.py

Pins the output writers: the key layout of the shipped run artefact
(/root/reference/extracted_visuals_excelSS: JSON + CSV) and, byte for byte, what the reference's own
`_save_results` / `_save_summary_csv` (pdf_image_segmentation.py:3900-3952) write for a small list of
constructed segments (one per visual type, default detail dataclasses)."""
import json
import os
import sys
import tempfile
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

ref = ref_import.load()
A1 = os.path.join(ref_import.REFERENCE_DIR, "extracted_visuals_excelSS")


def segments(mod):
    """The same constructed segments for the reference module and for ours (fields set at detection time + hints)."""
    BB, VT = mod.BoundingBox, mod.VisualType
    segs = []
    segs.append(mod.VisualSegment(segment_id="textbook_001_p000_aaaaaaaa", segment_type=VT.CHART, book_id="textbook_001", page_no=1,
                                  bbox=BB(72.0, 100.5, 300.25, 260.0, 612.0, 792.0), image_path="out/textbook_001_p000_aaaaaaaa.png",
                                  extraction_method="raster_cc", confidence=0.8, notes="Validation: good_size",
                                  chart_data=mod.ChartSpecificData(chart_subtype="bar", grid_detected=True, color_scheme=["#112233", "#445566"],
                                                                   estimated_data_points=23),
                                  classification_confidence=0.5))
    segs.append(mod.VisualSegment(segment_id="textbook_001_p001_bbbbbbbb", segment_type=VT.DIAGRAM, book_id="textbook_001", page_no=2,
                                  bbox=BB(10.0, 20.0, 110.0, 220.0, 612.0, 792.0), caption_text="Figure 2.1 A diagram, with \"quotes\"\nand a newline",
                                  figure_number="2.1", extraction_method="raster_cluster", confidence=0.9, notes="Validation: good_size, good_position",
                                  diagram_data=mod.DiagramSpecificData(arrow_count=4, shapes_detected={"rectangles": 3, "circles": 1, "diamonds": 0},
                                                                       connections=[{"id": "conn_0", "type": "arrow"}]),
                                  summary="s" * 150))
    segs.append(mod.VisualSegment(segment_id="textbook_001_p002_cccccccc", segment_type=VT.IMAGE, book_id="textbook_001", page_no=3,
                                  bbox=BB(0.0, 0.0, 612.0, 792.0, 612.0, 792.0), extraction_method="embedded_image",
                                  image_data=mod.ImageSpecificData(image_subtype="photo", dominant_colors=["#a0b0c0"])))
    segs.append(mod.VisualSegment(segment_id="textbook_001_p003_dddddddd", segment_type=VT.FIGURE, book_id="textbook_001", page_no=4,
                                  bbox=BB(1.5, 2.5, 3.5, 4.5, 612.0, 792.0),
                                  figure_data=mod.FigureSpecificData(contains_chart=True, contains_image=True)))
    segs.append(mod.VisualSegment(segment_id="textbook_001_p004_eeeeeeee", segment_type=VT.UNKNOWN, book_id="textbook_001", page_no=5,
                                  bbox=BB(5.0, 6.0, 7.0, 8.0, 612.0, 792.0)))
    return segs


def main():
    a1 = json.load(open(os.path.join(A1, "textbook_001_visual_segments.json"), encoding="utf-8"))
    seg = a1["segments"][0]
    out = dict(a1_top_keys=list(a1.keys()), a1_segment_keys=list(seg.keys()), a1_bbox_keys=list(seg["bbox"].keys()),
               a1_image_details_keys=list(seg["image_details"].keys()), a1_image_data_keys=list(seg["image_data"].keys()),
               a1_notes=seg["notes"], a1_confidence=seg["confidence"], a1_extraction_method=seg["extraction_method"],
               a1_csv_header=open(os.path.join(A1, "textbook_001_visual_summary.csv"), encoding="utf-8").readline().strip())
    pl = ref.VisualSegmentationPipeline.__new__(ref.VisualSegmentationPipeline)      # __init__ is never run
    with tempfile.TemporaryDirectory() as td:
        pl.book_id, pl.pdf_path, pl.output_dir = "textbook_001", "book.pdf", Path(td)
        pl.output_json = Path(td) / "textbook_001_visual_segments.json"
        pl.segments = segments(ref)
        pl._initialize_json_file()
        out["initialized_json"] = pl.output_json.read_text(encoding="utf-8")
        pl._save_results()
        out["json"] = pl.output_json.read_text(encoding="utf-8")
        out["csv"] = (Path(td) / "textbook_001_visual_summary.csv").read_bytes().decode("utf-8")
    json.dump(out, open(os.path.join(HERE, "reference_writers.json"), "w"), indent=1)
    print("wrote reference_writers.json:", len(out["json"]), "bytes of JSON,", len(out["csv"]), "bytes of CSV")


if __name__ == "__main__":
    main()
