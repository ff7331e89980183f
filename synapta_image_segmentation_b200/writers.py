"""Output writers of the stage: `{book}_visual_segments.json` and `{book}_visual_summary.csv`.

Mirrors the reference's `_initialize_json_file`, `_append_segment_to_json`, `_save_results` and `_save_summary_csv`
(pdf_image_segmentation.py:3852-3952): same file names, same JSON layout (`book_id`, `pdf_path`, `total_segments`,
`segments` = `VisualSegment.to_dict()` records, indent 2, non-ASCII kept), same CSV columns and truncations, the same
duplicate rule (a segment id is written once, :3886-3887).  Byte-identical output on the same segments is pinned by
tests/golden/reference_writers.json.

One deliberate difference: the reference re-reads and re-writes the whole JSON file for every appended segment
(O(n^2) over a book, :3866-3898).  Here `append` only records the segment; the file is rewritten by `flush()` --
called every `flush_every` appends and by `save_results()` -- through a temporary file and an atomic rename, so a
crash never leaves a truncated JSON behind.
"""
from __future__ import annotations

import csv
import io
import json
import os
from pathlib import Path
from typing import Dict, Iterable, List, Optional

from .datamodel import VisualSegment


class SegmentWriter:
    def __init__(self, book_id: str, pdf_path: str, output_dir, flush_every: int = 64):
        self.book_id = book_id
        self.pdf_path = pdf_path
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.output_json = self.output_dir / f"{book_id}_visual_segments.json"
        self.output_csv = self.output_dir / f"{book_id}_visual_summary.csv"
        self.flush_every = max(1, int(flush_every))
        self.segments: List[VisualSegment] = []
        self._ids: Dict[str, int] = {}
        self._dirty = 0

    # ---- JSON ---------------------------------------------------------------------------------------
    def _document(self) -> dict:
        return {"book_id": self.book_id, "pdf_path": self.pdf_path, "total_segments": len(self.segments),
                "segments": [s.to_dict() for s in self.segments]}

    def _write_json(self, doc: dict) -> None:
        tmp = self.output_json.with_suffix(".json.tmp")
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(doc, f, indent=2, ensure_ascii=False)
        os.replace(tmp, self.output_json)

    def initialize(self) -> None:
        """`_initialize_json_file` (:3852-3864): truncate to the empty document."""
        self.segments.clear()
        self._ids.clear()
        self._dirty = 0
        self._write_json(self._document())

    def append(self, segment: VisualSegment) -> bool:
        """`_append_segment_to_json` (:3866-3898).  Returns False for a duplicate segment id."""
        if segment.segment_id in self._ids:
            return False
        self._ids[segment.segment_id] = len(self.segments)
        self.segments.append(segment)
        self._dirty += 1
        if self._dirty >= self.flush_every:
            self.flush()
        return True

    def extend(self, segments: Iterable[VisualSegment]) -> int:
        return sum(1 for s in segments if self.append(s))

    def flush(self) -> None:
        self._write_json(self._document())
        self._dirty = 0

    def save_results(self) -> None:
        """`_save_results` (:3900-3931): final JSON, then the summary CSV."""
        self.flush()
        self.save_summary_csv()

    # ---- CSV ----------------------------------------------------------------------------------------
    @staticmethod
    def summary_row(seg: VisualSegment) -> Dict[str, object]:
        """One row of `_save_summary_csv` (:3933-3952)."""
        return {"segment_id": seg.segment_id, "page": seg.page_no, "type": seg.segment_type.value,
                "confidence": f"{seg.classification_confidence:.2f}", "figure_number": seg.figure_number or "",
                "caption": seg.caption_text[:100] if seg.caption_text else "",
                "ocr_text": seg.ocr_result.raw_text[:100] if seg.ocr_result else "",
                "linked_concepts": len(seg.linked_concept_ids), "summary": seg.summary[:100] if seg.summary else ""}

    COLUMNS = ["segment_id", "page", "type", "confidence", "figure_number", "caption", "ocr_text", "linked_concepts", "summary"]

    def summary_csv_text(self) -> str:
        """Same bytes as pandas' `DataFrame(rows).to_csv(index=False)`: minimal quoting, '\\n' line ends."""
        if not self.segments:
            return "\n"                       # pandas writes an empty header line for an empty frame
        buf = io.StringIO()
        w = csv.writer(buf, quoting=csv.QUOTE_MINIMAL, lineterminator="\n")
        w.writerow(self.COLUMNS)
        for s in self.segments:
            row = self.summary_row(s)
            w.writerow([row[c] for c in self.COLUMNS])
        return buf.getvalue()

    def save_summary_csv(self) -> Path:
        tmp = self.output_csv.with_suffix(".csv.tmp")
        with open(tmp, "w", encoding="utf-8", newline="") as f:
            f.write(self.summary_csv_text())
        os.replace(tmp, self.output_csv)
        return self.output_csv


def load_segments_json(path) -> Optional[dict]:
    p = Path(path)
    if not p.exists():
        return None
    with open(p, "r", encoding="utf-8") as f:
        return json.load(f)
