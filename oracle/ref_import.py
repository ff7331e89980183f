"""Import the reference in place (build container only) -- TEST INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box; nothing under ``-m gpu`` tests, ``smoke()`` or
``bench.py`` calls this.  It exists to (a) pin the oracle against the reference's own functions in
CPU tests when the reference tree is mounted and (b) generate ``tests/golden`` fixtures.

Recipe: SURVEY.md Appendix C.  ``fitz`` (PyMuPDF) and ``paddleocr`` are absent from the image and
are imported unconditionally by the reference (pdf_image_segmentation.py:18-19), so two stub modules
are seeded into ``sys.modules`` first.  ``VisualSegmentationPipeline.__init__`` is never run
(it creates directories and embeds a credential literal, pdf_image_segmentation.py:2707).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_DIR = os.environ.get("SYNSEG_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_DIR, "pdf_image_segmentation.py"))


def load(old_algo: bool = False):
    """Return the imported reference module (current algorithm, or the old one)."""
    if not available():
        raise FileNotFoundError(REFERENCE_DIR)
    if "fitz" not in sys.modules:
        fitz = types.ModuleType("fitz")
        for n in ("Page", "Rect", "Document", "Matrix"):
            setattr(fitz, n, object)
        sys.modules["fitz"] = fitz
    if "paddleocr" not in sys.modules:
        pd_ = types.ModuleType("paddleocr")
        pd_.PaddleOCR = object
        sys.modules["paddleocr"] = pd_
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    return importlib.import_module("pdf_image_segmentation_old_algo" if old_algo else "pdf_image_segmentation")


def pipeline_instance(ref=None):
    """An uninitialised VisualSegmentationPipeline (pure-geometry methods are callable on it)."""
    ref = ref or load()
    return ref.VisualSegmentationPipeline.__new__(ref.VisualSegmentationPipeline)
