"""lp.translation_unit.TranslationUnit, used in type annotations of annotation scripts (reference tests/sem.py:11-14)."""
from nomp_bridge.ir import Kernel as TranslationUnit  # noqa: F401
