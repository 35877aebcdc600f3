"""B200-native STMiner spatial-pattern hot path.  ``from stminer_b200 import SPFinder`` mirrors
``from STMiner import SPFinder`` (reference: STMiner/__init__.py:1)."""
from .SPFinder import SPFinder  # noqa: F401

__all__ = ["SPFinder"]
