"""Import stub used only by tests/golden/make_golden.py to import the reference (pasero/evaluation.py:8,18)."""
from . import metrics  # noqa: F401


def __getattr__(name):
    raise AttributeError(f'sacrebleu stub: {name} is not available')
