"""symbanafis_b200 -- B200-native batch evaluator for SymbAnaFis compiled expressions.

The reference's Python surface for this path, under the same names (`symbanafis_b200/evaluator.py`):

    from symbanafis_b200 import CompiledEvaluator, eval_f64, evaluate_parallel

Resolved lazily: importing the package does not load the CUDA library.
"""
_EXPORTS = {"CompiledEvaluator": "evaluator", "eval_f64": "evaluator", "evaluate_parallel": "evaluator"}


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'symbanafis_b200' has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)


def __dir__():
    return sorted(list(globals()) + list(_EXPORTS))
