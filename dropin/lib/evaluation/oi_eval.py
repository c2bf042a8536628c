"""Drop-in for `lib/evaluation/oi_eval.py` of the reference: same import names, served by egtr_b200.oi_evaluation."""
from egtr_b200.oi_evaluation import OIEvaluator, eval_rel_results  # noqa: F401
