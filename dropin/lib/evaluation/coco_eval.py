"""Drop-in for `lib/evaluation/coco_eval.py` of the reference (bbox): same import names, served by egtr_b200.oi_evaluation."""
from egtr_b200.oi_evaluation import CocoEvaluator  # noqa: F401
