"""Drop-in for `lib/evaluation/sg_eval.py` of the reference: same import names, served by egtr_b200.evaluation."""
from egtr_b200.evaluation import *  # noqa: F401,F403
from egtr_b200.evaluation import BasicSceneGraphEvaluator, calculate_mR_from_evaluator_list  # noqa: F401
